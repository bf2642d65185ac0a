"""Group-convolution networks (tcgen05 GEMM with the implicit 13-neighbour gather + pack / tails) against the oracle on seeded random
weights, and against files the UNMODIFIED reference wrote (tests/golden/s256.npz: gf_eqv_*, det_score_*, trans_pre_*)."""
import os
import types
import numpy as np
import pytest
import torch
from conftest import load_golden
from oracle import roreg_oracle as O
from roreg_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from roreg_b200 import ops
    c = ops.Context(0)
    yield c
    c.close()


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("npass,tol", [(3, 5e-5), (1, 3e-3)])   # 3xTF32: float32-class products; the tensor core's accumulator truncation grows with K
@pytest.mark.parametrize("R,K,Odim", [(300, 416, 256), (129, 64, 4), (1000, 3328, 512), (60, 6656, 256)])
def test_gemm_tc_against_float64(ctx, npass, tol, R, K, Odim):
    from roreg_b200 import nets
    rng = np.random.default_rng(R + K)
    A = rng.standard_normal((R, K)).astype(np.float32); W = (rng.standard_normal((Odim, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(Odim).astype(np.float32); res = rng.standard_normal((R, Odim)).astype(np.float32)
    sc = (1 + 0.1 * rng.standard_normal(Odim)).astype(np.float32); sh = (0.1 * rng.standard_normal(Odim)).astype(np.float32)
    g = nets.GroupNets(ctx, npass)
    L = nets.Layer(ctx, W.reshape(Odim, K, 1, 1), b)
    hi, lo = nets.tf32_split(A)
    raw, (ahi, alo) = g.gemm((ctx.dev(hi), ctx.dev(lo)), R, L, residual=ctx.dev(res), res_ld=Odim, want_raw=True, bn=(sc, sh), relu=True)
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ W.astype(np.float64).T + b + res
    assert np.abs(_np(raw) - ref).max() < tol * max(1.0, np.abs(ref).max())
    act = np.maximum(ref * sc + sh, 0)
    assert np.abs(_np(ahi) + _np(alo) - act).max() < tol * max(1.0, np.abs(act).max())
    assert (np.abs(_np(alo)) <= np.abs(_np(ahi)) * 2.0 ** -10 + 1e-30).all()          # hi carries 11 significant bits


def test_gf_net_against_oracle_and_reference(ctx, tables):
    from roreg_b200 import nets
    z, n, keynum, max_iter, seeds = load_golden("s256")
    ds = synth.SynthDataset(seeds[:1], n=n, name="synth/s256", max_res_deg=2.0)
    sd = O.random_state_dict("GF", 101)
    net = nets.GFNet(ctx, sd, npass=3, chunk=100)
    for cid in ds.pc_ids:
        x = ds.get_feats(cid, "fcgf")
        got = _np(net.forward(ctx.dev(x[:140])))
        ref, _ = O.gf_forward(x[:140], sd, tables.nei)
        assert np.abs(got - ref).max() < 1e-4                        # stated tolerance on descriptors: 1e-4 (measured 4e-5)
        assert np.abs(got[:40] - z[f"gf_eqv_{cid}"]).max() < 1e-4    # vs the file the reference's yoho_des.run wrote
    # equivariance known answer (SURVEY 8c ii): permuting the input group axis by P[a] permutes the output
    x = ds.get_feats("0", "fcgf")[:64]
    a = 23
    ya = _np(net.forward(ctx.dev(np.ascontiguousarray(x[:, :, tables.perm[a]]))))
    y = _np(net.forward(ctx.dev(x)))
    assert np.abs(ya - y[:, :, tables.perm[a]]).max() < 1e-4


def test_gf_net_single_pass_tf32(ctx, tables):
    from roreg_b200 import nets
    sd = O.random_state_dict("GF", 7)
    x = synth.make_pair(5, n=64, with_fcgf=True)["fcgf0"]
    got = _np(nets.GFNet(ctx, sd, npass=1).forward(ctx.dev(x)))
    ref, _ = O.gf_forward(x, sd, tables.nei)
    assert np.abs(got - ref).max() < 5e-3            # one TF32 pass: the reference's own cuDNN arithmetic class, not parity-grade


@pytest.mark.parametrize("npass", [1, 3])
def test_gather_producers_agree_bitwise(ctx, npass, monkeypatch):
    """The implicit group convolution's two operand producers (cp.async row copies - the default - and TMA tile::gather4) fill the
    same shared-memory image: every layer output must be IDENTICAL, at a size with several tiles per CTA, ragged last tile, and
    three chunks (the cp.async data reaches the tensor core through a proxy fence: a missed fence would show here as stale rows)."""
    from roreg_b200 import nets
    sd = O.random_state_dict("GF", 33)
    rng = np.random.default_rng(12)
    x = rng.standard_normal((1111, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True)
    xd = ctx.dev(x)
    net = nets.GFNet(ctx, sd, npass=npass, chunk=400)
    monkeypatch.delenv("ROREG_GEMM_GATHER", raising=False)
    a = [net.forward(xd).clone() for _ in range(3)]
    monkeypatch.setenv("ROREG_GEMM_GATHER", "tma")
    b = net.forward(xd).clone()
    monkeypatch.delenv("ROREG_GEMM_GATHER", raising=False)
    assert torch.equal(a[0], a[1]) and torch.equal(a[0], a[2])      # run-to-run
    assert torch.equal(a[0], b)                                      # producer-to-producer
    assert bool(torch.isfinite(a[0]).all())


def test_et_net_against_oracle(ctx, tables):
    from roreg_b200 import nets
    pr = synth.make_pair(33, n=500, with_fcgf=True, max_res_deg=2.0)
    pps, _ = O.mutual_run(pr["feats0"], pr["feats1"])
    dr = O.rindex(pr["feats0"], pr["feats1"], pps, tables.perm)
    sd = O.random_state_dict("ET", 102)
    net = nets.ETNet(ctx, sd, npass=3, chunk=128)
    i0 = ctx.dev(pps[:, 0].astype(np.int32)); i1 = ctx.dev(pps[:, 1].astype(np.int32))
    q = _np(net.forward(ctx.dev(pr["fcgf1"]), i1, ctx.dev(pr["fcgf0"]), i0, ctx.dev(pr["feats1"]), i1, ctx.dev(pr["feats0"]), i0,
                        ctx.dev(dr.astype(np.int32))))
    ref = O.et_forward(pr["fcgf1"][pps[:, 1]], pr["fcgf0"][pps[:, 0]], pr["feats1"][pps[:, 1]], pr["feats0"][pps[:, 0]], dr, sd,
                       tables.nei, tables.perm)
    assert np.abs(q - ref).max() < 2e-5
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-6


def test_rd_net_against_oracle(ctx, tables):
    from roreg_b200 import nets
    sd = O.random_state_dict("RD", 103)
    x = synth.make_pair(44, n=300)["feats0"]
    s = _np(nets.RDNet(ctx, sd, npass=3, chunk=128).forward(ctx.dev(x)))
    ref = O.rd_forward(x, sd, tables.nei, tables.perm)
    assert np.abs(s - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())


def _write_ckpt(path, sd):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    full = {}
    for k, v in sd.items():
        full[k] = torch.from_numpy(v)
    torch.save({"best_para": 0, "network_state_dict": full}, path)


def test_net_plugins_reproduce_reference_files(tmp_path):
    """yoho_des.run / yoho_det.run / extractor_localtrans.Rt_pre through the reference signatures, against the
    files the reference wrote with the same (seeded random) checkpoints."""
    import roreg_b200.test as rt
    z, n, keynum, max_iter, seeds = load_golden("s256")
    ds = synth.SynthDataset(seeds, n=n, name="synth/s256", max_res_deg=2.0)
    model_fn = str(tmp_path / "ckpt")
    for kind, seed in (("GF", 101), ("ET", 102), ("RD", 103)):
        _write_ckpt(f"{model_fn}/{kind}/model_best.pth", O.random_state_dict(kind, seed))
    # per-cloud stages on clouds 0, 1 (what the fixture recorded)
    cache_pc = str(tmp_path / "cache_pc")
    ds.write_cache(cache_pc, yoho=False)
    cfg = types.SimpleNamespace(output_cache_fn=cache_pc, model_fn=model_fn, SO3_related_files=None, backbone="FCGF", bs_GF=1250,
                                bs_ET=1000, RD=False, RM=False, match_n=0.5, ransac_ird=0.1)
    ds_pc = synth.SynthDataset(seeds[:1], n=n, name="synth/s256", max_res_deg=2.0)
    rt.yoho_des(cfg).run(ds_pc)
    rt.yoho_det(cfg).run(ds_pc)
    for cid in ds_pc.pc_ids:
        eqv = np.load(f"{cache_pc}/synth/s256/YOHO_Output_Group_feature/{cid}.npy")
        assert eqv.dtype == np.float32 and eqv.shape == (n, 32, 60)
        assert np.abs(eqv[:40] - z[f"gf_eqv_{cid}"]).max() < 1e-4
        det = np.load(f"{cache_pc}/synth/s256/det_score/{cid}.npy")
        assert np.mean(np.abs(det - z[f"det_score_{cid}"]) * n <= 1.0) > 0.97       # rank statistic: near ties may swap neighbours
    # ET stage on the reference's own matches / DR_index
    cache = str(tmp_path / "cache")
    ds.write_cache(cache)
    cfg.output_cache_fn = cache
    base = f"{cache}/synth/s256/match_{keynum}"
    os.makedirs(f"{base}/DR_index", exist_ok=True); os.makedirs(f"{base}/scores", exist_ok=True)
    for (id0, id1) in ds.pair_ids:
        np.save(f"{base}/{id0}-{id1}.npy", z[f"match_{id0}-{id1}"]); np.save(f"{base}/DR_index/{id0}-{id1}.npy", z[f"dr_index_{id0}-{id1}"])
        np.save(f"{base}/scores/{id0}-{id1}.npy", z[f"scores_{id0}-{id1}"])
    rt.extractor_localtrans(cfg).Rt_pre(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        tr = np.load(f"{base}/Trans_pre/{id0}-{id1}.npy")
        assert tr.dtype == np.float64 and np.abs(tr - z[f"trans_pre_{id0}-{id1}"]).max() < 1e-4
    # the complete yohoo estimator (Rindex -> Rt_pre -> one-shot RANSAC -> refine) against the reference's final poses
    np.random.seed(4321)
    rt.yohoo(cfg).run(ds, keynum, max_iter)
    for (id0, id1) in ds.pair_ids:
        r = np.load(f"{base}/yohoo/{max_iter}iters/{id0}-{id1}.npz")
        assert int(r["recalltime"]) == int(z[f"yohoo_recall_{id0}-{id1}"])
        assert np.abs(r["trans"] - z[f"yohoo_trans_{id0}-{id1}"]).max() < 1e-4


@pytest.mark.parametrize("npass,tol", [(3, 1e-4), (1, 2e-2)])
def test_group_corr_allpairs(ctx, tables, npass, tol):
    """All-pairs 60-rotation correlation (north_star kernel 1): max_a / argmax_a of cor_a(n,m) for every (n,m) against the
    float64 restatement of test/estimator.py:85-89 on every pair, and the rotation-invariant NN per row."""
    import ctypes as C
    from roreg_b200 import nets, _lib
    from roreg_b200.ops import _ptr, _stream
    pr = synth.make_pair(71, n=200)
    X = pr["feats1"][:150]; Y = pr["feats0"][:130]
    g = nets.GroupNets(ctx, npass)
    xh, xl = g.pack([ctx.dev(X)], [None], [0], None, X.shape[0]); yh, yl = g.pack([ctx.dev(Y)], [None], [0], None, Y.shape[0])
    N, M = X.shape[0], Y.shape[0]
    best = torch.empty((N, M), dtype=torch.float32, device=ctx.device); ba = torch.empty((N, M), dtype=torch.uint8, device=ctx.device)
    nn = torch.empty(N, dtype=torch.int32, device=ctx.device); nna = torch.empty(N, dtype=torch.int32, device=ctx.device)
    nd = torch.empty(N, dtype=torch.float32, device=ctx.device)
    rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, _ptr(xh), _ptr(xl), N, _ptr(yh), _ptr(yl), M, npass, _ptr(best), _ptr(ba), _ptr(nn), _ptr(nna),
                                           _ptr(nd), _stream())
    _lib.check(ctx.h, rc, "roreg_group_corr_allpairs")
    torch.cuda.synchronize()
    # float64 reference: cor[n,m,a] = sum_{f,g} X[n,f,P[a,g]] Y[m,f,g]
    Xp = X.astype(np.float64)[:, :, tables.perm]            # [N,32,60(a),60(g)]
    cor = np.einsum("nfag,mfg->nma", Xp, Y.astype(np.float64))
    ref_best = cor.max(2); ref_a = cor.argmax(2)
    # relative tolerance: the tensor core's fp32 accumulator truncates, so the error of a K = 1920 (x3 passes) sum grows with
    # its magnitude (measured 2.5e-5 relative at |cor| = 55, i.e. on true matches; 5e-5 absolute on the bulk)
    assert (np.abs(_np(best) - ref_best) / np.maximum(1.0, np.abs(ref_best))).max() < tol
    top2 = np.sort(cor, axis=2)[:, :, -2:]
    clear = (top2[:, :, 1] - top2[:, :, 0]) > tol * 20
    assert (_np(ba)[clear] == ref_a[clear]).all()
    dist = (X.astype(np.float64) ** 2).sum((1, 2))[:, None] + (Y.astype(np.float64) ** 2).sum((1, 2))[None] - 2 * ref_best
    part = np.partition(dist, 1, axis=1)
    ok = (part[:, 1] - part[:, 0]) > tol * 200
    assert (_np(nn)[ok] == dist.argmin(1)[ok]).all()
    assert np.abs(_np(nd) - dist.min(1))[ok].max() < tol * 200
    # a size that takes the vectorised store path and several tiles in both directions (M % 4 == 0, M % 16 != 0: run c16's bug)
    if True:
        rng = np.random.default_rng(3)
        X3 = rng.standard_normal((300, 32, 60)).astype(np.float32); Y3 = rng.standard_normal((520, 32, 60)).astype(np.float32)
        X3 /= np.linalg.norm(X3, axis=1, keepdims=True); Y3 /= np.linalg.norm(Y3, axis=1, keepdims=True)
        xh3, xl3 = g.pack([ctx.dev(X3)], [None], [0], None, 300); yh3, yl3 = g.pack([ctx.dev(Y3)], [None], [0], None, 520)
        b3 = torch.empty((300, 520), dtype=torch.float32, device=ctx.device); a3 = torch.empty((300, 520), dtype=torch.uint8, device=ctx.device)
        rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, _ptr(xh3), _ptr(xl3), 300, _ptr(yh3), _ptr(yl3), 520, npass, _ptr(b3), _ptr(a3), None, None, None, _stream())
        _lib.check(ctx.h, rc, "roreg_group_corr_allpairs")
        torch.cuda.synchronize()
        cor3 = np.einsum("nfag,mfg->nma", X3.astype(np.float64)[:, :, tables.perm], Y3.astype(np.float64))
        assert (np.abs(_np(b3) - cor3.max(2)) / np.maximum(1.0, np.abs(cor3.max(2)))).max() < tol
        t3 = np.sort(cor3, axis=2)[:, :, -2:]
        clear3 = (t3[:, :, 1] - t3[:, :, 0]) > tol * 20
        assert (_np(a3)[clear3] == cor3.argmax(2)[clear3]).all()
        if npass == 3:
            assert clear3.mean() > 0.9
    # planted rotation: Y rows that are group-permuted copies of X rows must be found with their rotation index
    a = 17
    Y2 = np.ascontiguousarray(X[:64][:, :, tables.perm[a]])
    yh2, yl2 = g.pack([ctx.dev(Y2)], [None], [0], None, 64)
    best2 = torch.empty((N, 64), dtype=torch.float32, device=ctx.device); ba2 = torch.empty((N, 64), dtype=torch.uint8, device=ctx.device)
    rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, _ptr(xh), _ptr(xl), N, _ptr(yh2), _ptr(yl2), 64, npass, _ptr(best2), _ptr(ba2), _ptr(nn), _ptr(nna),
                                           _ptr(nd), _stream())
    _lib.check(ctx.h, rc, "roreg_group_corr_allpairs")
    torch.cuda.synchronize()
    assert (_np(nn)[:64] == np.arange(64)).all() and (_np(nna)[:64] == a).all()       # Des2R(X, X[:,:,P[a]]) = a on the diagonal


def test_yohoo_engine_against_oracle(ctx, tables):
    """Batched yohoo throughput engine (matcher -> Des2R -> ET on the scored hypotheses -> one-shot RANSAC -> refine) against the
    oracle pipeline with the same hypothesis order."""
    from roreg_b200 import pipeline
    seeds = [81, 82]; n = 500; H = 120
    prs = [synth.make_pair(s, n=n, with_fcgf=True, max_res_deg=2.0) for s in seeds]
    sd = O.random_state_dict("ET", 102)
    desc = ctx.dev(np.stack([x for pr in prs for x in (pr["feats0"], pr["feats1"])]))
    fcgf = ctx.dev(np.stack([x for pr in prs for x in (pr["fcgf0"], pr["fcgf1"])]))
    keys = ctx.dev(np.stack([x for pr in prs for x in (pr["keys0"], pr["keys1"])]), torch.float64)
    pc = ctx.dev(np.array([[0, 1], [2, 3]], np.int32))
    rng = np.random.default_rng(0)
    stages = []; order = np.zeros((2, H), np.int64)
    for i, pr in enumerate(prs):
        pps, sc = O.mutual_run(pr["feats0"], pr["feats1"])
        dr = O.rindex(pr["feats0"], pr["feats1"], pps, tables.perm)
        order[i] = rng.permutation(pps.shape[0])[:H]
        stages.append((pps, sc, dr))
    eng = pipeline.YohooEngine(ctx, sd, npass=3, max_iter=H, ird=0.1, nn_mode=0)
    out = eng.register(desc, fcgf, keys, pc, order=ctx.dev(order))
    torch.cuda.synchronize()
    for i, (pr, (pps, sc, dr)) in enumerate(zip(prs, stages)):
        sel = order[i]
        q = O.et_forward(pr["fcgf1"][pps[sel, 1]], pr["fcgf0"][pps[sel, 0]], pr["feats1"][pps[sel, 1]], pr["feats0"][pps[sel, 0]], dr[sel], sd,
                         tables.nei, tables.perm)
        k0 = pr["keys0"][pps[:, 0]]; k1 = pr["keys1"][pps[:, 1]]
        tr = O.hypotheses_from_quat(q, dr[sel], k0[sel], k1[sel], tables.rot)
        assert np.abs(_np(out["hyps"][i]) - tr).max() < 1e-4
        best, bov, ovs = O.oneshot_ransac(k0, k1, sc, tr, 0.1)
        top = np.sort(ovs)[-2:]
        if top[1] > top[0]:                                  # unique winner: index and pose must agree
            assert int(out["recall"][i]) == best
            T = O.refine(k0, k1, tr[best], sc, 0.1)
            assert np.abs(_np(out["poses"][i]) - T).max() < 1e-6
        assert np.abs(_np(out["poses"][i])[:3] - pr["gt"]).max() < 2e-2
    # default device shuffle: poses still agree with ground truth
    out2 = pipeline.YohooEngine(ctx, sd, npass=1, max_iter=200, ird=0.1).register(desc, fcgf, keys, pc, seed=3)
    torch.cuda.synchronize()
    for i, pr in enumerate(prs):
        assert np.abs(_np(out2["poses"][i])[:3] - pr["gt"]).max() < 2e-2
