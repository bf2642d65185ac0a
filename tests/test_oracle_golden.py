"""Pin the NumPy oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py running the reference's own mutual / yohoo / yohoc / yoho_des / yoho_det
classes on CPU).  Integer outputs must be bit-exact; float outputs to the stated tolerance."""
import numpy as np
import pytest
from conftest import load_golden
from oracle import roreg_oracle as O
from roreg_b200 import synth


def _replay_mutual(ds, keynum, seed):
    np.random.seed(seed)
    res = []
    for (id0, id1) in ds.pair_ids:
        f0 = ds.get_feats(id0); f1 = ds.get_feats(id1)
        s0 = np.arange(f0.shape[0]); s1 = np.arange(f1.shape[0])
        np.random.shuffle(s0); np.random.shuffle(s1)                  # test/matcher.py:85-88
        res.append(O.mutual_run(f0, f1, s0[:keynum], s1[:keynum]))
    return res


@pytest.mark.parametrize("name", ["s256", "s700"])
def test_pipeline_against_reference_outputs(name, tables):
    z, n, keynum, max_iter, seeds = load_golden(name)
    ds = synth.SynthDataset(seeds, n=n, name=f"synth/{name}", max_res_deg=2.0)
    matches = _replay_mutual(ds, keynum, 1234)
    sd_et = O.random_state_dict("ET", 102)
    np.random.seed(4321)
    trans_all = []
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        pps, sc = matches[pi]
        assert np.array_equal(pps, z[f"match_{id0}-{id1}"])
        assert np.array_equal(sc, z[f"scores_{id0}-{id1}"])
        f0 = ds.get_feats(id0); f1 = ds.get_feats(id1)
        dr = O.rindex(f0, f1, pps, tables.perm)
        assert np.array_equal(dr, z[f"dr_index_{id0}-{id1}"])
        q = O.et_forward(ds.get_feats(id1, "fcgf")[pps[:, 1]], ds.get_feats(id0, "fcgf")[pps[:, 0]],
                         f1[pps[:, 1]], f0[pps[:, 0]], dr, sd_et, tables.nei, tables.perm)
        k0 = ds.get_kps(id0)[pps[:, 0]]; k1 = ds.get_kps(id1)[pps[:, 1]]
        tr = O.hypotheses_from_quat(q, dr, k0, k1, tables.rot)
        # float32 network on two BLAS paths: 1e-5 on the rotation part is the stated tolerance
        assert np.abs(tr - z[f"trans_pre_{id0}-{id1}"]).max() < 2e-5
        trans_all.append((k0, k1, sc, z[f"trans_pre_{id0}-{id1}"], dr))
    for pi, (id0, id1) in enumerate(ds.pair_ids):              # yohoo_ransac.ransac runs after Rt_pre for all pairs
        k0, k1, sc, tr, dr = trans_all[pi]
        T, best, info = O.yohoo_ransac(k0, k1, sc, tr, 0.1, max_iter)
        assert best == int(z[f"yohoo_recall_{id0}-{id1}"])
        assert np.abs(T - z[f"yohoo_trans_{id0}-{id1}"]).max() < 1e-9
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        k0, k1, sc, tr, dr = trans_all[pi]
        np.random.seed(777 + pi)
        T, recall, info = O.yohoc_ransac(k0, k1, sc, dr, 0.1, max_iter)
        assert recall == int(z[f"yohoc_recall_{id0}-{id1}"])
        assert np.abs(T - z[f"yohoc_trans_{id0}-{id1}"]).max() < 1e-9


def test_per_cloud_nets_against_reference_outputs(tables):
    z, n, keynum, max_iter, seeds = load_golden("s256")
    ds = synth.SynthDataset(seeds[:1], n=n, name="synth/s256", max_res_deg=2.0)
    sd_gf = O.random_state_dict("GF", 101); sd_rd = O.random_state_dict("RD", 103)
    for cid in ds.pc_ids:
        x = ds.get_feats(cid, "fcgf")
        eqv, _ = O.gf_forward(x[:40], sd_gf, tables.nei)
        assert np.abs(eqv - z[f"gf_eqv_{cid}"]).max() < 5e-6
    # detector: rank-normalised scores are a permutation statistic -> compare ranks on the full cloud
    cid = ds.pc_ids[0]
    eqv, _ = O.gf_forward(ds.get_feats(cid, "fcgf"), sd_gf, tables.nei)
    s = O.rank_normalise(O.rd_forward(eqv, sd_rd, tables.nei, tables.perm))
    ref = z[f"det_score_{cid}"]
    assert np.mean(np.abs(s - ref) * n <= 1.0) > 0.98      # ranks agree up to float32 near-ties


def test_kat_des2r(tables):
    """Des2R(X, X[:,:,P[a]]) = a and Des2R(X[:,:,P[a]], X) = inv[a]."""
    rng = np.random.default_rng(0)
    X = rng.standard_normal((5, 32, 60)).astype(np.float32)
    for a in (0, 1, 7, 33, 59):
        assert (O.des2r(X, X[:, :, tables.perm[a]], tables.perm) == a).all()
        assert (O.des2r(X[:, :, tables.perm[a]], X, tables.perm) == tables.inv[a]).all()


def test_kat_planted_pose(tables):
    """a planted SE(3) among random hypotheses is returned at its position and
    refining exact correspondences returns the planted pose."""
    pr = synth.make_pair(3, n=300, sigma_xyz=0.0)
    m = pr["corr0"] >= 0
    k0 = pr["keys0"][m]; k1 = pr["keys1"][pr["corr0"][m]]
    rng = np.random.default_rng(1)
    H = np.concatenate([np.linalg.qr(rng.standard_normal((50, 3, 3)))[0], rng.standard_normal((50, 3, 1))], 2)
    H[17] = pr["gt"]
    best, ov, _ = O.oneshot_ransac(k0, k1, np.ones(k0.shape[0]), H, 0.1)
    assert best == 17 and ov == 1.0
    assert np.abs(O.refine(k0, k1, H[17], np.ones(k0.shape[0]), 0.1)[:3] - pr["gt"]).max() < 1e-9


def _replay_rm(ds, keynum, seed, fn):
    """yoho_mat.run's host logic (test/matcher.py:166-210) around a Match_ot forward `fn(feats_src, feats_tgt, keys_src, keys_tgt)`."""
    np.random.seed(seed)
    res = []
    for (id0, id1) in ds.pair_ids:
        f0 = ds.get_feats(id0); f1 = ds.get_feats(id1)
        s0 = np.arange(f0.shape[0]); s1 = np.arange(f1.shape[0])
        np.random.shuffle(s0); np.random.shuffle(s1)
        s0 = s0[:keynum]; s1 = s1[:keynum]
        m0, sc0 = fn(f1[s1], f0[s0], ds.get_kps(id1)[s1].astype(np.float32), ds.get_kps(id0)[s0].astype(np.float32))
        sel = np.where(m0 != -1)[0]
        pairs = np.stack([sel, m0[sel]], 1)
        res.append((np.stack([s0[pairs[:, 1]], s1[pairs[:, 0]]], 1), sc0[sel]))
    return res


def test_match_ot_against_reference_outputs(tables):
    """Match_ot restatement vs the files the reference's yoho_mat.run wrote (tests/golden/make_golden_rm.py)."""
    z, n, keynum, _, seeds = load_golden("rm300")
    ds = synth.SynthDataset(seeds, n=n, name="synth/rm", with_fcgf=False)
    sd = O.random_state_dict("RM", 104)
    out = _replay_rm(ds, keynum, 2468, lambda a, b, ka, kb: O.match_ot_forward(a, b, ka, kb, sd, tables.perm)[:2])
    for (id0, id1), (m, s) in zip(ds.pair_ids, out):
        assert np.array_equal(m, z[f"match_{id0}-{id1}"])
        assert np.abs(s - z[f"scores_{id0}-{id1}"]).max() < 1e-4          # stated tolerance; matches are bit-exact


def test_torch_mirror_equals_numpy_oracle():
    """oracle/torch_mirror.py (the reference's tensor operations on CPU, used only for bench.py's `reference_ops` timing) returns
    what the NumPy restatement returns: same matches, same coarse-rotation indices, a pose at the planted ground truth."""
    from oracle import torch_mirror as TM
    from roreg_b200 import group, synth
    t = group.load()
    pr = synth.make_pair(5, n=500)
    T, pps, dr = TM.register_pair(pr, t.perm, 200, 0.1, 0)
    p2, _ = O.mutual_run(pr["feats0"], pr["feats1"])
    assert np.array_equal(pps, p2)
    assert np.array_equal(dr, O.rindex(pr["feats0"], pr["feats1"], p2, t.perm))
    assert np.abs(T[:3] - pr["gt"]).max() < 1e-2


def test_oracle_nms_sampler_and_yohoc_helpers_equal_reference():
    """a14 NMS_sample (test/matcher.py:11-42) and the host helpers of a20 (DR_statictic, Threepps2Tran, test/estimator.py:119-147)
    against tests/golden/s400rdrm.npz (tests/golden/make_golden_rd_rm.py: the unmodified reference)."""
    from conftest import load_golden
    from roreg_b200 import synth
    z, n, keynum, max_iter, seeds = load_golden("s400rdrm")
    ds = synth.SynthDataset(seeds, n=n, name="synth/s400rdrm", max_res_deg=2.0)
    keys = ds.get_kps(ds.pc_ids[0]); sc0 = z[f"det_score_{ds.pc_ids[0]}"]
    for num in (40, 300, 380, 500):
        assert np.array_equal(O.nms_sample(keys, sc0, num), z[f"nms_{num}"])
    for num in (40, 300):
        assert np.array_equal(O.nms_sample(keys, z["nms_flat_scores"], num), z[f"nms_flat_{num}"])
    dr = z[f"dr_index_{ds.pair_ids[0][0]}-{ds.pair_ids[0][1]}"]
    stat, prob = O.dr_statistic(dr)
    assert np.array_equal(prob, z["drstat_prob"])
    assert np.array_equal(np.concatenate([np.asarray(stat[i], np.int64) for i in range(60)]), z["drstat_members"])
    k0 = ds.get_kps(ds.pair_ids[0][0]); k1 = ds.get_kps(ds.pair_ids[0][1])
    for t, T in zip(z["kabsch_triplets"], z["kabsch_T"]):
        assert np.array_equal(O.threepps2tran(k0[t], k1[t]), T)
    # --RM selection + one-shot RANSAC + refine of the reference (test/estimator.py:404-441) through the oracle
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        m = z[f"match_{id0}-{id1}"]; s = z[f"scores_{id0}-{id1}"]
        k0m = ds.get_kps(id0)[m[:, 0]]; k1m = ds.get_kps(id1)[m[:, 1]]
        if pi == 0:
            rng = np.random.RandomState(1357)
        T, best, _ = O.yohoo_ransac(k0m, k1m, s, z[f"trans_pre_{id0}-{id1}"], 0.1, max_iter, RM=True, match_n=0.5, rng=rng)
        assert best == int(z[f"yohoo_recall_{id0}-{id1}"]) and np.abs(T - z[f"yohoo_trans_{id0}-{id1}"]).max() < 1e-10
        T, recall, _ = O.yohoc_ransac(k0m, k1m, s, z[f"dr_index_{id0}-{id1}"], 0.1, max_iter, RM=True, match_n=0.5,
                                      rng=np.random.RandomState(555 + pi))
        assert recall == int(z[f"yohoc_recall_{id0}-{id1}"]) and np.abs(T - z[f"yohoc_trans_{id0}-{id1}"]).max() < 1e-10


def test_kat_equivariance_and_invariance_of_the_oracle_networks(tables):
    """permuting the input's group axis by P[a] permutes the GF output the same way (and leaves the
    detector's saliency and the invariant pooling unchanged)."""
    rng = np.random.default_rng(5)
    sd_gf = O.random_state_dict("GF", 101); sd_rd = O.random_state_dict("RD", 103)
    x = rng.standard_normal((6, 32, 60)).astype(np.float32)
    y, _ = O.gf_forward(x, sd_gf, tables.nei)
    s = O.rd_forward(y, sd_rd, tables.nei, tables.perm)
    for a in (1, 7, 33, 59):
        ya, _ = O.gf_forward(np.ascontiguousarray(x[:, :, tables.perm[a]]), sd_gf, tables.nei)
        assert np.abs(ya - y[:, :, tables.perm[a]]).max() < 2e-6
        assert np.abs(O.inv_pool(ya) - O.inv_pool(y)).max() < 1e-6
        sa = O.rd_forward(np.ascontiguousarray(y[:, :, tables.perm[a]]), sd_rd, tables.nei, tables.perm)
        assert np.abs(sa - s).max() < 1e-5 * max(1.0, np.abs(s).max())
