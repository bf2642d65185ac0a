"""Rotation-coherence matcher on the GPU: building blocks against NumPy, MatchOT.forward against the oracle, and the
yoho_mat plugin against the files the UNMODIFIED reference wrote (tests/golden/rm300.npz)."""
import os
import types
import numpy as np
import pytest
import torch
from conftest import load_golden
from oracle import roreg_oracle as O
from roreg_b200 import synth
from test_oracle_golden import _replay_rm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from roreg_b200 import ops
    c = ops.Context(0)
    yield c
    c.close()


def _np(t):
    return t.cpu().numpy()


def test_topk_gather_stats(ctx):
    from roreg_b200 import matchot
    mo = matchot.MatchOT(ctx, O.random_state_dict("RM", 1))
    rng = np.random.default_rng(0)
    S = rng.standard_normal((77, 1300)).astype(np.float32)
    S[5, 10] = S[5, 900] = 9.0                                   # tie: lower column first
    idx = _np(mo.topk(ctx.dev(S), 77, 1300, 16))
    ref = np.argsort(-S, axis=1, kind="stable")[:, :16]
    assert np.array_equal(idx, ref)
    src = rng.standard_normal((1300, 32)).astype(np.float32)
    g = _np(mo.gather(ctx.dev(src), ctx.dev(ref.astype(np.int32)), 32))
    assert np.array_equal(g, src[ref.reshape(-1)])
    x = rng.standard_normal((5000, 64)).astype(np.float32) * 3 + 1
    mean, rstd = mo.stats(ctx.dev(x), 5000, 64)
    assert np.abs(_np(mean) - x.mean(0)).max() < 1e-5 and np.abs(_np(rstd) - 1 / np.sqrt(x.var(0) + 1e-5)).max() < 1e-5


@pytest.mark.parametrize("m,n", [(211, 190), (204, 192), (37, 4), (1500, 1208)],
                         ids=["launch-per-pass (n % 4 != 0)", "persistent", "persistent tiny", "persistent, rows split over all CTAs"])
def test_sinkhorn_against_oracle(ctx, m, n):
    """roreg_sinkhorn_match: the persistent cooperative kernel (n, ld multiples of 4) and the launch-per-pass kernels it falls
    back to, against the oracle's float32 log-domain Sinkhorn: OT matrix to 2e-4, mutual assignment exact."""
    from roreg_b200 import matchot, _lib
    from roreg_b200.ops import _ptr, _stream
    import ctypes as C
    rng = np.random.default_rng(1)
    S = (rng.standard_normal((m, n)) * 3).astype(np.float32)
    u = torch.empty(m + 1, dtype=torch.float32, device=ctx.device); v = torch.empty(n + 1, dtype=torch.float32, device=ctx.device)
    m0 = torch.empty(m, dtype=torch.int32, device=ctx.device); s0 = torch.empty(m, dtype=torch.float32, device=ctx.device)
    rc = ctx.lib.roreg_sinkhorn_match(ctx.h, _ptr(ctx.dev(S)), m, n, n, C.c_float(0.7), 100, _ptr(u), _ptr(v), _ptr(m0), _ptr(s0), _stream())
    _lib.check(ctx.h, rc, "sinkhorn")
    Z = O.log_sinkhorn(S, np.float32(0.7), 100)
    norm = -np.log(np.float32(m + n))
    got = S + _np(u)[:m, None] + _np(v)[None, :n] - norm
    assert np.abs(got - Z[:m, :n]).max() < 2e-4
    inner = Z[:-1, :-1]; i0 = inner.argmax(1); i1 = inner.argmax(0)
    mut = np.arange(m) == i1[i0]
    assert np.array_equal(_np(m0), np.where(mut, i0, -1))
    assert np.abs(_np(s0) - np.where(mut, np.exp(inner.max(1)), 0)).max() < 1e-4


KNN_BAND = 2e-4      # relative to the row's largest |score|: float32 (3xTF32 GEMM) evaluation error of a score_mat entry
OT_BAND = 2e-3       # absolute, log domain: error of a Sinkhorn-normalised score (100 iterations of float32 logsumexp)


def _adjudicate_forward(ctx, tables, seed, n, capsys):
    """Teacher-forced parity of MatchOT.forward (rot_coh_match.py:339-390).  The matcher is a chain of DISCRETE decisions
    (8 top-k neighbour lists, then the mutual argmax on the OT matrix) joined by continuous float32 layers, so:
      1. every neighbour list the CUDA path chose must be a valid top-k of the ORACLE's score matrix for that block (the
         oracle runs with the CUDA lists forced, so both see the same inputs up to rounding): chosen scores in descending
         order and no unchosen column better than the k-th, both within KNN_BAND; rows where the oracle's own stable argsort
         would have picked differently are counted (near ties), not tolerated silently;
      2. the final score matrix agrees to 1e-3 of its scale;
      3. matches0: equal to the oracle's, except rows whose OT row / column top-2 gap is inside OT_BAND."""
    from roreg_b200 import matchot
    pr = synth.make_pair(seed, n=n)
    sd = O.random_state_dict("RM", 104)
    mo = matchot.MatchOT(ctx, sd, npass=3)
    mo.trace = {}
    m0, s0 = mo.forward(ctx.dev(pr["feats1"]), ctx.dev(pr["feats0"]), ctx.dev(pr["keys1"].astype(np.float32)), ctx.dev(pr["keys0"].astype(np.float32)))
    torch.cuda.synchronize()
    forced = {k: _np(v).astype(np.int64) for k, v in mo.trace.items() if k != "final_score"}
    assert len(forced) == 8
    tr = {}
    r0, rs0, _, _, Z = O.match_ot_forward(pr["feats1"], pr["feats0"], pr["keys1"], pr["keys0"], sd, tables.perm, forced=forced, trace=tr)
    knn_ties = 0
    for p, knn in forced.items():
        Sc = tr[p]["score"]; own = tr[p]["knn"]
        band = KNN_BAND * np.abs(Sc).max(axis=1)
        chosen = np.take_along_axis(Sc, knn, 1)
        assert np.all(chosen[:, :-1] - chosen[:, 1:] >= -band[:, None]), p                 # descending order
        rest = Sc.copy(); np.put_along_axis(rest, knn, -np.inf, 1)
        assert np.all(chosen[:, -1] >= rest.max(axis=1) - band), p                          # nothing better left out
        assert all(len(set(r)) == len(r) for r in knn.tolist()), p
        knn_ties += int((knn != own).any(axis=1).sum())
    S = _np(mo.trace["final_score"]); Sr = tr["final_score"]
    assert np.abs(S - Sr).max() < 1e-3 * max(1.0, np.abs(Sr).max())
    got = _np(m0); inner = Z[:-1, :-1]
    bad = np.flatnonzero(got != r0)
    if bad.size:
        row2 = np.sort(inner, axis=1)[:, -2:]; col2 = np.sort(inner, axis=0)[-2:, :]
        row_tie = (row2[:, 1] - row2[:, 0]) < OT_BAND; col_tie = (col2[1] - col2[0]) < OT_BAND
        i0 = inner.argmax(1)
        for i in bad:
            cols = {int(i0[i])} | ({int(got[i])} if got[i] >= 0 else set())
            assert row_tie[i] or any(col_tie[j] for j in cols), (int(i), int(got[i]), int(r0[i]))
    both = (got == r0) & (r0 >= 0)
    assert np.abs(_np(s0)[both] - rs0[both]).max() < 1e-3 * max(1.0, rs0.max())
    with capsys.disabled():
        print(f"\n[Match_ot parity n={n}] rows with a near-tied neighbour list (of {8 * n}): {knn_ties}; "
              f"assignments inside the OT near-tie band: {bad.size} of {n}; matched {int((r0 >= 0).sum())}")


def test_match_ot_forward_against_oracle(ctx, tables, capsys):
    _adjudicate_forward(ctx, tables, 58, 400, capsys)


def test_match_ot_forward_against_oracle_larger(ctx, tables, capsys):
    _adjudicate_forward(ctx, tables, 59, 1100, capsys)


def test_yoho_mat_plugin_reproduces_reference_files(tmp_path, capsys):
    import roreg_b200.test as rt
    z, n, keynum, _, seeds = load_golden("rm300")
    ds = synth.SynthDataset(seeds, n=n, name="synth/rm", with_fcgf=False)
    cache = str(tmp_path / "cache"); ds.write_cache(cache)
    model_fn = str(tmp_path / "ckpt"); os.makedirs(f"{model_fn}/RM")
    torch.save({"best_para": 0, "network_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in O.random_state_dict("RM", 104).items()}},
               f"{model_fn}/RM/model_best.pth")
    cfg = types.SimpleNamespace(output_cache_fn=cache, model_fn=model_fn, SO3_related_files=None, backbone="FCGF", bs_GF=1250, bs_ET=1000,
                                RD=False, RM=True, match_n=0.5, ransac_ird=0.1)
    np.random.seed(2468)
    rt.name2matcher["yoho_mat"](cfg).run(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        m = np.load(f"{cache}/synth/rm/match_{keynum}/{id0}-{id1}.npy"); s = np.load(f"{cache}/synth/rm/match_{keynum}/scores/{id0}-{id1}.npy")
        ref = z[f"match_{id0}-{id1}"]
        a = {tuple(r) for r in m.tolist()}; b = {tuple(r) for r in ref.tolist()}
        with capsys.disabled():
            print(f"\n[yoho_mat vs reference file {id0}-{id1}] {len(b)} reference matches, {len(a ^ b)} in the symmetric difference")
        # The reference file comes from torch on the CPU, the plugin from the 3xTF32 GEMMs: a neighbour list decided inside the
        # float32 near-tie band (test_match_ot_forward_against_oracle counts them: 0-2 rows of 8 n) changes that row's features and
        # with them a handful of assignments.  The teacher-forced test above is the exact one; here the bound is a few matches.
        assert len(a ^ b) <= max(2, len(b) // 50), (len(a), len(b), len(a ^ b))
        common = sorted(a & b)
        if common:
            ia = {tuple(r): i for i, r in enumerate(m.tolist())}; ib = {tuple(r): i for i, r in enumerate(ref.tolist())}
            sa = np.array([s[ia[r]] for r in common]); sb = np.array([z[f"scores_{id0}-{id1}"][ib[r]] for r in common])
            if len(a ^ b) == 0:
                assert np.abs(sa - sb).max() < 1e-3
