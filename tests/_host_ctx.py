"""TEST INFRASTRUCTURE: an oracle-backed stand-in for roreg_b200.ops.Context (same method names, argument order and return
layout, CPU torch tensors) so that the HOST logic of the plugin mirrors in roreg_b200/test - sampling, global-RNG order, file
layout, hypothesis bookkeeping, pre.log - can run in the `-m "not gpu"` suite against the reference-generated fixtures.  The
product never imports this; on a GPU the same plugin code runs against the real context (tests/test_gpu_dropin.py)."""
import numpy as np
import torch
from oracle import roreg_oracle as O
from roreg_b200 import group


def _np(t):
    return None if t is None else (t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t))


def _chk(t, dtype, name, optional=False):
    """What the C ABI assumes about every buffer it is handed (ops.Context._chk checks some of these, the kernels assume the rest):
    a contiguous torch tensor of exactly this dtype."""
    if t is None:
        assert optional, f"{name}: required"
        return
    dts = dtype if isinstance(dtype, tuple) else (dtype,)
    assert torch.is_tensor(t) and t.dtype in dts and t.is_contiguous(), f"{name}: need a contiguous {dts} tensor, got {getattr(t, 'dtype', type(t))}"


class HostContext:
    def __init__(self, so3_dir=None):
        self.tables = group.load(so3_dir)
        self.device = torch.device("cpu")
        self.corr_mode = 0

    def set_corr_mode(self, mode):
        self.corr_mode = int(mode)

    def dev(self, a, dtype=None):
        t = torch.as_tensor(np.ascontiguousarray(a)) if not torch.is_tensor(a) else a
        return (t.to(dtype) if dtype is not None else t).contiguous()

    def inv_pool(self, eqv, sample=None, normalise=True):
        _chk(eqv, torch.float32, "eqv"); _chk(sample, torch.int32, "sample", True)
        f = _np(eqv)
        if sample is not None:
            f = f[_np(sample).astype(np.int64)]
        return torch.from_numpy(O.inv_pool(f, normalise))

    def knn(self, target, source, k=1):
        _chk(target, torch.float32, "target"); _chk(source, torch.float32, "source")
        d, i = O.knn(_np(target), _np(source), k)
        return torch.from_numpy(d), torch.from_numpy(i.astype(np.int32))

    def mutual_match(self, f0, f1, mode=0):
        _chk(f0, torch.float32, "f0"); _chk(f1, torch.float32, "f1")
        pps, nn01, nn10 = O.mutual_matches(_np(f0), _np(f1))
        out = np.zeros((min(f0.shape[0], f1.shape[0]), 2), np.int32); out[:pps.shape[0]] = pps
        return (torch.from_numpy(out), torch.tensor([pps.shape[0]], dtype=torch.int32), torch.from_numpy(nn01.astype(np.int32)),
                torch.from_numpy(nn10.astype(np.int32)))

    def group_corr(self, X, Y, idxX=None, idxY=None, variant=1, want_cor=True, want_argmax=True):
        _chk(X, torch.float32, "X"); _chk(Y, torch.float32, "Y"); _chk(idxX, torch.int32, "idxX", True); _chk(idxY, torch.int32, "idxY", True)
        x = _np(X); y = _np(Y)
        if idxX is not None:
            x = x[_np(idxX).astype(np.int64)]
        if idxY is not None:
            y = y[_np(idxY).astype(np.int64)]
        assert variant == 1
        cor = O.group_corr_v1(x, y, self.tables.perm)
        return (torch.from_numpy(cor.astype(np.float32)) if want_cor else None,
                torch.from_numpy(np.argmax(cor, 1).astype(np.int32)) if want_argmax else None)

    def hypotheses_from_quat(self, quat, pre_idx, k0m, k1m):
        _chk(quat, torch.float32, "quat"); _chk(pre_idx, torch.int32, "pre_idx"); _chk(k0m, torch.float64, "k0m"); _chk(k1m, torch.float64, "k1m")
        return torch.from_numpy(O.hypotheses_from_quat(_np(quat), _np(pre_idx), _np(k0m), _np(k1m), self.tables.rot))

    @staticmethod
    def _scores(scores, K):
        return np.ones(K) if scores is None else _np(scores)

    def ransac_oneshot(self, k0m, k1m, scores, trans, order, ird, want_overlaps=False):
        _chk(k0m, torch.float64, "k0m"); _chk(k1m, torch.float64, "k1m"); _chk(scores, (torch.float32, torch.float64), "scores", True)
        _chk(trans, torch.float64, "trans"); _chk(order, torch.int32, "order", True)
        T = _np(trans)
        if order is not None:
            T = T[_np(order).astype(np.int64)]
        k0 = _np(k0m); k1 = _np(k1m)
        best, bov, ovs = O.oneshot_ransac(k0, k1, self._scores(scores, k0.shape[0]), T, ird)
        return (torch.tensor([best], dtype=torch.int32), torch.tensor([float(bov)], dtype=torch.float64),
                torch.from_numpy(ovs) if want_overlaps else None)

    def _pick(self, T_in, order, T_index):
        T = _np(T_in)
        if T_index is None:
            return T
        j = int(_np(T_index).reshape(-1)[0])
        if order is not None:
            j = int(_np(order)[j])
        return T[j]

    def refine(self, k0m, k1m, scores, T_in, ird, order=None, T_index=None, want_mask=False):
        _chk(k0m, torch.float64, "k0m"); _chk(k1m, torch.float64, "k1m"); _chk(scores, (torch.float32, torch.float64), "scores", True)
        _chk(T_in, torch.float64, "T_in"); _chk(order, torch.int32, "order", True); _chk(T_index, torch.int32, "T_index", True)
        k0 = _np(k0m); k1 = _np(k1m); s = self._scores(scores, k0.shape[0])
        T0 = self._pick(T_in, order, T_index)
        T1 = O.refine_once(k0, k1, T0, s, 2.0 * ird)
        T2 = O.refine_once(k0, k1, T1, s, ird)
        mask = torch.from_numpy(O.inlier_mask(k0, k1, T1, ird).astype(np.uint8)) if want_mask else None
        return torch.from_numpy(T2), mask

    def refine_once(self, k0m, k1m, scores, T_in, radius, want_mask=False):
        _chk(k0m, torch.float64, "k0m"); _chk(k1m, torch.float64, "k1m"); _chk(scores, (torch.float32, torch.float64), "scores", True)
        _chk(T_in, torch.float64, "T_in")
        k0 = _np(k0m); k1 = _np(k1m); s = self._scores(scores, k0.shape[0])
        T0 = _np(T_in)
        mask = torch.from_numpy(O.inlier_mask(k0, k1, T0, radius).astype(np.uint8)) if want_mask else None
        return torch.from_numpy(O.refine_once(k0, k1, T0, s, radius)), mask

    def register_batch(self, desc, keys, pair_cloud, keynum=None, sample=None, nn_mode=0, estimator=0, max_iter=1000,
                       ird=0.1, seed=0, triplets=None, hyps=None, out=None):
        """Oracle version of the batched engine (estimator 0, draws from a per-pair NumPy RandomState instead of the device's
        counter-based RNG): same output dict layout as ops.Context.register_batch."""
        assert estimator == 0 and triplets is None and hyps is None
        _chk(desc, torch.float32, "desc"); _chk(keys, torch.float64, "keys"); _chk(pair_cloud, torch.int32, "pair_cloud"); _chk(sample, torch.int32, "sample", True)
        D = _np(desc); Kp = _np(keys); pc = _np(pair_cloud); B = pc.shape[0]; n = D.shape[1]; S = keynum or n
        smp = None if sample is None else _np(sample).astype(np.int64)
        o = dict(matches=np.zeros((B, S, 2), np.int32), n_matches=np.zeros(B, np.int32), dr_index=np.zeros((B, S), np.int32),
                 poses=np.zeros((B, 4, 4)), recall=np.full(B, -1, np.int32), best_overlap=np.zeros(B))
        for p in range(B):
            f0, f1 = D[pc[p, 0]], D[pc[p, 1]]
            pps, sc = O.mutual_run(f0, f1, None if smp is None else smp[p, 0], None if smp is None else smp[p, 1])
            k = pps.shape[0]
            o["n_matches"][p] = k
            if k == 0:
                continue
            dr = O.rindex(f0, f1, pps, self.tables.perm)
            o["matches"][p, :k] = pps; o["dr_index"][p, :k] = dr
            T, rec, info = O.yohoc_ransac(Kp[pc[p, 0]][pps[:, 0]], Kp[pc[p, 1]][pps[:, 1]], sc, dr, ird, max_iter,
                                          rng=np.random.RandomState((int(seed) + p) % (2 ** 31)))
            if T is not None and rec > 0:
                o["poses"][p] = T; o["recall"][p] = rec - 1; o["best_overlap"][p] = info["best_overlap"]
        return {k: torch.from_numpy(v) for k, v in o.items()}

    def kabsch3(self, k0s, k1s, triplets):
        _chk(k0s, torch.float64, "k0s"); _chk(k1s, torch.float64, "k1s"); _chk(triplets, torch.int32, "triplets")
        k0 = _np(k0s); k1 = _np(k1s)
        return torch.from_numpy(np.stack([O.threepps2tran(k0[t], k1[t]) for t in _np(triplets).astype(np.int64)]))


def install(monkeypatch):
    """Route `context(cfg)` of every plugin module to one HostContext."""
    import roreg_b200.test._common as common
    import roreg_b200.test.matcher as matcher
    import roreg_b200.test.estimator as estimator
    import roreg_b200.test.extractor as extractor
    import roreg_b200.test.detector as detector
    ctx = HostContext()

    def fake_context(cfg=None):
        cm = getattr(cfg, "corr_mode", None) if cfg is not None else None
        if cm is not None:
            ctx.set_corr_mode(cm)
        return ctx
    for mod in (common, matcher, estimator, extractor, detector):
        monkeypatch.setattr(mod, "context", fake_context, raising=True)
    return ctx


class HostGFNet:
    """Stand-in for roreg_b200.nets.GFNet (same constructor / forward signature), oracle arithmetic."""

    def __init__(self, ctx, sd, npass=3, chunk=500):
        self.ctx, self.sd = ctx, sd

    def forward(self, x):
        return torch.from_numpy(O.gf_forward(_np(x), self.sd, self.ctx.tables.nei)[0])


class HostRDNet:
    """Stand-in for roreg_b200.nets.RDNet."""

    def __init__(self, ctx, sd, npass=3):
        self.ctx, self.sd = ctx, sd

    def forward(self, x):
        return torch.from_numpy(O.rd_forward(_np(x), self.sd, self.ctx.tables.nei, self.ctx.tables.perm))


class HostETNet:
    """Stand-in for roreg_b200.nets.ETNet: forward(before0, rows, before1, rows, after0, rows, after1, rows, pre_idx) -> [K,4]."""

    def __init__(self, ctx, sd, npass=3, chunk=1000):
        self.ctx, self.sd = ctx, sd

    def forward(self, before0, rows_b0, before1, rows_b1, after0, rows_a0, after1, rows_a1, pre_idx):
        r = lambda t, i: _np(t)[_np(i).astype(np.int64)]
        q = O.et_forward(r(before0, rows_b0), r(before1, rows_b1), r(after0, rows_a0), r(after1, rows_a1),
                         _np(pre_idx).astype(np.int64), self.sd, self.ctx.tables.nei, self.ctx.tables.perm)
        return torch.from_numpy(q)


class HostMatchOT:
    """Stand-in for roreg_b200.matchot.MatchOT: forward(src_eqv, tgt_eqv, keys_src, keys_tgt) -> (matches0, scores0)."""

    def __init__(self, ctx, sd, npass=3, sinkhorn_iters=100):
        self.ctx, self.sd, self.iters = ctx, sd, sinkhorn_iters

    def forward(self, src_eqv, tgt_eqv, keys_src, keys_tgt):
        out = O.match_ot_forward(_np(src_eqv), _np(tgt_eqv), _np(keys_src), _np(keys_tgt), self.sd, self.ctx.tables.perm, self.iters)
        return torch.from_numpy(np.asarray(out[0]).astype(np.int32)), torch.from_numpy(np.asarray(out[1]).astype(np.float32))


def install_nets(monkeypatch):
    import roreg_b200.nets as nets_mod
    import roreg_b200.matchot as matchot_mod
    monkeypatch.setattr(matchot_mod, "MatchOT", HostMatchOT, raising=True)
    monkeypatch.setattr(nets_mod, "ETNet", HostETNet, raising=True)
    import roreg_b200.test.extractor as extractor
    import roreg_b200.test.detector as detector
    monkeypatch.setattr(extractor.nets, "GFNet", HostGFNet, raising=True)
    monkeypatch.setattr(detector.nets, "RDNet", HostRDNet, raising=True)
