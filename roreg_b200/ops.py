"""Thin torch-tensor front end over the C ABI.  PyTorch is used only for device memory and streams;
every computation below is a call into libroreg_b200.so.  All tensors must live on the context's
CUDA device and be contiguous; outputs are allocated here and returned."""
import ctypes as C
import numpy as np
import torch

from . import _lib, group as _group


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """One roreg_ctx per (device, host thread).  Holds the icosahedral group tables on the device."""

    def __init__(self, device=0, tables=None, so3_dir=None):
        if not torch.cuda.is_available():
            raise _lib.RoregLibraryError("roreg_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.tables = tables or _group.load(so3_dir)
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        perm = np.ascontiguousarray(self.tables.perm, np.int32)
        nei = np.ascontiguousarray(self.tables.nei, np.int32)
        rot = np.ascontiguousarray(self.tables.rot, np.float64)
        rc = self.lib.roreg_ctx_create(device, perm.ctypes.data, nei.ctypes.data, rot.ctypes.data, C.byref(h))
        self.h = h
        self._pipe_keep = []
        _lib.check(self.h, rc, "roreg_ctx_create")

    def close(self):
        if getattr(self, "h", None):
            self.lib.roreg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.roreg_launch_count(self.h))

    STAGES = ("inv_pool", "nn", "compact", "group_corr", "hypotheses", "score_select", "refine")

    def set_overlap(self, on):
        """Two-stream schedule of register_batch (default off: it measured slower on a B200); results are identical either way."""
        _lib.check(self.h, self.lib.roreg_set_overlap(self.h, int(on)), "roreg_set_overlap")

    def set_score_mode(self, mode):
        """0: float64 one-shot scoring (default); 1: float32 pre-filter with exact float64 re-check (same results)."""
        _lib.check(self.h, self.lib.roreg_set_score_mode(self.h, int(mode)), "roreg_set_score_mode")

    def set_corr_mode(self, mode):
        _lib.check(self.h, self.lib.roreg_set_corr_mode(self.h, int(mode)), "roreg_set_corr_mode")

    def set_timing(self, on=True):
        _lib.check(self.h, self.lib.roreg_set_timing(self.h, int(on)), "roreg_set_timing")

    def stage_ms(self):
        buf = (C.c_float * len(self.STAGES))()
        _lib.check(self.h, self.lib.roreg_get_stage_ms(self.h, buf), "roreg_get_stage_ms")
        return dict(zip(self.STAGES, [float(x) for x in buf]))

    # ---- helpers -------------------------------------------------------------------------------
    def _chk(self, t, dtype, name):
        assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), f"{name}: need contiguous {dtype} CUDA tensor"
        return t

    def dev(self, a, dtype=None):
        t = torch.as_tensor(np.ascontiguousarray(a)) if not torch.is_tensor(a) else a
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device).contiguous()

    # ---- a13 -----------------------------------------------------------------------------------
    def inv_pool(self, eqv, sample=None, normalise=True):
        self._chk(eqv, torch.float32, "eqv")
        n_out = eqv.shape[0] if sample is None else sample.shape[0]
        if sample is not None:
            self._chk(sample, torch.int32, "sample")
        out = torch.empty((n_out, 32), dtype=torch.float32, device=self.device)
        rc = self.lib.roreg_inv_pool(self.h, _ptr(eqv), _ptr(sample), n_out, int(normalise), _ptr(out), _stream())
        _lib.check(self.h, rc, "roreg_inv_pool")
        return out

    # ---- a15 -----------------------------------------------------------------------------------
    def knn(self, target, source, k=1):
        self._chk(target, torch.float32, "target"); self._chk(source, torch.float32, "source")
        n, f = target.shape; m = source.shape[0]
        dist = torch.empty((m, k), dtype=torch.float32, device=self.device)
        idx = torch.empty((m, k), dtype=torch.int32, device=self.device)
        rc = self.lib.roreg_knn(self.h, _ptr(target), n, _ptr(source), m, f, k, _ptr(dist), _ptr(idx), _stream())
        _lib.check(self.h, rc, "roreg_knn")
        return dist, idx

    def mutual_match(self, f0, f1, mode=0):
        self._chk(f0, torch.float32, "f0"); self._chk(f1, torch.float32, "f1")
        n0, n1 = f0.shape[0], f1.shape[0]
        matches = torch.empty((min(n0, n1), 2), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        nn01 = torch.empty(n0, dtype=torch.int32, device=self.device)
        nn10 = torch.empty(n1, dtype=torch.int32, device=self.device)
        rc = self.lib.roreg_mutual_match(self.h, _ptr(f0), n0, _ptr(f1), n1, mode, _ptr(matches), _ptr(cnt),
                                         _ptr(nn01), _ptr(nn10), _stream())
        _lib.check(self.h, rc, "roreg_mutual_match")
        return matches, cnt, nn01, nn10

    # ---- a4 / a5 -------------------------------------------------------------------------------
    def group_corr(self, X, Y, idxX=None, idxY=None, variant=1, want_cor=True, want_argmax=True):
        self._chk(X, torch.float32, "X"); self._chk(Y, torch.float32, "Y")
        K = (idxX.shape[0] if idxX is not None else X.shape[0])
        cor = torch.empty((K, 60), dtype=torch.float32, device=self.device) if want_cor else None
        am = torch.empty(K, dtype=torch.int32, device=self.device) if want_argmax else None
        rc = self.lib.roreg_group_corr(self.h, _ptr(X), _ptr(idxX), _ptr(Y), _ptr(idxY), K, variant, _ptr(cor),
                                       _ptr(am), _stream())
        _lib.check(self.h, rc, "roreg_group_corr")
        return cor, am

    # ---- a17 -----------------------------------------------------------------------------------
    def hypotheses_from_quat(self, quat, pre_idx, k0m, k1m):
        K = quat.shape[0]
        trans = torch.empty((K, 3, 4), dtype=torch.float64, device=self.device)
        rc = self.lib.roreg_hypotheses_from_quat(self.h, _ptr(self._chk(quat, torch.float32, "quat")),
                                                 _ptr(self._chk(pre_idx, torch.int32, "pre_idx")),
                                                 _ptr(self._chk(k0m, torch.float64, "k0m")),
                                                 _ptr(self._chk(k1m, torch.float64, "k1m")), K, _ptr(trans), _stream())
        _lib.check(self.h, rc, "roreg_hypotheses_from_quat")
        return trans

    # ---- a18 / a19 -----------------------------------------------------------------------------
    def ransac_oneshot(self, k0m, k1m, scores, trans, order, ird, want_overlaps=False):
        K = k0m.shape[0]
        H = order.shape[0] if order is not None else trans.shape[0]
        f64 = int(scores is not None and scores.dtype == torch.float64)
        ov = torch.empty(H, dtype=torch.float64, device=self.device) if want_overlaps else None
        best = torch.empty(1, dtype=torch.int32, device=self.device)
        bov = torch.empty(1, dtype=torch.float64, device=self.device)
        rc = self.lib.roreg_ransac_oneshot(self.h, _ptr(k0m), _ptr(k1m), _ptr(scores), f64, K, _ptr(trans), _ptr(order),
                                           H, float(ird), _ptr(ov), _ptr(best), _ptr(bov), _stream())
        _lib.check(self.h, rc, "roreg_ransac_oneshot")
        return best, bov, ov

    def refine(self, k0m, k1m, scores, T_in, ird, order=None, T_index=None, want_mask=False):
        K = k0m.shape[0]
        f64 = int(scores is not None and scores.dtype == torch.float64)
        out = torch.empty((4, 4), dtype=torch.float64, device=self.device)
        mask = torch.empty(K, dtype=torch.uint8, device=self.device) if want_mask else None
        rc = self.lib.roreg_refine(self.h, _ptr(k0m), _ptr(k1m), _ptr(scores), f64, K, _ptr(T_in), _ptr(order),
                                   _ptr(T_index), float(ird), _ptr(out), _ptr(mask), _stream())
        _lib.check(self.h, rc, "roreg_refine")
        return out, mask

    def refine_once(self, k0m, k1m, scores, T_in, radius, want_mask=False):
        K = k0m.shape[0]
        f64 = int(scores is not None and scores.dtype == torch.float64)
        out = torch.empty((4, 4), dtype=torch.float64, device=self.device)
        mask = torch.empty(K, dtype=torch.uint8, device=self.device) if want_mask else None
        rc = self.lib.roreg_refine_once(self.h, _ptr(k0m), _ptr(k1m), _ptr(scores), f64, K, _ptr(T_in), float(radius),
                                        _ptr(out), _ptr(mask), _stream())
        _lib.check(self.h, rc, "roreg_refine_once")
        return out, mask

    def kabsch3(self, k0s, k1s, triplets):
        H = triplets.shape[0]
        trans = torch.empty((H, 3, 4), dtype=torch.float64, device=self.device)
        rc = self.lib.roreg_kabsch3(self.h, _ptr(k0s), _ptr(k1s), _ptr(self._chk(triplets, torch.int32, "triplets")), H,
                                    _ptr(trans), _stream())
        _lib.check(self.h, rc, "roreg_kabsch3")
        return trans

    # ---- batched engine ------------------------------------------------------------------------
    def register_batch(self, desc, keys, pair_cloud, keynum=None, sample=None, nn_mode=0, estimator=0, max_iter=1000,
                       ird=0.1, seed=0, triplets=None, hyps=None, out=None, pipelined=False):
        """desc [n_clouds,n,32,60] f32, keys [n_clouds,n,3] f64, pair_cloud [B,2] int32 (device tensors).
        Returns dict of device tensors: matches [B,keynum,2], n_matches [B], dr_index [B,keynum],
        poses [B,4,4], recall [B], best_overlap [B]."""
        self._chk(desc, torch.float32, "desc"); self._chk(keys, torch.float64, "keys")
        self._chk(pair_cloud, torch.int32, "pair_cloud")
        n_clouds, n = desc.shape[0], desc.shape[1]
        B = pair_cloud.shape[0]
        S = keynum or n
        if out is None:
            dev = self.device
            out = dict(matches=torch.empty((B, S, 2), dtype=torch.int32, device=dev),
                       n_matches=torch.empty(B, dtype=torch.int32, device=dev),
                       dr_index=torch.empty((B, S), dtype=torch.int32, device=dev),
                       poses=torch.empty((B, 4, 4), dtype=torch.float64, device=dev),
                       recall=torch.empty(B, dtype=torch.int32, device=dev),
                       best_overlap=torch.empty(B, dtype=torch.float64, device=dev))
        b = _lib.RoregBatch()
        b.n_clouds, b.n, b.keynum, b.B = n_clouds, n, S, B
        b.desc, b.keys, b.pair_cloud, b.sample = desc.data_ptr(), keys.data_ptr(), pair_cloud.data_ptr(), (sample.data_ptr() if sample is not None else None)
        b.nn_mode, b.estimator, b.max_iter, b.ird, b.seed = nn_mode, estimator, max_iter, float(ird), int(seed)
        b.triplets = triplets.data_ptr() if triplets is not None else None
        b.hyp_host_svd = hyps.data_ptr() if hyps is not None else None
        b.matches, b.n_matches, b.dr_index = out["matches"].data_ptr(), out["n_matches"].data_ptr(), out["dr_index"].data_ptr()
        b.poses, b.recall, b.best_overlap = out["poses"].data_ptr(), out["recall"].data_ptr(), out["best_overlap"].data_ptr()
        if pipelined:
            rc = self.lib.roreg_register_batch_pipelined(self.h, C.byref(b), _stream())
            _lib.check(self.h, rc, "roreg_register_batch_pipelined")
            self._pipe_keep = (self._pipe_keep + [(out, keys, pair_cloud, sample, triplets, hyps)])[-2:]   # inputs of the owed tail stay alive
            return out
        rc = self.lib.roreg_register_batch(self.h, C.byref(b), _stream())
        _lib.check(self.h, rc, "roreg_register_batch")
        self._last_batch = b       # kept for estimate_batch (same buffers; `out` keeps the tensors alive)
        return out

    def register_batch_pipelined(self, *args, **kw):
        """Throughput form of register_batch for back-to-back batches (roreg_register_batch_pipelined): the RANSAC tail of the
        previous batch runs beside this batch's pooling.  `poses`, `recall`, `best_overlap` of the returned dict are complete only
        after the NEXT register_batch_pipelined / flush_batches on the same stream - pass a different `out` (and different keys /
        pair_cloud tensors if they change) for consecutive calls."""
        return self.register_batch(*args, pipelined=True, **kw)

    def flush_batches(self):
        """Enqueue the tail still owed by the last register_batch_pipelined call."""
        _lib.check(self.h, self.lib.roreg_register_batch_flush(self.h, _stream()), "roreg_register_batch_flush")

    def estimate_batch(self, out, hyps, n_hyp=None):
        """One-shot RANSAC + refine on caller-provided hypotheses [B,max_iter,3,4] f64 after register_batch(estimator=2)."""
        self._chk(hyps, torch.float64, "hyps")
        rc = self.lib.roreg_estimate_batch(self.h, C.byref(self._last_batch), _ptr(hyps), _ptr(n_hyp), _stream())
        _lib.check(self.h, rc, "roreg_estimate_batch")
        return out
