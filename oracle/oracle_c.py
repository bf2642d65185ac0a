"""ctypes front end of oracle/oracle_c.c (C + pthreads restatement).  TEST INFRASTRUCTURE ONLY - the
timed CPU baseline of bench.py and a second checker; never imported by the product."""
import ctypes as C
import os
import numpy as np
from . import roreg_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.orc_threads.restype = C.c_int
    return _lib


def threads():
    return int(lib().orc_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def inv_pool(feats, normalise=True):
    feats = np.ascontiguousarray(feats, np.float32)
    out = np.empty((feats.shape[0], 32), np.float32)
    lib().orc_inv_pool(_p(feats), C.c_int(feats.shape[0]), C.c_int(int(normalise)), _p(out))
    return out


def nn(target, source):
    t = np.ascontiguousarray(target, np.float32); s = np.ascontiguousarray(source, np.float32)
    idx = np.empty(s.shape[0], np.int32); d = np.empty(s.shape[0], np.float32)
    lib().orc_nn(_p(t), C.c_int(t.shape[0]), _p(s), C.c_int(s.shape[0]), C.c_int(t.shape[1]), _p(idx), _p(d))
    return d, idx


def des2r(X, Y, ix, iy, perm, want_cor=False):
    X = np.ascontiguousarray(X, np.float32); Y = np.ascontiguousarray(Y, np.float32)
    ix = np.ascontiguousarray(ix, np.int32); iy = np.ascontiguousarray(iy, np.int32)
    pm = np.ascontiguousarray(perm, np.int32)
    out = np.empty(ix.shape[0], np.int32)
    cor = np.empty((ix.shape[0], 60), np.float32) if want_cor else None
    lib().orc_des2r(_p(X), _p(Y), _p(ix), _p(iy), C.c_int(ix.shape[0]), _p(pm), _p(out), _p(cor) if want_cor else None)
    return (out, cor) if want_cor else out


def score(k0, k1, scores, hyps, ird):
    k0 = np.ascontiguousarray(k0, np.float64); k1 = np.ascontiguousarray(k1, np.float64)
    sc = np.ascontiguousarray(scores, np.float64); hy = np.ascontiguousarray(hyps, np.float64)
    ov = np.empty(hy.shape[0], np.float64)
    lib().orc_score(_p(k0), _p(k1), _p(sc), C.c_int(k0.shape[0]), _p(hy), C.c_int(hy.shape[0]), C.c_double(ird), _p(ov))
    return ov


def register_pair(pr, tables, max_iter, ird, seed=0):
    """mutual matcher -> Rindex -> yohoc RANSAC -> 2x refine for one pair, hot loops in C/pthreads, the
    O(iterations) host logic (draws, 3-point Kabsch, refiner) in NumPy exactly as the reference."""
    f0 = inv_pool(pr["feats0"]); f1 = inv_pool(pr["feats1"])
    _, nn01 = nn(f1, f0); _, nn10 = nn(f0, f1)
    i = np.arange(f0.shape[0]); keep = nn10[nn01] == i
    pps = np.stack([i[keep], nn01[keep]], 1).astype(np.int64)
    dr = des2r(pr["feats1"], pr["feats0"], pps[:, 1], pps[:, 0], tables.perm).astype(np.int64)
    k0 = pr["keys0"][pps[:, 0]]; k1 = pr["keys1"][pps[:, 1]]
    sc = np.ones(pps.shape[0])
    draws, _, prob = O.yohoc_draws(dr, max_iter, np.random.RandomState(seed))
    hyps = np.stack([O.threepps2tran(k0[d[1]], k1[d[1]]) for d in draws])
    ov = score(k0, k1, sc, hyps, ird)
    best = int(np.argmax(ov))                                   # first maximum == strict '>' scan
    return O.refine(k0, k1, hyps[best], sc, ird), pps, dr
