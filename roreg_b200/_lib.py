"""ctypes binding of libroreg_b200.so (the C ABI declared in include/roreg_b200.h).

There is NO CPU fallback: if the CUDA library cannot be loaded every product entry point raises.
Build it with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ROREG_B200_LIB: another build of the same library (e.g. the -DROREG_GEMM_TRACE debugging build)
LIB_PATH = os.environ.get("ROREG_B200_LIB") or os.path.join(_HERE, "csrc", "libroreg_b200.so")

_p = C.c_void_p
_i = C.c_int
_d = C.c_double


class RoregBatch(C.Structure):
    """struct roreg_batch (include/roreg_b200.h)."""
    _fields_ = [("n_clouds", C.c_int32), ("n", C.c_int32), ("keynum", C.c_int32), ("B", C.c_int32),
                ("desc", _p), ("keys", _p), ("pair_cloud", _p), ("sample", _p),
                ("nn_mode", C.c_int32), ("estimator", C.c_int32), ("max_iter", C.c_int32),
                ("ird", _d), ("seed", C.c_uint64), ("triplets", _p), ("hyp_host_svd", _p),
                ("matches", _p), ("n_matches", _p), ("dr_index", _p), ("poses", _p), ("recall", _p),
                ("best_overlap", _p)]


# name -> (restype, argtypes); every symbol include/roreg_b200.h declares
SIGNATURES = {
    "roreg_version": (_i, []),
    "roreg_ctx_create": (_i, [_i, _p, _p, _p, C.POINTER(_p)]),
    "roreg_ctx_destroy": (_i, [_p]),
    "roreg_last_error": (C.c_char_p, [_p]),
    "roreg_launch_count": (C.c_int64, [_p]),
    "roreg_inv_pool": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "roreg_knn": (_i, [_p, _p, _i, _p, _i, _i, _i, _p, _p, _p]),
    "roreg_mutual_match": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p]),
    "roreg_group_corr": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "roreg_set_corr_mode": (_i, [_p, _i]),
    "roreg_hypotheses_from_quat": (_i, [_p, _p, _p, _p, _p, _i, _p, _p]),
    "roreg_ransac_oneshot": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _i, _d, _p, _p, _p, _p]),
    "roreg_refine": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p, _d, _p, _p, _p]),
    "roreg_refine_once": (_i, [_p, _p, _p, _p, _i, _i, _p, _d, _p, _p, _p]),
    "roreg_kabsch3": (_i, [_p, _p, _p, _p, _i, _p, _p]),
    "roreg_pack_descriptors": (_i, [_p, _i, _p, _p, _p, _p, _i, _p, _p, _i, _p, _p, _p]),
    "roreg_gconv_gemm": (_i, [_p, _p, _p, _i, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _p, _i, _p, _p, _i, _p]),
    "roreg_gemm": (_i, [_p, _p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _p, _i, _p, _p, _i, _p]),
    "roreg_gf_finalize": (_i, [_p, _p, _p, _i, _p, _p]),
    "roreg_rd_finalize": (_i, [_p, _p, _i, _p, _p]),
    "roreg_row_std60": (_i, [_p, _p, _i, _p, _p]),
    "roreg_quat_normalize": (_i, [_p, _p, _i, _i, _p, _p]),
    "roreg_group_corr_allpairs": (_i, [_p, _p, _p, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "roreg_topk_rows": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "roreg_gather_rows": (_i, [_p, _p, _p, C.c_longlong, _i, _p, _p]),
    "roreg_rel_coor": (_i, [_p, _p, _p, _i, _i, C.c_float, _p, _p]),
    "roreg_chan_stats": (_i, [_p, _p, C.c_longlong, _i, _p, _p, _p]),
    "roreg_prep_rows": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _i, C.c_longlong, _i, _p, _p, _p, _p]),
    "roreg_mha": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "roreg_rind_rows": (_i, [_p, _p, _i, _p, _p]),
    "roreg_sinkhorn_match": (_i, [_p, _p, _i, _i, _i, C.c_float, _i, _p, _p, _p, _p, _p]),
    "roreg_register_batch": (_i, [_p, C.POINTER(RoregBatch), _p]),
    "roreg_register_batch_pipelined": (_i, [_p, C.POINTER(RoregBatch), _p]),
    "roreg_register_batch_flush": (_i, [_p, _p]),
    "roreg_estimate_batch": (_i, [_p, C.POINTER(RoregBatch), _p, _p, _p]),
    "roreg_set_timing": (_i, [_p, _i]),
    "roreg_set_overlap": (_i, [_p, _i]),
    "roreg_set_score_mode": (_i, [_p, _i]),
    "roreg_get_stage_ms": (_i, [_p, C.POINTER(C.c_float)]),
    "roreg_write_pair_files": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _p, _p, _i, _p, C.c_longlong]),
}

_lib = None


class RoregLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library and declare every prototype.  No compute happens here."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RoregLibraryError(
            f"{LIB_PATH} is missing - the CUDA extension has not been built (run __graft_entry__.build()). "
            "roreg_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc, what):
    if rc != 0:
        msg = load().roreg_last_error(ctx)
        raise RoregLibraryError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
