"""roreg_b200 - B200-native (sm_100a) implementation of RoReg's per-pair registration hot path.

    ops        thin ctypes front end of libroreg_b200.so (include/roreg_b200.h): Context, one method per C entry point
    test       mirrors of the reference's plugin classes (test/{extractor,detector,matcher,estimator}.py): Test.py runs unchanged
    scene      whole-dataset driver on the batched engine (clouds uploaded once, pairs sharded over ranks, reference file contract)
    dataio     the reference's dataset duck type over its origin-data directories without open3d
    nets, matchot, pipeline   schedules of the GF / ET / RD networks, Match_ot and the yohoo engine on the C ABI's GEMM / glue kernels
    shard      pair sharding + final gather of poses (torch.distributed)
    group, synth   icosahedral group tables; seeded synthetic pairs / scenes

Nothing here falls back to the CPU: without the CUDA library or a GPU the entry points raise (see _lib.RoregLibraryError).
Submodules are imported on demand (importing the package does not load the shared library).
"""
