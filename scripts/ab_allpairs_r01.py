"""A/B of roreg_group_corr_allpairs: this build against the round-1 library (scripts/ab/libroreg_b200_r01.so, built from commit
12a8102 by hand) on the same inputs.  GPU box only; not part of the product or the tests."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
which = sys.argv[1]
from roreg_b200 import _lib
if which == "r01":
    import ctypes as C
    _lib.LIB_PATH = "scripts/ab/libroreg_b200_r01.so"
    probe = C.CDLL(_lib.LIB_PATH)
    for name in list(_lib.SIGNATURES):
        if not hasattr(probe, name):
            del _lib.SIGNATURES[name]
from roreg_b200 import ops, nets, synth
from roreg_b200.ops import _ptr, _stream
ctx = ops.Context(0)
pr = synth.make_pair(2, n=5000)
N = M = 5000
for npass in (1, 3):
    g = nets.GroupNets(ctx, npass)
    xh, xl = g.pack([ctx.dev(pr["feats1"])], [None], [0], None, N); yh, yl = g.pack([ctx.dev(pr["feats0"])], [None], [0], None, M)
    best = torch.empty((N, M), dtype=torch.float32, device=ctx.device); ba = torch.empty((N, M), dtype=torch.uint8, device=ctx.device)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        e0.record()
        rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, _ptr(xh), _ptr(xl), N, _ptr(yh), _ptr(yl), M, npass, _ptr(best), _ptr(ba), None, None, None, _stream())
        e1.record(); torch.cuda.synchronize()
    print(f"{which}: all-pairs npass {npass}: {e0.elapsed_time(e1):.2f} ms, checksum {float(best.double().sum()):.6e} {int(ba.long().sum())}")
