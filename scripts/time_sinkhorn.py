"""GPU-box timing of roreg_sinkhorn_match alone (100 iterations + assignment) on a 5000 x 5000 score matrix:
the persistent kernel (default) against ROREG_SINKHORN_LAUNCHES=1 (200 launches), with the equality of the assignments.   python scripts/time_sinkhorn.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from roreg_b200 import ops, _lib

ctx = ops.Context(0)
lib = ctx.lib
m = n = 5000
rng = np.random.default_rng(3)
S = ctx.dev((rng.standard_normal((m, n)) * 3).astype(np.float32))
res = {}
for mode in ("0", "1"):
    os.environ.pop("ROREG_SINKHORN_LAUNCHES", None)
    if mode == "1":
        os.environ["ROREG_SINKHORN_LAUNCHES"] = "1"
    u = torch.empty(m + 1, dtype=torch.float32, device=ctx.device); v = torch.empty(n + 1, dtype=torch.float32, device=ctx.device)
    m0 = torch.empty(m, dtype=torch.int32, device=ctx.device); s0 = torch.empty(m, dtype=torch.float32, device=ctx.device)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: _lib.check(ctx.h, lib.roreg_sinkhorn_match(ctx.h, S.data_ptr(), m, n, n, 1.0, 100, u.data_ptr(), v.data_ptr(), m0.data_ptr(), s0.data_ptr(), st), "sinkhorn")
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        call()
    e1.record(); torch.cuda.synchronize()
    res[mode] = (u.clone(), v.clone(), m0.clone())
    print(f"{'200 launches' if mode == '1' else 'persistent kernel'}: {e0.elapsed_time(e1) / 5:.2f} ms per call (100 iterations + assignment)", flush=True)
du = (res["1"][0] - res["0"][0]).abs().max().item(); dv = (res["1"][1] - res["0"][1]).abs().max().item()
print(f"max |u_launches - u| {du:.2e}, max |v_launches - v| {dv:.2e}, assignments equal: {bool(torch.equal(res['1'][2], res['0'][2]))}, matched {int((res['1'][2] >= 0).sum())}")
