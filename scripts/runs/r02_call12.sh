#!/bin/bash
# Round 2, GPU call 12: producer-warp count of the gather GEMM (4 / 8 / 16); rank-deficient Kabsch fix on the device.
set -x
mkdir -p gpurun_out
for np in 4 8 16; do
  ROREG_GEMM_PRODUCERS=$np timeout 600 python -m pytest tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/c12_pytest_nets_p$np.txt 2>&1; tail -2 gpurun_out/c12_pytest_nets_p$np.txt
  ROREG_GEMM_PRODUCERS=$np timeout 600 python - > gpurun_out/c12_nets_p$np.txt 2>&1 <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth
ctx = ops.Context(0); rng = np.random.default_rng(0)
def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
x = rng.standard_normal((5000, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True); xd = ctx.dev(x)
K = 3400; rows = ctx.dev(rng.integers(0, 5000, K).astype(np.int32)); pre = ctx.dev(rng.integers(0, 60, K).astype(np.int32))
for npass in (1, 3):
    gf = nets.GFNet(ctx, synth.random_weights("GF", 101), npass=npass, chunk=500)
    et = nets.ETNet(ctx, synth.random_weights("ET", 102), npass=npass, chunk=1000)
    print(f"npass {npass}: GF {timed(lambda: gf.forward(xd)):.2f} ms   ET {timed(lambda: et.forward(xd, rows, xd, rows, xd, rows, xd, rows, pre)):.2f} ms")
PY
  echo "producers $np:"; cat gpurun_out/c12_nets_p$np.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kabsch or device_draws or register_batch" > gpurun_out/c12_pytest_kabsch.txt 2>&1; tail -2 gpurun_out/c12_pytest_kabsch.txt
