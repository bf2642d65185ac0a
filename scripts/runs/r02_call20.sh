#!/bin/bash
# Round 2, GPU call 20: cp.async loader + 8 epilogue warps (trace build): correctness, timeline, timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/c20_pytest.txt
for p in 4 8; do for l in 1 2; do
  ROREG_GEMM_PRODUCERS=$p ROREG_DEBUG_GEMM_TRACE=gpurun_out/c20_trace_cp${p}_$l.txt ROREG_DEBUG_GEMM_TRACE_LAUNCH=$l timeout 300 python scripts/gf_one_chunk.py 1 > /dev/null 2>&1
done; done
cat > /tmp/tg.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth
ctx = ops.Context(0); rng = np.random.default_rng(0)
x = rng.standard_normal((5000, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True); xd = ctx.dev(x)
for npass in (1, 3):
  for chunk in (500, 2500, 5000):
    gf = nets.GFNet(ctx, synth.random_weights("GF", 101), npass=npass, chunk=chunk)
    for _ in range(2): gf.forward(xd)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gf.forward(xd)
    e1.record(); torch.cuda.synchronize()
    print(f"GF npass {npass} chunk {chunk}: {e0.elapsed_time(e1) / 5:.2f} ms", flush=True)
    del gf
PY
for p in 4 8; do echo "producers $p"; ROREG_GEMM_PRODUCERS=$p timeout 300 python /tmp/tg.py 2>&1 | tail -6; done | tee gpurun_out/c20_chunks.txt
