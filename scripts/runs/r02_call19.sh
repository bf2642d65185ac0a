#!/bin/bash
# Round 2, GPU call 19: clock64 timeline of the gather GEMM (trace build: -DROREG_GEMM_TRACE) + GF chunk-size sweep
set -x
mkdir -p gpurun_out
for g in ldg tma; do for l in 1 2; do
  ROREG_GEMM_GATHER=$g ROREG_DEBUG_GEMM_TRACE=gpurun_out/c19_trace_${g}_$l.txt ROREG_DEBUG_GEMM_TRACE_LAUNCH=$l timeout 300 python scripts/gf_one_chunk.py 1 > /dev/null 2>&1
done; done
cat > /tmp/tg.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth
ctx = ops.Context(0); rng = np.random.default_rng(0)
x = rng.standard_normal((5000, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True); xd = ctx.dev(x)
for chunk in (500, 1250, 2500, 5000):
    gf = nets.GFNet(ctx, synth.random_weights("GF", 101), npass=1, chunk=chunk)
    for _ in range(2): gf.forward(xd)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gf.forward(xd)
    e1.record(); torch.cuda.synchronize()
    print(f"GF npass 1 chunk {chunk}: {e0.elapsed_time(e1) / 5:.2f} ms", flush=True)
    del gf
PY
ROREG_GEMM_GATHER=tma timeout 300 python /tmp/tg.py 2>&1 | tail -4 | tee gpurun_out/c19_chunks.txt
