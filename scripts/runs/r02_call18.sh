#!/bin/bash
# Round 2, GPU call 18: where does the gather GEMM lose its time?  (debug knobs: results wrong, timing only)
set -x
mkdir -p gpurun_out
cat > /tmp/tg.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth
ctx = ops.Context(0); rng = np.random.default_rng(0)
x = rng.standard_normal((5000, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True); xd = ctx.dev(x)
gf = nets.GFNet(ctx, synth.random_weights("GF", 101), npass=1, chunk=500)
for _ in range(2): gf.forward(xd)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): gf.forward(xd)
e1.record(); torch.cuda.synchronize()
print(f"GF npass 1: {e0.elapsed_time(e1) / 5:.2f} ms")
PY
for g in ldg tma; do for d in 0 1 2; do echo -n "gather=$g dbg=$d: "; ROREG_GEMM_GATHER=$g ROREG_DEBUG_GEMM=$d timeout 300 python /tmp/tg.py 2>&1 | tail -1; done; done | tee gpurun_out/c18_gemm_dbg.txt
ROREG_GEMM_GATHER=ldg timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 1 -o gpurun_out/c18_gemm_ldg python scripts/gf_one_chunk.py 1 > gpurun_out/c18_ncu.log 2>&1
ncu -i gpurun_out/c18_gemm_ldg.ncu-rep --page source --csv --print-source sass > gpurun_out/c18_gemm_ldg_source.csv 2>/dev/null; rm -f gpurun_out/c18_gemm_ldg.ncu-rep
