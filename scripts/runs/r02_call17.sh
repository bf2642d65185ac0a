#!/bin/bash
# Round 2, GPU call 17: LDG / STS loader for the gathered GEMM operand vs the TMA gather4 producer.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_matchot.py -x -q -m gpu > gpurun_out/c17_pytest.txt 2>&1; tail -3 gpurun_out/c17_pytest.txt
cat > /tmp/tn.py <<'PY'
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
    rd = nets.RDNet(ctx, synth.random_weights("RD", 103), npass=npass, chunk=1000)
    print(f"npass {npass}: GF {timed(lambda: gf.forward(xd)):.2f} ms   ET {timed(lambda: et.forward(xd, rows, xd, rows, xd, rows, xd, rows, pre)):.2f} ms   RD {timed(lambda: rd.forward(xd)):.2f} ms")
PY
for mode in ldg tma; do echo "gather = $mode"; ROREG_GEMM_GATHER=$mode timeout 600 python /tmp/tn.py 2>&1 | tail -2; done | tee gpurun_out/c17_nets_gather_ab.txt
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 5 -c 2 -o gpurun_out/c17_gemm_ldg python scripts/gf_one_chunk.py 1 > gpurun_out/c17_ncu.log 2>&1; tail -1 gpurun_out/c17_ncu.log
ncu -i gpurun_out/c17_gemm_ldg.ncu-rep --page raw --csv > gpurun_out/c17_gemm_ldg_raw.csv 2>/dev/null; rm -f gpurun_out/c17_gemm_ldg.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/c17_gemm_ldg_raw.csv"))); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d.get("gpu__time_duration.sum"), d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"), d.get("launch__block_size"))
PY
