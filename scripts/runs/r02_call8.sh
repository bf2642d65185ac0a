#!/bin/bash
# Round 2, GPU call 8: all-pairs epilogue without the serialised read-modify-write chain; ncu of the fused Sinkhorn kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matchot.py tests/test_gpu_nets.py -x -q -m gpu -s > gpurun_out/c8_pytest.txt 2>&1; tail -5 gpurun_out/c8_pytest.txt; grep "yoho_mat vs" gpurun_out/c8_pytest.txt
timeout 300 python scripts/ab_allpairs_r01.py now > gpurun_out/c8_ab_now.txt 2>&1; tail -3 gpurun_out/c8_ab_now.txt
timeout 900 python scripts/time_nets.py > gpurun_out/c8_nets_timing.txt 2> gpurun_out/c8_nets_timing.err; cat gpurun_out/c8_nets_timing.txt; tail -3 gpurun_out/c8_nets_timing.err
cat > /tmp/mo.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, synth, matchot
ctx = ops.Context(0)
pr = synth.make_pair(2, n=5000)
f0 = ctx.dev(pr["feats0"]); f1 = ctx.dev(pr["feats1"]); k0 = ctx.dev(pr["keys0"].astype(np.float32)); k1 = ctx.dev(pr["keys1"].astype(np.float32))
mo = matchot.MatchOT(ctx, synth.random_weights("RM", 104), npass=1)
for _ in range(2): mo.forward(f1, f0, k1, k0)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_fused -s 1 -c 1 -o gpurun_out/c8_sinkhorn python /tmp/mo.py > gpurun_out/c8_ncu.log 2>&1; tail -2 gpurun_out/c8_ncu.log
ncu -i gpurun_out/c8_sinkhorn.ncu-rep --page source --csv --print-source sass > gpurun_out/c8_sinkhorn_source.csv 2>/dev/null
ncu -i gpurun_out/c8_sinkhorn.ncu-rep --page raw --csv > gpurun_out/c8_sinkhorn_raw.csv 2>/dev/null
rm -f gpurun_out/c8_sinkhorn.ncu-rep
