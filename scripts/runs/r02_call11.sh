#!/bin/bash
# Round 2, GPU call 11: full GPU suite + smoke at the current state; Match_ot timing; ncu of the gather GEMM layers of GF.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c11_pytest_all.txt 2>&1; tail -4 gpurun_out/c11_pytest_all.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c11_smoke.txt 2>&1; tail -2 gpurun_out/c11_smoke.txt
timeout 300 python /dev/stdin > gpurun_out/c11_matchot.txt 2>&1 <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, synth, matchot
ctx = ops.Context(0); ctx.set_corr_mode(3)
pr = synth.make_pair(2, n=5000)
f0 = ctx.dev(pr["feats0"]); f1 = ctx.dev(pr["feats1"]); k0 = ctx.dev(pr["keys0"].astype(np.float32)); k1 = ctx.dev(pr["keys1"].astype(np.float32))
for npass in (1, 3):
    mo = matchot.MatchOT(ctx, synth.random_weights("RM", 104), npass=npass)
    for _ in range(2): mo.forward(f1, f0, k1, k0)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches; e0.record(); m0, s0 = mo.forward(f1, f0, k1, k0); e1.record(); torch.cuda.synchronize()
    print(f"Match_ot 5000 x 5000 npass {npass} eager: {e0.elapsed_time(e1):.2f} ms, {ctx.launches - l0} launches, matched {int((m0 >= 0).sum())}")
    mo.forward_graphed(f1, f0, k1, k0); torch.cuda.synchronize()
    e0.record(); mg, sg = mo.forward_graphed(f1, f0, k1, k0); e1.record(); torch.cuda.synchronize()
    print(f"Match_ot 5000 x 5000 npass {npass} CUDA-graph replay: {e0.elapsed_time(e1):.2f} ms, equal {bool((mg == m0).all())}")
PY
cat gpurun_out/c11_matchot.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 2 -o gpurun_out/c11_gemm_gather python scripts/gf_one_chunk.py 1 > gpurun_out/c11_ncu.log 2>&1; tail -2 gpurun_out/c11_ncu.log
ncu -i gpurun_out/c11_gemm_gather.ncu-rep --page raw --csv > gpurun_out/c11_gemm_gather_raw.csv 2>/dev/null
ncu -i gpurun_out/c11_gemm_gather.ncu-rep --page source --csv --print-source sass > gpurun_out/c11_gemm_gather_source.csv 2>/dev/null
rm -f gpurun_out/c11_gemm_gather.ncu-rep
