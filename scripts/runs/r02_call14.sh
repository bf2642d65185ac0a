#!/bin/bash
# Round 2, GPU call 14: pipelined scene driver with the C file writer (drop-in tests + bench e2e); Sinkhorn barrier experiments.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_matchot.py tests/test_gpu_parity.py -x -q -m gpu -k "scene or dropin or sinkhorn or kabsch3 or reproduce" > gpurun_out/c14_pytest.txt 2>&1; tail -4 gpurun_out/c14_pytest.txt
cat > /tmp/mo.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, synth, matchot
ctx = ops.Context(0); ctx.set_corr_mode(3)
pr = synth.make_pair(2, n=5000)
f0 = ctx.dev(pr["feats0"]); f1 = ctx.dev(pr["feats1"]); k0 = ctx.dev(pr["keys0"].astype(np.float32)); k1 = ctx.dev(pr["keys1"].astype(np.float32))
mo = matchot.MatchOT(ctx, synth.random_weights("RM", 104), npass=1)
for _ in range(2): mo.forward(f1, f0, k1, k0)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); m0, s0 = mo.forward(f1, f0, k1, k0); e1.record(); torch.cuda.synchronize()
print("Match_ot eager:", round(e0.elapsed_time(e1), 2), "ms; matched", int((m0 >= 0).sum()))
PY
for dbg in 0 1 2; do ROREG_DEBUG_SINK=$dbg timeout 300 python /tmp/mo.py 2>&1 | tail -1 | sed "s/^/ROREG_DEBUG_SINK=$dbg: /"; done | tee gpurun_out/c14_sink_dbg.txt
timeout 900 python bench.py --steps 20 --extras 0 > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c14_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e(scene)", round(d["e2e"]["value"]), d["e2e"]["seconds"], "readers", d["e2e"]["reader_threads_per_rank"], "err", d["e2e"]["max_abs_err_vs_gt"], "pair_upload", round(d["e2e_pair_upload"]["value"]))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/c14_bench.err").read()[-2500:])
PY
