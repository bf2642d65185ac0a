#!/bin/bash
# Round 2, GPU call 10: vectorised Sinkhorn passes.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matchot.py -x -q -m gpu > gpurun_out/c10_pytest.txt 2>&1; tail -4 gpurun_out/c10_pytest.txt
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
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); m0, s0 = mo.forward(f1, f0, k1, k0); e1.record(); torch.cuda.synchronize()
print("Match_ot eager:", e0.elapsed_time(e1), "ms; matched", int((m0 >= 0).sum()))
mo.forward_graphed(f1, f0, k1, k0); torch.cuda.synchronize()
e0.record(); mg, sg = mo.forward_graphed(f1, f0, k1, k0); e1.record(); torch.cuda.synchronize()
print("Match_ot graph replay:", e0.elapsed_time(e1), "ms; equal", bool((mg == m0).all()))
PY
timeout 300 python /tmp/mo.py > gpurun_out/c10_matchot.txt 2>&1; cat gpurun_out/c10_matchot.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/c10_mo_launches.csv python /tmp/mo.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/c10_mo_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
agg = collections.OrderedDict()
if hdr:
    h = rows[hdr[0]]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
    for r in rows[hdr[0] + 1:]:
        if len(r) > mv:
            try: v = float(r[mv].replace(",", ""))
            except ValueError: continue
            a = agg.setdefault(r[kn][:60], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(t for _, t in agg.values())
    print("total kernel time over the 3 eager forwards + capture + replays (us):", tot / 1e3)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print(f"{t/1e3:10.1f} us total {n:5d} launches {t/n/1e3:9.1f} us avg  {k}")
PY
