#!/bin/bash
# Round 2, GPU call 9: GEMM under the 196 KiB carve-out + vectorised running-maximum epilogue; Sinkhorn with 8/16 loads in flight.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matchot.py tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/c9_pytest.txt 2>&1; tail -4 gpurun_out/c9_pytest.txt
timeout 300 python scripts/ab_allpairs_r01.py now > gpurun_out/c9_ab_now.txt 2>&1; tail -3 gpurun_out/c9_ab_now.txt
timeout 900 python scripts/time_nets.py > gpurun_out/c9_nets_timing.txt 2> gpurun_out/c9_nets_timing.err; cat gpurun_out/c9_nets_timing.txt; tail -3 gpurun_out/c9_nets_timing.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c9_nets_launches.csv python scripts/time_nets.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/c9_nets_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
agg = collections.OrderedDict()
if hdr:
    h = rows[hdr[0]]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
    for r in rows[hdr[0] + 1:]:
        if len(r) > mv:
            try: v = float(r[mv].replace(",", ""))
            except ValueError: continue
            a = agg.setdefault(r[kn][:60], [0, 0.0]); a[0] += 1; a[1] += v
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print(f"{t/1e3:10.1f} us total {n:5d} launches {t/n/1e3:9.1f} us avg  {k}")
PY
timeout 900 python bench.py --steps 20 > gpurun_out/c9_bench_full.json 2> gpurun_out/c9_bench_full.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c9_bench_full.json").read().strip().splitlines()[-1])
    print("full:", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    y=d.get("workload_yohoo",{}); print("yohoo:", y.get("value"), y.get("stage_ms_per_step"), y.get("roofline",{}).get("frac"), y.get("unavailable"))
    print("match_ot:", json.dumps(d.get("workload_match_ot"))[:500])
except Exception as e:
    print("full FAILED", e); print(open("gpurun_out/c9_bench_full.err").read()[-2500:])
PY
