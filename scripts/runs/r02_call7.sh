#!/bin/bash
# Round 2, GPU call 7: fast-exp branch-free Sinkhorn accumulation, chan_stats block reduction, CUDA-graph replay of Match_ot;
# A/B of the all-pairs GEMM against the round-1 library; ncu of the plain-mode GEMM.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matchot.py tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/c7_pytest.txt 2>&1; tail -5 gpurun_out/c7_pytest.txt
timeout 300 python scripts/ab_allpairs_r01.py r01 > gpurun_out/c7_ab_r01.txt 2>&1; tail -3 gpurun_out/c7_ab_r01.txt
timeout 300 python scripts/ab_allpairs_r01.py now > gpurun_out/c7_ab_now.txt 2>&1; tail -3 gpurun_out/c7_ab_now.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 70 -c 1 -o gpurun_out/c7_gemm_plain python scripts/ab_allpairs_r01.py now > gpurun_out/c7_ncu1.log 2>&1; tail -2 gpurun_out/c7_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 70 -c 1 -o gpurun_out/c7_gemm_plain_r01 python scripts/ab_allpairs_r01.py r01 > gpurun_out/c7_ncu2.log 2>&1; tail -2 gpurun_out/c7_ncu2.log
for f in c7_gemm_plain c7_gemm_plain_r01; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
timeout 900 python bench.py --steps 20 > gpurun_out/c7_bench_full.json 2> gpurun_out/c7_bench_full.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c7_bench_full.json").read().strip().splitlines()[-1])
    print("full:", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    y=d.get("workload_yohoo",{}); print("yohoo:", y.get("value"), y.get("stage_ms_per_step"), y.get("roofline",{}).get("frac"), y.get("unavailable"))
    print("match_ot:", json.dumps(d.get("workload_match_ot"))[:700])
except Exception as e:
    print("full FAILED", e); print(open("gpurun_out/c7_bench_full.err").read()[-2500:])
PY
