#!/bin/bash
# Round 2, GPU call 3: group_corr_tc3 epilogue diet (packed smem read table, cached match counts, one barrier less); gather4 micro-test.
set -x
mkdir -p gpurun_out
./scripts/gather4_test > gpurun_out/c3_gather4.txt 2>&1; cat gpurun_out/c3_gather4.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_matchot.py tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/c3_pytest.txt 2>&1
tail -5 gpurun_out/c3_pytest.txt
run_bench() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py "$@" --cpu-sample-pairs 0 --value-only 1 > gpurun_out/c3_bench_$tag.json 2> gpurun_out/c3_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c3_bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag:", round(d["value"]), "pairs/s", {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d.get("pose_check"))
except Exception as e:
    print("$tag: FAILED", e)
PY
}
run_bench base
ROREG_SCORE_CTAS_PER_SM=2 run_bench pipe_score1_cap2 --pipelined 1 --score-mode 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_corr_tc3 -c 1 -o gpurun_out/c3_corr3 python bench.py --steps 2 --warmup 1 --cpu-sample-pairs 0 --value-only 1 > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log
ncu -i gpurun_out/c3_corr3.ncu-rep --page source --csv --print-source sass > gpurun_out/c3_corr3_source.csv 2>/dev/null
ncu -i gpurun_out/c3_corr3.ncu-rep --page raw --csv > gpurun_out/c3_corr3_raw.csv 2>/dev/null
