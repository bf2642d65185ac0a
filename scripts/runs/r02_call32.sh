#!/bin/bash
# Round 2, GPU call 32: validation of the final build - full GPU suite, smoke, default bench + reference arm, ncu launch list and
# ncu --set full of the step's kernels (the files profiles/r02_* are made from).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c32_pytest_all.txt 2>&1; tail -4 gpurun_out/c32_pytest_all.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c32_smoke.txt 2>&1; tail -2 gpurun_out/c32_smoke.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c32_bench_reference.json 2> gpurun_out/c32_bench_reference.err; tail -c 600 gpurun_out/c32_bench_reference.json
timeout 900 python bench.py > gpurun_out/c32_bench_default.json 2> gpurun_out/c32_bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c32_bench_default.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms/step", d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"]["seconds"], "launches", d["gpu_launches"], "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "fused", {k: round(v,3) for k,v in d["roofline"]["fused_step"].items() if isinstance(v,float)})
    print("stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, "clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    print("yohoo", d["workload_yohoo"].get("value"), d["workload_yohoo"].get("roofline",{}).get("frac"), "match_ot", d["workload_match_ot"].get("ms_per_pair"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/c32_bench_default.err").read()[-2500:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 200 --csv --log-file gpurun_out/c32_launches_raw.csv python bench.py --steps 3 --warmup 3 --cpu-sample-pairs 0 --extras 0 --value-only 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:inv_pool_t4_kernel|nn_tc4_kernel|group_corr_tc3_kernel|ransac_score_pre_kernel|coarse_hyp_kernel|refine_kernel" -s 12 -c 6 -o gpurun_out/c32_full python bench.py --steps 2 --warmup 3 --cpu-sample-pairs 0 --extras 0 --value-only 1 > gpurun_out/c32_ncu_full.log 2>&1; tail -2 gpurun_out/c32_ncu_full.log
ncu -i gpurun_out/c32_full.ncu-rep --page raw --csv > gpurun_out/c32_full_raw.csv 2>/dev/null
rm -f gpurun_out/c32_full.ncu-rep
