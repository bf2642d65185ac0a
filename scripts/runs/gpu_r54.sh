#!/bin/bash
# run 54: final verification of the round: smoke, full GPU suite, default bench (+cpu baseline), reference arm, ncu launch list, ncu --set full
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r54_smoke.txt 2>&1; tail -2 gpurun_out/r54_smoke.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r54_pytest.txt 2>&1
tail -4 gpurun_out/r54_pytest.txt
timeout 600 python bench.py > gpurun_out/r54_bench_default.json 2> gpurun_out/r54_bench_default.err
tail -2 gpurun_out/r54_bench_default.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r54_bench_reference.json 2> gpurun_out/r54_bench_reference.err
tail -2 gpurun_out/r54_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 150 --csv --log-file gpurun_out/r54_launches_raw.csv python bench.py --steps 3 --warmup 3 --cpu-sample-pairs 0 > gpurun_out/r54_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nn_tc4_kernel|group_corr_tc3_kernel|inv_pool_t4_kernel|ransac_score_kernel' -s 12 -c 4 -o gpurun_out/r54_prof -f python bench.py --steps 2 --warmup 3 --cpu-sample-pairs 0 > gpurun_out/r54_ncu_full.log 2>&1
ls -la gpurun_out/r54_prof.ncu-rep
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r54_bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(round(d["value"]), "pairs/s e2e", round(d["e2e"]["value"]), "scene", round(d["e2e_scene"]["value"]), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("reference_ops"), r["kernel"], round(r["frac"],3), {k:round(v,3) for k,v in r["stage_ms_per_step"].items()}, round(r["fused_step"]["hbm_frac"],3), round(r["fused_step"]["hbm_frac_survey_8d"],3), d["clocks"], d["gpu_launches"])
d=json.loads(open("gpurun_out/r54_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", d["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("reference_ops"))
PY
