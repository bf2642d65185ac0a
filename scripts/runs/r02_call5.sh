#!/bin/bash
# Round 2, GPU call 5: implicit group-convolution GEMM (TMA gather4, one operand load per k-chunk): parity of the networks,
# timings; bench defaults (pipelined + score mode 1) with the extra tensor-core workloads.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_nets.py tests/test_gpu_matchot.py -x -q -m gpu > gpurun_out/c5_pytest_nets.txt 2>&1
tail -6 gpurun_out/c5_pytest_nets.txt
timeout 900 python scripts/time_nets.py > gpurun_out/c5_nets_timing.txt 2> gpurun_out/c5_nets_timing.err; cat gpurun_out/c5_nets_timing.txt; tail -3 gpurun_out/c5_nets_timing.err
timeout 900 python bench.py > gpurun_out/c5_bench_full.json 2> gpurun_out/c5_bench_full.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c5_bench_full.json").read().strip().splitlines()[-1])
    print("full:", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["seconds"], "stages", {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()})
    print("yohoo:", json.dumps(d.get("workload_yohoo"))[:1500])
    print("match_ot:", json.dumps(d.get("workload_match_ot"))[:600])
    print("cpu:", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"))
except Exception as e:
    print("full FAILED", e); print(open("gpurun_out/c5_bench_full.err").read()[-2500:])
PY
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/c5_pytest_all.txt 2>&1; tail -4 gpurun_out/c5_pytest_all.txt
