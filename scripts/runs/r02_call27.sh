#!/bin/bash
# Round 2, GPU call 27: chunked scene loader: GPU tests that go through it + e2e with 8 / 12 / 15 reader threads
set -x
mkdir -p gpurun_out
nproc
timeout 900 python -m pytest tests -q -m gpu -k "scene or dropin or evaluator or plugin" 2>&1 | tail -3 | tee gpurun_out/c27_pytest.txt
for r in 8 12 15; do
  ROREG_SCENE_READERS=$r timeout 600 python bench.py --steps 10 --warmup 3 --extras 0 --cpu-sample-pairs 0 > gpurun_out/c27_bench_r$r.json 2> gpurun_out/c27_bench_r$r.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c27_bench_r$r.json").read().strip().splitlines()[-1])
    print("readers $r: value", round(d["value"]), "e2e(scene)", round(d["e2e"]["value"]), d["e2e"].get("seconds"), "h2d GB/s", round(d["e2e"].get("h2d_gb_per_s_per_rank",0),1), "err", d["e2e"].get("max_abs_err_vs_gt"))
except Exception as e:
    print("readers $r FAILED", e); print(open("gpurun_out/c27_bench_r$r.err").read()[-2000:])
PY
done | tee gpurun_out/c27_readers.txt
