#!/bin/bash
# Round 2, GPU call 26 (--gpus 8): weak scaling of the final build at N = 2, 4, 8 (value, scene e2e) + N = 1 on the same box
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 600 python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/c26_bench_g1.json 2> gpurun_out/c26_bench_g1.err
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --extras 0 > gpurun_out/c26_bench_g$n.json 2> gpurun_out/c26_bench_g$n.err
done
for n in 1 2 4 8; do
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c26_bench_g$n.json").read().strip().splitlines()[-1])
    print("N=$n value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e(scene)", round(d["e2e"]["value"]), d["e2e"].get("seconds"), "h2d GB/s/rank", round(d["e2e"].get("h2d_gb_per_s_per_rank",0),1), "readers", d["e2e"].get("reader_threads_per_rank"), "cores", d["e2e"].get("host_cores"), d["clocks"])
except Exception as e:
    print("N=$n FAILED", e); print(open("gpurun_out/c26_bench_g$n.err").read()[-2000:])
PY
done | tee gpurun_out/c26_scaling_summary.txt
