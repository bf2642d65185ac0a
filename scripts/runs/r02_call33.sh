#!/bin/bash
# Round 2, GPU call 33 (--gpus 2): torchrun check of the final build (chunked scene loader, new GEMM) at N = 2
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --extras 0 --cpu-sample-pairs 0 > gpurun_out/c33_bench_g2.json 2> gpurun_out/c33_bench_g2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c33_bench_g2.json").read().strip().splitlines()[-1])
    print("N=2 value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e(scene)", round(d["e2e"]["value"]), d["e2e"].get("seconds"), "h2d GB/s/rank", round(d["e2e"].get("h2d_gb_per_s_per_rank",0),1), "readers", d["e2e"].get("reader_threads_per_rank"), "err", d["e2e"].get("max_abs_err_vs_gt"))
except Exception as e:
    print("N=2 FAILED", e); print(open("gpurun_out/c33_bench_g2.err").read()[-2500:])
PY
