#!/bin/bash
# Round 2, GPU call 13 (--gpus 8): weak scaling of the default bench at N = 2 and 8 (value, scene e2e, per-pair-upload e2e) + reference arm under torchrun.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c13_topo.txt 2>&1; nproc; free -g | head -2; df -h /dev/shm | tail -1
for n in 2 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/c13_bench_g$n.json 2> gpurun_out/c13_bench_g$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c13_bench_g$n.json").read().strip().splitlines()[-1])
    print("N=$n value", round(d["value"]), "e2e(scene)", round(d["e2e"]["value"]), d["e2e"].get("seconds"), "h2d GB/s/rank", round(d["e2e"].get("h2d_gb_per_s_per_rank",0),1), "pair_upload", round(d["e2e_pair_upload"]["value"]), "reuse4", round(d["e2e_cloud_reuse4"]["value"]), d["clocks"])
except Exception as e:
    print("N=$n FAILED", e); print(open("gpurun_out/c13_bench_g$n.err").read()[-2000:])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/c13_ref_g2.json 2> gpurun_out/c13_ref_g2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c13_ref_g2.json").read().strip().splitlines()[-1]); print("reference arm under torchrun:", d["value"], "cores", d["cpu_baseline"]["cores"])
except Exception as e:
    print("ref FAILED", e); print(open("gpurun_out/c13_ref_g2.err").read()[-1500:])
PY
