#!/bin/bash
# First gpurun call of the next round (prepared at the end of round 1, when the GPU budget was spent; NOT run yet).
# 1. confirm what was written without a GPU: score mode 1 (float32 pre-filter, kernels_ransac.cuh), the rewritten plugin mirrors,
#    the RD/RM drop-in test and the scene driver (roreg_b200/scene.py);
# 2. A/B the bench with and without it, and with the pipelined entry point (roreg_register_batch_pipelined); 3. multi-rank e2e with / without the NUMA binding of bench.py.
set -x
mkdir -p gpurun_out
ROREG_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "score_mode1 or ransac or pipelined" > gpurun_out/n01_pytest_experimental.txt 2>&1
tail -4 gpurun_out/n01_pytest_experimental.txt
ROREG_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_nets.py tests/test_gpu_matchot.py -x -q -m gpu > gpurun_out/n01_pytest_plugins.txt 2>&1
tail -4 gpurun_out/n01_pytest_plugins.txt
ROREG_TEST_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "score_mode1 or pipelined" > gpurun_out/n01_sanitizer.txt 2>&1
tail -3 gpurun_out/n01_sanitizer.txt
for m in 0 1; do
  timeout 400 python bench.py --score-mode $m --cpu-sample-pairs 0 --value-only 1 > gpurun_out/n01_bench_score$m.json 2> gpurun_out/n01_bench_score$m.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/n01_bench_score$m.json").read().strip().splitlines()[-1])
print("score_mode $m:", round(d["value"]), "pairs/s", {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d["pose_check"])
PY
done
for cfg in "--pipelined 1" "--pipelined 1 --score-mode 1" "--pipelined 1 --score-mode 1 --pairs-per-step 128" "--pairs-per-step 128"; do
  tag=$(echo $cfg | tr -d ' -')
  timeout 400 python bench.py $cfg --cpu-sample-pairs 0 --value-only 1 > gpurun_out/n01_bench_$tag.json 2> gpurun_out/n01_bench_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/n01_bench_$tag.json").read().strip().splitlines()[-1])
print("$cfg:", round(d["value"]), "pairs/s", d["pose_check"])
PY
done
# which phase wins the freed CTA slots while T(i-1) and P(i) co-run
for prio in lo same; do
  ROREG_PIPE_TAIL_PRIO=$prio timeout 400 python bench.py --pipelined 1 --cpu-sample-pairs 0 --value-only 1 > gpurun_out/n01_bench_pipe_$prio.json 2> gpurun_out/n01_bench_pipe_$prio.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/n01_bench_pipe_$prio.json").read().strip().splitlines()[-1])
print("pipelined, tail priority $prio:", round(d["value"]), "pairs/s", d["pose_check"])
PY
done
# pipelined + score mode 1 with a cap on the scoring kernel's resident CTAs (room for the co-running pooling kernel)
for cap in 1 2 4; do
  for prio in hi same; do
    ROREG_SCORE_CTAS_PER_SM=$cap ROREG_PIPE_TAIL_PRIO=$prio timeout 400 python bench.py --pipelined 1 --score-mode 1 --cpu-sample-pairs 0 --value-only 1 > gpurun_out/n01_bench_pipe_cap${cap}_$prio.json 2> gpurun_out/n01_bench_pipe_cap${cap}_$prio.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/n01_bench_pipe_cap${cap}_$prio.json").read().strip().splitlines()[-1])
print("pipelined, score mode 1, $cap scoring CTAs per SM, tail priority $prio:", round(d["value"]), "pairs/s", d["pose_check"])
PY
  done
done
# Registration-Recall parity on synthetic pairs of graded difficulty (CUDA engine vs reference arithmetic on the host)
timeout 900 python scripts/rr_parity.py --pairs 200 --n 2000 > gpurun_out/n01_rr_parity.json 2> gpurun_out/n01_rr_parity.err; cat gpurun_out/n01_rr_parity.json
