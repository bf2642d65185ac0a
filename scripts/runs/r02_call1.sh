#!/bin/bash
# Round 2, GPU call 1: everything written without a GPU at the end of round 1 (score mode 1, pipelined engine, RD/RM drop-in,
# scene driver) now un-gated + the new full-size oracle test + the teacher-forced Match_ot parity; A/B bench lines; RR parity.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/c1_pytest.txt 2>&1
tail -15 gpurun_out/c1_pytest.txt; grep -n "parity\|Match_ot" gpurun_out/c1_pytest.txt | head
# if -x stopped early, run the remaining files anyway so one failure does not hide the others
for f in tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_nets.py tests/test_gpu_matchot.py; do
  timeout 900 python -m pytest $f -q -m gpu -s > gpurun_out/c1_pytest_$(basename $f .py).txt 2>&1; tail -3 gpurun_out/c1_pytest_$(basename $f .py).txt
done
run_bench() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py "$@" --cpu-sample-pairs 0 --value-only 1 > gpurun_out/c1_bench_$tag.json 2> gpurun_out/c1_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c1_bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag:", round(d["value"]), "pairs/s", {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d.get("pose_check"))
except Exception as e:
    print("$tag: FAILED", e)
PY
}
run_bench score0 --score-mode 0
run_bench score1 --score-mode 1
run_bench pipe --pipelined 1
run_bench pipe_score1 --pipelined 1 --score-mode 1
run_bench pipe_score1_128 --pipelined 1 --score-mode 1 --pairs-per-step 128
run_bench b128 --pairs-per-step 128
run_bench b128_score1 --pairs-per-step 128 --score-mode 1
ROREG_PIPE_TAIL_PRIO=lo run_bench pipe_score1_lo --pipelined 1 --score-mode 1
ROREG_SCORE_CTAS_PER_SM=2 run_bench pipe_score1_cap2 --pipelined 1 --score-mode 1
timeout 900 python scripts/rr_parity.py --pairs 200 --n 2000 > gpurun_out/c1_rr_parity.json 2> gpurun_out/c1_rr_parity.err; cat gpurun_out/c1_rr_parity.json; tail -3 gpurun_out/c1_rr_parity.err
