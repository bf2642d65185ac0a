#!/bin/bash
# Round 2, GPU call 4: corr3 pair-wise item walk + optional third epilogue set; scene-level e2e in bench.py.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_matchot.py -x -q -m gpu > gpurun_out/c4_pytest.txt 2>&1
tail -4 gpurun_out/c4_pytest.txt
ROREG_CORR_SETS=3 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "corr or des2r or full_size or register_batch" > gpurun_out/c4_pytest_sets3.txt 2>&1
tail -4 gpurun_out/c4_pytest_sets3.txt
run_bench() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py "$@" --cpu-sample-pairs 0 --value-only 1 > gpurun_out/c4_bench_$tag.json 2> gpurun_out/c4_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c4_bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag:", round(d["value"]), "pairs/s", {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d.get("pose_check"))
except Exception as e:
    print("$tag: FAILED", e); print(open("gpurun_out/c4_bench_$tag.err").read()[-1500:])
PY
}
run_bench base
ROREG_CORR_SETS=3 run_bench sets3
ROREG_SCORE_CTAS_PER_SM=2 run_bench pipe_score1_cap2 --pipelined 1 --score-mode 1
ROREG_CORR_SETS=3 ROREG_SCORE_CTAS_PER_SM=2 run_bench sets3_pipe_score1_cap2 --pipelined 1 --score-mode 1
timeout 900 python bench.py > gpurun_out/c4_bench_full.json 2> gpurun_out/c4_bench_full.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c4_bench_full.json").read().strip().splitlines()[-1])
    print("full:", round(d["value"]), "e2e", d["e2e"], "pair_upload", round(d["e2e_pair_upload"]["value"]), "reuse4", round(d["e2e_cloud_reuse4"]["value"]), d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("full FAILED", e); print(open("gpurun_out/c4_bench_full.err").read()[-2500:])
PY
