#!/bin/bash
# run 23: TMEM read microbenchmark + corr TC on 1-D bulk copies (parity, timings with 4/8 convert warps, bare skeleton)
set -x
mkdir -p gpurun_out
timeout 120 ./scripts/tmem_bench > gpurun_out/r23_tmem_bench.txt 2>&1
cat gpurun_out/r23_tmem_bench.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "corr or des2r or register_batch" > gpurun_out/r23_pytest_corr.txt 2>&1
tail -5 gpurun_out/r23_pytest_corr.txt
B="python bench.py --steps 30 --warmup 3 --cpu-sample-pairs 0"
timeout 300 $B > gpurun_out/r23_bench_cw4.json 2> gpurun_out/r23_bench_cw4.err
ROREG_DEBUG_CORR_CW=8 timeout 300 $B > gpurun_out/r23_bench_cw8.json 2> gpurun_out/r23_bench_cw8.err
ROREG_DEBUG_CORR_SKIP=3 ROREG_DEBUG_CORR_PASSES=1 timeout 300 $B > gpurun_out/r23_bench_skel.json 2> gpurun_out/r23_bench_skel.err
ROREG_DEBUG_CORR_CW=8 ROREG_DEBUG_CORR_SKIP=1 timeout 300 $B > gpurun_out/r23_bench_cw8_nodiag.json 2> gpurun_out/r23_bench_cw8_nodiag.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r23_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
