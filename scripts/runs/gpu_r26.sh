#!/bin/bash
# run 26: corr mode-2 timeline trace; nn mode 4 parity + timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or mode4" > gpurun_out/r26_pytest_nn4.txt 2>&1
tail -15 gpurun_out/r26_pytest_nn4.txt
B="python bench.py --steps 30 --warmup 3 --cpu-sample-pairs 0"
ROREG_DEBUG_CORR_TRACE=gpurun_out/r26_corr2_trace.txt timeout 300 $B --corr-mode 2 > gpurun_out/r26_bench_trace.json 2> gpurun_out/r26_bench_trace.err
timeout 300 $B --corr-mode 2 --nn-mode 4 > gpurun_out/r26_bench_nn4.json 2> gpurun_out/r26_bench_nn4.err
ROREG_DEBUG_NN_PASSES=1 timeout 300 $B --corr-mode 2 --nn-mode 4 > gpurun_out/r26_bench_nn4_p1.json 2> gpurun_out/r26_bench_nn4_p1.err
tail -3 gpurun_out/r26_bench_nn4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r26_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
head -40 gpurun_out/r26_corr2_trace.txt
