#!/bin/bash
# run 30: nn mode 4 v2 (fp16 two-accumulator, deferred row index): parity, timing, timeline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_mode or mode4 or tensor_core_nn" > gpurun_out/r30_pytest_nn4.txt 2>&1
tail -15 gpurun_out/r30_pytest_nn4.txt
B="python bench.py --steps 20 --warmup 3 --cpu-sample-pairs 0 --corr-mode 2 --nn-mode 4"
timeout 300 $B > gpurun_out/r30_bench_nn4.json 2> gpurun_out/r30_bench_nn4.err
ROREG_DEBUG_NN_TRACE=gpurun_out/r30_nn4_trace.txt timeout 300 $B > gpurun_out/r30_bench_trace.json 2> gpurun_out/r30_bench_trace.err


tail -3 gpurun_out/r30_bench_nn4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r30_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
sed -n 1,6p gpurun_out/r30_nn4_trace.txt; sed -n 100,112p gpurun_out/r30_nn4_trace.txt
