#!/bin/bash
# run 41: corr3 with two epilogue sets
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "corr or des2r" > gpurun_out/r41_pytest.txt 2>&1
tail -5 gpurun_out/r41_pytest.txt
B="python bench.py --steps 20 --warmup 3 --cpu-sample-pairs 0 --corr-mode 3 --nn-mode 4"
timeout 200 $B > gpurun_out/r41_bench.json 2> gpurun_out/r41_bench.err
ROREG_DEBUG_CORR_TRACE=gpurun_out/r41_corr3_trace.txt timeout 200 $B > gpurun_out/r41_bench_trace.json 2> gpurun_out/r41_bench_trace.err
tail -3 gpurun_out/r41_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r41_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
sed -n 1,2p gpurun_out/r41_corr3_trace.txt; sed -n 60,74p gpurun_out/r41_corr3_trace.txt
