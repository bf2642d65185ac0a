#!/bin/bash
# run 40: shared-space pointer fix (STS/LDS everywhere): full GPU suite, bench, corr3 timeline
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r40_pytest.txt 2>&1
tail -5 gpurun_out/r40_pytest.txt
B="python bench.py --steps 20 --warmup 3 --cpu-sample-pairs 0 --corr-mode 3 --nn-mode 4"
timeout 200 $B > gpurun_out/r40_bench.json 2> gpurun_out/r40_bench.err
ROREG_DEBUG_CORR_TRACE=gpurun_out/r40_corr3_trace.txt ROREG_DEBUG_NN_TRACE=gpurun_out/r40_nn4_trace.txt timeout 200 $B > gpurun_out/r40_bench_trace.json 2> gpurun_out/r40_bench_trace.err
timeout 200 python bench.py --steps 20 --warmup 3 --cpu-sample-pairs 0 --corr-mode 1 --nn-mode 2 > gpurun_out/r40_bench_old.json 2> gpurun_out/r40_bench_old.err
tail -3 gpurun_out/r40_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r40_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
sed -n 1,2p gpurun_out/r40_corr3_trace.txt; sed -n 60,72p gpurun_out/r40_corr3_trace.txt
sed -n 100,108p gpurun_out/r40_nn4_trace.txt
