#!/bin/bash
# run 28: nn mode 4 timeline + epilogue knock-out experiments
set -x
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --cpu-sample-pairs 0 --corr-mode 2 --nn-mode 4"
ROREG_DEBUG_NN_TRACE=gpurun_out/r28_nn4_trace.txt timeout 300 $B > gpurun_out/r28_bench_trace.json 2> gpurun_out/r28_bench_trace.err
for sk in 1 2 4 7; do
ROREG_DEBUG_NN_SKIP=$sk timeout 300 $B > gpurun_out/r28_bench_skip$sk.json 2> gpurun_out/r28_bench_skip$sk.err
done
ROREG_DEBUG_NN_SKIP=7 ROREG_DEBUG_NN_PASSES=1 timeout 300 $B > gpurun_out/r28_bench_skip7_p1.json 2> gpurun_out/r28_bench_skip7_p1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r28_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
sed -n 1,12p gpurun_out/r28_nn4_trace.txt; sed -n 100,130p gpurun_out/r28_nn4_trace.txt
