#!/bin/bash
# run 53: nn4 with the redux-based column reduce; new tiny/degenerate test
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_mode or mode4 or tensor_core_nn or fast_path or fast_modes or overlapped" > gpurun_out/r53_pytest.txt 2>&1
tail -5 gpurun_out/r53_pytest.txt
timeout 400 python bench.py --steps 30 --cpu-sample-pairs 0 > gpurun_out/r53_bench.json 2> gpurun_out/r53_bench.err
ROREG_DEBUG_NN_TRACE=gpurun_out/r53_nn4_trace.txt timeout 400 python bench.py --steps 5 --cpu-sample-pairs 0 > gpurun_out/r53_bench_trace.json 2> gpurun_out/r53_bench_trace.err
tail -3 gpurun_out/r53_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r53_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, round(r['fused_step']['hbm_frac'],3), d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
sed -n 30,46p gpurun_out/r53_nn4_trace.txt
