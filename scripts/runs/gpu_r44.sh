#!/bin/bash
# run 44: pooling fused with the nn-mode-4 operand image; nn4 with PRMT tile index + FMNMX3 row groups
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or mode4 or register_batch" > gpurun_out/r44_pytest.txt 2>&1
tail -5 gpurun_out/r44_pytest.txt
B="python bench.py --steps 20 --warmup 3 --cpu-sample-pairs 0 --corr-mode 3 --nn-mode 4"
timeout 200 $B > gpurun_out/r44_bench.json 2> gpurun_out/r44_bench.err
tail -3 gpurun_out/r44_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r44_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
