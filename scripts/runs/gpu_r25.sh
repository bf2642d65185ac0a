#!/bin/bash
# run 25: corr mode 2 (1-match items, 2 CTAs/SM): parity + timing vs mode 1
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "corr or des2r or register_batch" > gpurun_out/r25_pytest_corr.txt 2>&1
tail -5 gpurun_out/r25_pytest_corr.txt
B="python bench.py --steps 30 --warmup 3 --cpu-sample-pairs 0"
timeout 300 $B --corr-mode 2 > gpurun_out/r25_bench_c2.json 2> gpurun_out/r25_bench_c2.err
ROREG_DEBUG_CORR_CTAS=1 timeout 300 $B --corr-mode 2 > gpurun_out/r25_bench_c2_1cta.json 2> gpurun_out/r25_bench_c2_1cta.err
ROREG_DEBUG_CORR_SKIP=3 ROREG_DEBUG_CORR_PASSES=1 timeout 300 $B --corr-mode 2 > gpurun_out/r25_bench_c2_skel.json 2> gpurun_out/r25_bench_c2_skel.err
ROREG_DEBUG_CORR_SKIP=1 timeout 300 $B --corr-mode 2 > gpurun_out/r25_bench_c2_nodiag.json 2> gpurun_out/r25_bench_c2_nodiag.err
ROREG_DEBUG_CORR_SKIP=2 timeout 300 $B --corr-mode 2 > gpurun_out/r25_bench_c2_noconv.json 2> gpurun_out/r25_bench_c2_noconv.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r25_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
