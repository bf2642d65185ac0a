#!/bin/bash
# run 48: deterministic bucket sort (device draws reproducible), overlap optional (default off): tests + bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_batch or fast_path or device_draws or ransac" > gpurun_out/r48_pytest.txt 2>&1
tail -5 gpurun_out/r48_pytest.txt
timeout 400 python bench.py --steps 30 --cpu-sample-pairs 0 > gpurun_out/r48_bench.json 2> gpurun_out/r48_bench.err
tail -3 gpurun_out/r48_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r48_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['value']), 'pairs/s serial', round(r['serial_schedule_pairs_per_s']), 'e2e', round(d['e2e']['value']), 'scene', round(d['e2e_scene']['value']), {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, round(r['fused_step']['hbm_frac_overlapped'],3), d['pose_check'], d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
