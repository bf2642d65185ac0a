#!/bin/bash
# run 43: nn4 with FMA-pipe tile index + corr3 with the half-warp-per-row loader mapping
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or mode4 or corr or des2r" > gpurun_out/r43_pytest.txt 2>&1
tail -5 gpurun_out/r43_pytest.txt
B="python bench.py --steps 20 --warmup 3 --cpu-sample-pairs 0 --corr-mode 3 --nn-mode 4"
timeout 200 $B > gpurun_out/r43_bench.json 2> gpurun_out/r43_bench.err
tail -3 gpurun_out/r43_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r43_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'pairs/s', {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['pose_check'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file gpurun_out/r43_launches.csv python bench.py --steps 3 --warmup 3 --pairs-per-step 32 --cpu-sample-pairs 0 --corr-mode 3 --nn-mode 4 > gpurun_out/r43_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r43_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1000
    elif r[ui]=='ms': v*=1000
    k=r[ki][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:62s} {n:4d} {t:10.1f} us {100*t/tot:5.1f}% avg {t/n:8.1f}")
PY
