#!/bin/bash
# run 32: ncu launch list of the nn-mode-4 / corr-mode-2 step
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r32_launches.csv python bench.py --steps 3 --warmup 3 --pairs-per-step 32 --cpu-sample-pairs 0 --corr-mode 2 --nn-mode 4 > gpurun_out/r32_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r32_launches.csv')) if len(r)>5]
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
