#!/bin/bash
# Round 2, GPU call 29: per-kernel times of one GF + one ET forward (ncu launch list, second pass)
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c29_gf_et_launches.csv python scripts/gf_et_once.py 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c29_gf_et_launches_3pass.csv python scripts/gf_et_once.py 3 > /dev/null 2>&1
