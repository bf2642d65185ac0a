#!/bin/bash
# Round 2, GPU call 16: fused all-pairs kernel (one launch, rotation loop inside): parity test, timing against the 60-launch path.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/c16_pytest_nets.txt 2>&1; tail -4 gpurun_out/c16_pytest_nets.txt
timeout 300 python scripts/ab_allpairs_r01.py now > gpurun_out/c16_allpairs_fused.txt 2>&1; tail -2 gpurun_out/c16_allpairs_fused.txt
ROREG_ALLPAIRS_LAUNCHES=1 timeout 300 python scripts/ab_allpairs_r01.py now > gpurun_out/c16_allpairs_60launch.txt 2>&1; tail -2 gpurun_out/c16_allpairs_60launch.txt
timeout 300 python scripts/time_allpairs.py > gpurun_out/c16_time_allpairs.txt 2>&1; tail -2 gpurun_out/c16_time_allpairs.txt
timeout 600 ncu --set full --clock-control none -k regex:allpairs_tc_kernel -c 1 -o gpurun_out/c16_allpairs python scripts/ab_allpairs_r01.py now > gpurun_out/c16_ncu.log 2>&1; tail -1 gpurun_out/c16_ncu.log
ncu -i gpurun_out/c16_allpairs.ncu-rep --page raw --csv > gpurun_out/c16_allpairs_raw.csv 2>/dev/null; rm -f gpurun_out/c16_allpairs.ncu-rep
