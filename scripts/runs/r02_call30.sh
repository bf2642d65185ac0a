#!/bin/bash
# Round 2, GPU call 30: epilogue store staging (this build) against the previous build (direct thread-per-row stores); ET chunk 16000
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_matchot.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/c30_pytest.txt
echo "## staged stores (this build)" | tee gpurun_out/c30_nets.txt
timeout 600 python scripts/time_nets.py 2>&1 | tee -a gpurun_out/c30_nets.txt
echo "## direct stores (previous build)" | tee -a gpurun_out/c30_nets.txt
ROREG_B200_LIB=$PWD/roreg_b200/csrc/libroreg_b200_prev.so timeout 600 python scripts/time_nets.py 2>&1 | grep "^GF\|^ET\|^RD\|^Match" | tee -a gpurun_out/c30_nets.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c30_gf_et_launches.csv python scripts/gf_et_once.py 1 > /dev/null 2>&1
