#!/bin/bash
# Round 2, GPU call 31: ncu source view of GF Conv_in (13 k-chunks per tile: epilogue-bound) in the staged-store build
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 10 -c 1 -o gpurun_out/c31_gemm python scripts/gf_et_once.py 1 > gpurun_out/c31_ncu.log 2>&1
ncu -i gpurun_out/c31_gemm.ncu-rep --page source --csv --print-source sass > gpurun_out/c31_gemm_source.csv 2>/dev/null
ncu -i gpurun_out/c31_gemm.ncu-rep --page raw --csv > gpurun_out/c31_gemm_raw.csv 2>/dev/null
rm -f gpurun_out/c31_gemm.ncu-rep
