#!/bin/bash
# Round 2, GPU call 23: ncu source view of the reworked gather GEMM (L_in: 13 k-chunks per tile = epilogue-bound; L_a: 104)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -o gpurun_out/c23_gemm python scripts/gf_one_chunk.py 1 > gpurun_out/c23_ncu.log 2>&1
ncu -i gpurun_out/c23_gemm.ncu-rep --page source --csv --print-source sass > gpurun_out/c23_gemm_source.csv 2>/dev/null
ncu -i gpurun_out/c23_gemm.ncu-rep --page raw --csv > gpurun_out/c23_gemm_raw.csv 2>/dev/null
rm -f gpurun_out/c23_gemm.ncu-rep
