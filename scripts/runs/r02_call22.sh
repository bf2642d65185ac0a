#!/bin/bash
# Round 2, GPU call 22: mmap + cudaHostRegister loader experiment; bench with the new GEMM
set -x
mkdir -p gpurun_out
nproc
timeout 300 python scripts/hostreg_test.py 2>&1 | tee gpurun_out/c22_hostreg.txt
timeout 900 python bench.py > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err; tail -3 gpurun_out/c22_bench.err
