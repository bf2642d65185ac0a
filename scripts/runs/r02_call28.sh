#!/bin/bash
# Round 2, GPU call 28: staged Sinkhorn kernel + producer-equality test of the gather GEMM
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matchot.py tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/c28_pytest.txt
timeout 300 python scripts/time_sinkhorn.py 2>&1 | tail -4 | tee gpurun_out/c28_sinkhorn.txt
timeout 300 python scripts/time_match_ot.py 2>&1 | tail -3 | tee -a gpurun_out/c28_sinkhorn.txt
