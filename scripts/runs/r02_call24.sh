#!/bin/bash
# Round 2, GPU call 24: GEMM epilogue fast path: correctness, timeline, timings
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nets.py tests/test_gpu_matchot.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/c24_pytest.txt
for l in 0 1 2; do
  ROREG_B200_LIB=$PWD/roreg_b200/csrc/libroreg_b200_trace.so ROREG_DEBUG_GEMM_TRACE=gpurun_out/c24_trace_$l.txt ROREG_DEBUG_GEMM_TRACE_LAUNCH=$l timeout 300 python scripts/gf_one_chunk.py 1 > /dev/null 2>&1
done
timeout 600 python scripts/time_nets.py 2>&1 | tee gpurun_out/c24_nets_timing.txt
