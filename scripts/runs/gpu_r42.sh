#!/bin/bash
# run 42: ncu --set full of the four dominant kernels of the step (nn mode 4, corr mode 3)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nn_tc4_kernel|group_corr_tc3_kernel|inv_pool_kernel|ransac_score_kernel' -s 12 -c 4 -o gpurun_out/r42_prof -f python bench.py --steps 2 --warmup 3 --pairs-per-step 32 --cpu-sample-pairs 0 --corr-mode 3 --nn-mode 4 > gpurun_out/r42_ncu.log 2>&1
tail -5 gpurun_out/r42_ncu.log
ls -la gpurun_out/r42_prof.ncu-rep
