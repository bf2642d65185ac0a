#!/bin/bash
# Round 2, GPU call 2: conflict-free coset layout of group_corr_tc3_kernel (icosa_cosets.cuh): parity, bench, ncu of the kernel.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_matchot.py -x -q -m gpu -s > gpurun_out/c2_pytest.txt 2>&1
tail -5 gpurun_out/c2_pytest.txt; grep -n "parity" gpurun_out/c2_pytest.txt | head
run_bench() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py "$@" --cpu-sample-pairs 0 --value-only 1 > gpurun_out/c2_bench_$tag.json 2> gpurun_out/c2_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c2_bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag:", round(d["value"]), "pairs/s", {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d.get("pose_check"))
except Exception as e:
    print("$tag: FAILED", e)
PY
}
run_bench base
ROREG_SCORE_CTAS_PER_SM=2 run_bench pipe_score1_cap2 --pipelined 1 --score-mode 1
ROREG_SCORE_CTAS_PER_SM=1 run_bench pipe_score1_cap1 --pipelined 1 --score-mode 1
ROREG_SCORE_CTAS_PER_SM=3 run_bench pipe_score1_cap3 --pipelined 1 --score-mode 1
# ncu: full capture of the Des2R kernel only (one launch), then the launch list of one bench step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_corr_tc3 -c 1 -o gpurun_out/c2_corr3 python bench.py --steps 2 --warmup 1 --cpu-sample-pairs 0 --value-only 1 > gpurun_out/c2_ncu.log 2>&1
tail -3 gpurun_out/c2_ncu.log
ncu -i gpurun_out/c2_corr3.ncu-rep --page raw --csv > gpurun_out/c2_corr3_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/c2_corr3_raw.csv")))
hdr=rows[0]; vals=rows[-1]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","l1tex__data_pipe_lsu_wavefronts.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","sm__inst_executed_pipe_lsu.sum","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
for h,v in zip(hdr,vals):
    if any(w in h for w in want): print(h,v)
PY
