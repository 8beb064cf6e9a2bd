#!/bin/bash
# ncu capture of every HBM-bound kernel (1 GPU, under gpurun): duration + DRAM bytes per launch, everything that is not a
# tcgen05 GEMM, in one detection step, one training step + Adam and the NMS sweep points (tools/hbm_kernels.py).
# Usage: tools/profile_hbm.sh <tag>   -> gpurun_out/hbm_<tag>.csv ; summarise here with python tools/hbm_summary.py <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__grid_size,launch__block_size
timeout 1200 ncu --profile-from-start off --metrics $M --clock-control none -k 'regex:^(?!.*(conv_tc_kernel|wgrad_tc_kernel))' \
    --csv --log-file gpurun_out/hbm_${TAG}.csv python tools/hbm_kernels.py $2 > gpurun_out/ncu_hbm_${TAG}.log 2>&1
echo "hbm kernels ncu rc=$?"; tail -3 gpurun_out/ncu_hbm_${TAG}.log; wc -l gpurun_out/hbm_${TAG}.csv
