#!/bin/bash
# ncu captures of the default bench workload (1 GPU, under gpurun).  Usage: tools/profile_r2.sh <tag>
#  a. launch list of the detection step (share of device time per kernel)
#  b. every tcgen05 conv launch of ONE step with --set full -> per-launch metrics + DRAM traffic of the step
#  c. launch list of one training step
TAG=${1:-r2}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-nms-sweep"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 105 -c 21 \
    -f -o gpurun_out/prof_convstep_${TAG} $BENCH > gpurun_out/ncu_convstep_${TAG}.log 2>&1
echo "conv step full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 500 --csv \
    --log-file gpurun_out/launches_train_${TAG}.csv python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_${TAG}.log 2>&1
echo "train launch list rc=$?"
ls -la gpurun_out/*${TAG}*
