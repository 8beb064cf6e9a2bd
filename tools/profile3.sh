#!/bin/bash
# Third-pass ncu captures (1 GPU, under gpurun) for the CTA-pair + chain-cap kernel.  Usage: tools/profile3.sh <tag>
#  a. launch list of the inference step (share of device time per kernel)
#  b. every tcgen05 conv launch of ONE inference step with --set full  -> per-launch metrics + DRAM traffic per step
#  c. launch list of one TRAINING step (B=64)
TAG=${1:-r1g}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 260 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 63 -c 21 \
    -f -o gpurun_out/prof_convstep_${TAG} $BENCH > gpurun_out/ncu_convstep_${TAG}.log 2>&1
echo "conv step full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 400 --csv \
    --log-file gpurun_out/launches_train_${TAG}.csv python tools/bench_train.py --steps 1 --warmup 1 > gpurun_out/ncu_train_${TAG}.log 2>&1
echo "train launch list rc=$?"
ls -la gpurun_out/*${TAG}*
