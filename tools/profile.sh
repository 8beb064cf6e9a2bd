#!/bin/bash
# ncu captures for profiles/ (run under gpurun, 1 GPU).  Usage: tools/profile.sh <round-tag>
# 1. launch list with device times (cold-cache, serialised: compare SHARES, not absolutes)
# 2. --set full on the dominant tcgen05 conv launches (conv20, final, conv1 of the next step)
# 3. --set full on the HBM-bound kernels (conv0, pool, reorg, decode, NMS)
TAG=${1:-r1}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 260 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 19 -c 3 \
    -f -o gpurun_out/prof_conv_${TAG} $BENCH > gpurun_out/ncu_conv_${TAG}.log 2>&1
echo "conv full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'conv0_pool|maxpool_planes|reorg_kernel|decode_kernel|nms_select|nms_apply|splitk_finish' -s 9 -c 10 \
    -f -o gpurun_out/prof_hbm_${TAG} $BENCH > gpurun_out/ncu_hbm_${TAG}.log 2>&1
echo "hbm full rc=$?"
ls -la gpurun_out/*.ncu-rep
