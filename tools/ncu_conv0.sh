#!/bin/bash
# ncu --set full on conv0, one launch.  Usage: tools/ncu_conv0.sh <tag> [conv0_tc mode: 0 = CUDA cores, 1 = tensor cores, 2 = + FAST gather]
TAG=${1:-r1j}
MODE=${2:-2}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv0_pool_kernel|conv0_tc_pool_kernel' -s 4 -c 1 \
    -f -o gpurun_out/prof_conv0_m${MODE}_${TAG} python bench.py --conv0-tc $MODE --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv0_m${MODE}_${TAG}.log 2>&1
echo "rc=$?"
ls -la gpurun_out/prof_conv0_m${MODE}_${TAG}*
