#!/bin/bash
# ncu --set full on conv0 (CUDA-core kernel and tensor-core kernel), one launch each.  Usage: tools/ncu_conv0.sh <tag>
TAG=${1:-r1h}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv0_pool_kernel|conv0_tc_pool_kernel' -s 4 -c 1 \
    -f -o gpurun_out/prof_conv0_simt_${TAG} python bench.py --conv0-tc 0 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv0_simt_${TAG}.log 2>&1
echo "simt rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv0_pool_kernel|conv0_tc_pool_kernel' -s 4 -c 1 \
    -f -o gpurun_out/prof_conv0_tc_${TAG} python bench.py --conv0-tc 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv0_tc_${TAG}.log 2>&1
echo "tc rc=$?"
ls -la gpurun_out/prof_conv0_*${TAG}*
