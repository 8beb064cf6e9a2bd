#!/bin/bash
# compute-sanitizer passes (B200, under gpurun): memcheck, racecheck and synccheck on the detection step (smoke(): backbone +
# decode + NMS at 64x64, B=2) and on one training step + Adam (tools/sanitize_step.py: Darknet-19 and tiny at 64x64, B=2).
# The kernels' hand-rolled mbarrier / flag protocols are what racecheck and synccheck are for.
# Usage: tools/sanitize.sh <tag>  ->  gpurun_out/sanitizer_<tag>.txt
TAG=${1:-r2}
OUT=gpurun_out/sanitizer_${TAG}.txt
mkdir -p gpurun_out
echo "# compute-sanitizer (B200, under gpurun), tag ${TAG}" > $OUT
for tool in memcheck racecheck synccheck; do
  for target in "python __graft_entry__.py smoke" "python tools/sanitize_step.py"; do
    echo "## --tool $tool : $target" >> $OUT
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 $target > gpurun_out/san_tmp.log 2>&1
    echo "rc=$?" >> $OUT
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|step ok|Error|hazard" gpurun_out/san_tmp.log | head -12 >> $OUT
  done
done
cat $OUT
