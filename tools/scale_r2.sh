#!/bin/bash
# One visit of an 8-GPU box: training step (BASELINE configs[2]) and 608x608 inference (configs[3]: 32 images per GPU) at
# 1 / 2 / 4 / 8 GPUs, one rank per GPU over NCCL.  Usage: gpurun --gpus 8 -- tools/scale_r2.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
run() {  # n, outfile, bench args...
  local n=$1 out=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > $out 2> ${out%.json}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > $out 2> ${out%.json}.err; fi
  echo "n=$n $* rc=$?"; python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$out") if l.startswith("{")][-1]
    print("   value %.1f e2e %.1f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("train", {}).get("allreduce_ms_alone"), d["clocks"]["sm_mhz"])
except Exception as e:
    print("   unreadable:", e)
PY
}
if [ "$2" = "infer" ]; then
  for n in 1 2 4 8; do run $n gpurun_out/scale_infer_${TAG}_n$n.json --steps 200 --no-cpu-baseline --no-nms-sweep; done
else
  for n in 1 2 4 8; do run $n gpurun_out/scale_train_${TAG}_n$n.json --workload train --steps 20 --no-cpu-baseline; done
  for n in 8; do run $n gpurun_out/scale_608_${TAG}_n$n.json --size 608 --steps 60 --no-cpu-baseline --no-nms-sweep; done
fi
