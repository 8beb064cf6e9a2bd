#!/bin/bash
# Second-pass ncu captures (1 GPU, under gpurun).  Usage: tools/profile2.sh <tag>
#  a. every tcgen05 conv launch of ONE inference step with --set full  -> DRAM traffic per step (roofline.traffic)
#  b. launch list of one TRAINING step (B=64)
#  c. --set full on the NMS kernels in the sweep configuration (B=512, N=845, K=1000)
TAG=${1:-r1b}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 63 -c 21 \
    -f -o gpurun_out/prof_convstep_${TAG} $BENCH > gpurun_out/ncu_convstep_${TAG}.log 2>&1
echo "conv step full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 400 --csv \
    --log-file gpurun_out/launches_train_${TAG}.csv python tools/bench_train.py --steps 1 --warmup 1 > gpurun_out/ncu_train_${TAG}.log 2>&1
echo "train launch list rc=$?"
cat > /tmp/nms_one.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
sys.argv = ["x", "prof"]
import tools.bench_nms as bn
from yolo_tf_b200 import _lib
L = _lib.lib()
rs = np.random.RandomState(5)
conf, lo, hi = bn.make_inputs(rs, 512, 13, 80, 1000)
d0, dlo, dhi = torch.from_numpy(conf).cuda(), torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
nb = L.y2_nms_workspace_bytes(512, 845, 80)
ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
for _ in range(3):
    w = d0.clone()
    _lib.check(L.y2_nms(_lib.ptr(w), _lib.ptr(dlo), _lib.ptr(dhi), 512, 845, 80, 0.3, 0.4, None, None, _lib.ptr(ws), nb, None))
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nms_select|nms_apply' -s 4 -c 2 \
    -f -o gpurun_out/prof_nms_${TAG} python /tmp/nms_one.py > gpurun_out/ncu_nms_${TAG}.log 2>&1
echo "nms full rc=$?"
ls -la gpurun_out/*${TAG}*
