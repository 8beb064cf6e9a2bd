"""Host-side cost of one bench step (Python + ctypes enqueue time, no device sync inside the loop) and a cProfile
breakdown -- the e2e loop is host-driven, so this bounds the end-to-end rate.  Diagnostic tool."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from yolo_tf_b200 import variables  # noqa: E402
from yolo_tf_b200.model.yolo2 import Builder  # noqa: E402
from yolo_tf_b200.utils.postprocess import non_max_suppress_device  # noqa: E402

B, size, C = 32, 416, 80
params = bench.synthetic_checkpoint(C, 5)
store = variables.reset_default_store()
store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
builder = Builder.from_values([str(i) for i in range(C)], size, size, bench.ANCHORS_COCO)
x = torch.from_numpy(np.random.RandomState(1).normal(0, 1, size=(B, size, size, 3)).astype(np.float32)).cuda()
N = (size // 32) ** 2 * 5


def step():
    builder(x)
    m = builder.model
    non_max_suppress_device(m.conf.view(B, N, C), m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2), 0.3, 0.4, check=False)


for _ in range(20):
    step()
torch.cuda.synchronize()
# enqueue-only time: the GPU queue is deep enough for 40 steps (26 launches each)
t0 = time.perf_counter()
for _ in range(40):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue per step: %.3f ms; incl. drain: %.3f ms/step" % ((t1 - t0) / 40 * 1e3, (t2 - t0) / 40 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
