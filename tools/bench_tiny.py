"""Tiny YOLOv2 (model/yolo2/inference.py:25-50) on one B200: images/s of the backbone + decode + NMS at batch 32, 416x416, 20
classes (config/yolo2/tiny-20.ini), device-resident inputs rotated over 4 batches, CUDA events on the launching stream.
    python tools/bench_tiny.py [--batch 32] [--size 416] [--classes 20] [--steps 100] [--warmup 5]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from oracle import head_oracle as ho                      # anchors table only
    from oracle.darknet_oracle import flops_per_image, init_params, tiny_layer_table
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder
    from yolo_tf_b200.utils.postprocess import non_max_suppress_device
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    table = tiny_layer_table(a.classes, 5)
    store = variables.reset_default_store()
    store.assign({"yolo2_tiny/" + k: v for k, v in init_params(a.classes, 5, seed=1, table=table).items()})
    anchors = ho.ANCHORS_VOC if a.classes == 20 else ho.ANCHORS_COCO
    builder = Builder.from_values([str(i) for i in range(a.classes)], a.size, a.size, anchors, inference_name="tiny")
    g = torch.Generator(device="cuda").manual_seed(2)
    xs = [torch.randn(a.batch, a.size, a.size, 3, device="cuda", generator=g) for _ in range(4)]

    def step(i):
        builder(xs[i % 4])
        m = builder.model
        conf = m.conf.reshape(a.batch, -1, a.classes)
        non_max_suppress_device(conf, m.xy_min.reshape(a.batch, -1, 2), m.xy_max.reshape(a.batch, -1, 2), 0.3, 0.4)
    for i in range(a.warmup):
        step(i)
    torch.cuda.synchronize()
    n0 = _lib.lib().y2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    ms = e0.elapsed_time(e1) / a.steps
    gf = flops_per_image(a.size, a.size, a.classes, 5, table=table) / 1e9
    print(json.dumps({"metric": "images/sec tiny-YOLOv2 inference+NMS", "value": a.batch / ms * 1e3, "unit": "images/s", "ms_per_step": ms,
                      "batch": a.batch, "size": a.size, "classes": a.classes, "gflop_per_image": gf,
                      "algorithmic_tflops": a.batch * gf / ms, "gpu_launches_per_step": (_lib.lib().y2_launch_count() - n0) / a.steps,
                      "steps": a.steps, "warmup": a.warmup}))


if __name__ == "__main__":
    main()
