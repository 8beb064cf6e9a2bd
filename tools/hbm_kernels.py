"""Driver for the ncu capture of the HBM-bound kernels (tools/profile_hbm.sh): after an unprofiled warm-up it runs, between
cudaProfilerStart / Stop, ONE detection step (B=32, 416x416, C=80: standardise, conv0, reorg, decode, NMS, detections), ONE
training step + Adam (B=64, 416x416, C=20: BN statistics / apply / backward passes, pool / reorg backward, loss, optimizer)
and the NMS at the BASELINE configs[4] sweep points B=512, N=845, K = 100 / 1000 / 10000.
    python tools/hbm_kernels.py [--skip-train]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from yolo_tf_b200 import _lib, variables  # noqa: E402
from yolo_tf_b200.model.yolo2 import Builder  # noqa: E402
from yolo_tf_b200.optimizer import AdamOptimizer, create_train_op  # noqa: E402
from yolo_tf_b200.utils.data import transform_labels_batch  # noqa: E402
from yolo_tf_b200.utils.postprocess import detections_device, non_max_suppress_device  # noqa: E402
from yolo_tf_b200.utils.preprocess import per_image_standardization  # noqa: E402


def nms_inputs(rs, B, g, C, K):
    A, cells = 5, g * g
    N = cells * A
    anch = np.asarray(bench.ANCHORS_COCO)
    gy, gx = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    centre = np.stack([gx, gy], -1).reshape(1, cells, 1, 2) + rs.uniform(0, 1, size=(B, cells, A, 2))
    wh = anch.reshape(1, 1, A, 2) * np.exp(rs.normal(0, 0.5, size=(B, cells, A, 2)))
    lo = (centre - wh / 2).astype(np.float32).reshape(B, N, 2)
    hi = (centre + wh / 2).astype(np.float32).reshape(B, N, 2)
    conf = rs.uniform(0, 0.29, size=(B, N * C)).astype(np.float32)
    for b in range(B):
        conf[b, rs.choice(N * C, size=K, replace=False)] = rs.uniform(0.3, 1.0, size=K)
    return conf.reshape(B, N, C), lo, hi


def main():
    skip_train = "--skip-train" in sys.argv
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    # ---- detection step
    B, size, C = 32, 416, 80
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in bench.synthetic_checkpoint(C, 5).items()})
    builder = Builder.from_values([str(i) for i in range(C)], size, size, bench.ANCHORS_COCO)
    u8 = torch.from_numpy(bench.synthetic_images_u8(rs, B, size)).to(dev)
    N = 13 * 13 * 5

    def detect_step():
        builder(per_image_standardization(u8))
        m = builder.model
        conf, lo, hi = m.conf.view(B, N, C), m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2)
        non_max_suppress_device(conf, lo, hi, 0.3, 0.4, check=False)
        detections_device(conf, lo, hi, 0.3, (32.0, 32.0))

    for _ in range(3):
        detect_step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    detect_step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    # ---- NMS sweep points
    for K in (100, 1000, 10000):
        conf, lo, hi = nms_inputs(np.random.RandomState(5), 512, 13, 80, K)
        d0, dlo, dhi = torch.from_numpy(conf).to(dev), torch.from_numpy(lo).to(dev), torch.from_numpy(hi).to(dev)
        work = d0.clone()
        non_max_suppress_device(work, dlo, dhi, 0.3, 0.4, check=False)
        work.copy_(d0)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        non_max_suppress_device(work, dlo, dhi, 0.3, 0.4, check=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        del d0, work
    if skip_train:
        return
    # ---- training step + Adam
    B, C = 64, 20
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in bench.synthetic_checkpoint(C, 5, dense_detections=False).items()})
    builder = Builder.from_values([str(i) for i in range(C)], size, size, bench.ANCHORS_VOC, hparam=bench.HPARAM)
    top = create_train_op(builder, AdamOptimizer(1e-6), clip_gradient_norm=1.0)
    x = torch.from_numpy(rs.normal(0, 1, size=(B, size, size, 3)).astype(np.float32)).to(dev)
    labels = list(transform_labels_batch(*bench.synthetic_boxes(rs, B, C), C, 13, 13, device=dev))
    for _ in range(2):
        top(x, labels)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    top(x, labels)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    _lib.check(_lib.lib().y2_check_async_errors())


if __name__ == "__main__":
    main()
