"""One small training step + Adam for Darknet-19 and for tiny (64x64, B=2), the target of tools/sanitize.sh."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from yolo_tf_b200 import _lib, variables  # noqa: E402
from yolo_tf_b200.model.yolo2 import Builder  # noqa: E402
from yolo_tf_b200.optimizer import AdamOptimizer, create_train_op  # noqa: E402
from yolo_tf_b200.utils.data import transform_labels_batch  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    for name in ("darknet", "tiny"):
        variables.reset_default_store()          # variables are created by the graph's own initializers
        builder = Builder.from_values([str(i) for i in range(20)], 64, 64, bench.ANCHORS_VOC, hparam=bench.HPARAM, inference_name=name)
        op = create_train_op(builder, AdamOptimizer(1e-4), clip_gradient_norm=1.0)
        x = torch.from_numpy(rs.normal(0, 1, size=(2, 64, 64, 3)).astype(np.float32)).to(dev)
        labels = list(transform_labels_batch(*bench.synthetic_boxes(rs, 2, 20), 20, 2, 2, device=dev))
        loss = float(op(x, labels))
        torch.cuda.synchronize()
        _lib.check(_lib.lib().y2_check_async_errors())
        assert np.isfinite(loss)
        print("step ok: %s loss %.4f" % (name, loss))


if __name__ == "__main__":
    main()
