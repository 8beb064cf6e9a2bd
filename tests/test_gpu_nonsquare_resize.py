"""-m gpu: non-square inputs and the device image resize (first executed on a B200 at the start of round 2: 18 / 18 green).

* non-square input: the reference takes width and height separately (config [yolo2] width / height, utils/__init__.py:52-56);
  the library's entry points do too: darknet / tiny at 96 x 64 against the reference's own graph builder (golden fixture).
* `resize` (detect.py:65): device bicubic / nearest against Pillow's own outputs (tests/golden/resize.npz), bit for bit."""
import os

import numpy as np
import pytest

from oracle.darknet_oracle import init_params, tiny_layer_table

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "backbone_reference.npz")


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


def test_darknet_forward_non_square_vs_the_reference_graph_golden(cuda):
    import torch
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import inference
    d = np.load(GOLD)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in init_params(20, 5, seed=1).items()})
    x = torch.from_numpy(d["x96"]).to(cuda)                       # [2, 96, 64, 3]: 3 x 2 cells
    _, out = inference.darknet(x, 20, 5)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    assert tuple(out.shape) == (2, 3, 2, 125)
    assert _rel(out.cpu().numpy(), d["darknet_rect_out"]) <= 1e-4


def test_tiny_forward_non_square_vs_the_reference_graph_golden(cuda):
    import torch
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import inference
    d = np.load(GOLD)
    store = variables.reset_default_store()
    store.assign({"yolo2_tiny/" + k: v for k, v in init_params(20, 5, seed=1, table=tiny_layer_table(20, 5)).items()})
    x = torch.from_numpy(d["x96"]).to(cuda)
    scope, out = inference.tiny(x, 20, 5)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    assert scope == "yolo2_tiny" and tuple(out.shape) == (2, 3, 2, 125)
    assert _rel(out.cpu().numpy(), d["tiny_out"]) <= 1e-4


@pytest.mark.parametrize("name", ["shrink", "enlarge", "mixed", "same_w", "tiny", "strong"])
def test_resize_on_the_device_reproduces_pillow_bit_for_bit(cuda, name):
    """csrc/y2_resize.cu against Pillow's own outputs (tests/golden/resize.npz); the same per-element code passes on the CPU
    (tests/test_resize.py), so what this adds is the kernels' launch geometry and table upload."""
    import torch
    from yolo_tf_b200.utils import preprocess
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize.npz"))
    oh, ow = (int(v) for v in g[name + "_size"])
    x = torch.from_numpy(g[name + "_in"]).to(cuda)
    for flt, key in ((preprocess.BICUBIC, "_bicubic"), (preprocess.NEAREST, "_nearest")):
        got = preprocess.resize(x, ow, oh, flt).cpu().numpy()
        assert np.array_equal(got, g[name + key]), (name, flt, int((got != g[name + key]).sum()))


def test_resize_then_standardize_is_the_detect_preprocessing(cuda):
    """detect.py:65: uint8 image -> resize -> float32 -> per_image_standardization, all on the device from one uint8 upload."""
    import torch
    from oracle.prepost_oracle import per_image_standardization_oracle
    from oracle.resize_oracle import resize_oracle
    from yolo_tf_b200.utils import preprocess
    rs = np.random.RandomState(8)
    img = rs.randint(0, 256, size=(375, 500, 3)).astype(np.uint8)
    dev = preprocess.per_image_standardization(preprocess.resize(torch.from_numpy(img).to(cuda), 416, 416)).cpu().numpy()
    ref = np.asarray(per_image_standardization_oracle(resize_oracle(img, 416, 416).astype(np.float32)), dtype=np.float64)
    assert dev.shape == (416, 416, 3) and np.abs(dev - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1.0)
