"""GPU, opt-in (Y2_EXPERIMENTAL=1): checks of SHIPPED code paths that no GPU run has exercised yet (written after the round's GPU
budget was spent).  Skipped by default so that the suite the driver runs only contains verified expectations; the first GPU
call of the next round runs them (tools/round2_first.sh) and the ones that pass move into the regular files.

* non-square input: the reference takes width and height separately (config [yolo2] width / height, utils/__init__.py:52-56);
  the library's entry points do too, but every GPU test so far used square images."""
import os

import numpy as np
import pytest

from oracle.darknet_oracle import init_params, tiny_layer_table

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("Y2_EXPERIMENTAL") != "1", reason="not yet run on a GPU: set Y2_EXPERIMENTAL=1")]
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
