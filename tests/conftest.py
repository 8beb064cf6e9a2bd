import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:      # tests/tests_gpu_train_helpers.py
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # layer-by-layer parity tests read every layer's output back after the forward: one workspace slot per layer
    # (the product default, two alternating arenas, is covered by test_gpu_options.py, smoke() and bench.py)
    from yolo_tf_b200.model.yolo2 import inference
    inference._Engine.KEEP_ACTIVATIONS = True
    return torch.device("cuda:0")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """A fresh checkout has no libyolo2_b200.so (built artefacts are git-ignored): build it once (nvcc cross-compiles without a
    GPU) so that the order in which test files run does not matter.  On the GPU box the prebuilt .so travels with the snapshot."""
    from yolo_tf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    yield
