"""Generates tests/golden/darknet_transpose.npz with the REFERENCE'S OWN final-layer permutation functions
(/root/reference/parse_darknet_yolo2.py: transpose_weights, transpose_biases), run in the authoring container.
The module imports TensorFlow / matplotlib at the top (absent here); they are replaced by inert stubs ONLY so that the
import statement succeeds -- the two functions called are pure numpy.  Run once, here:
    python tests/golden/make_darknet_golden.py
/root/reference does not exist on the GPU box; only the committed .npz travels."""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference/parse_darknet_yolo2.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    stubs = {}
    for name in ["tensorflow", "tensorflow.contrib", "tensorflow.contrib.slim", "tensorflow.python", "tensorflow.python.framework",
                 "tensorflow.python.framework.ops", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "model", "model.yolo2",
                 "model.yolo2.inference", "utils"]:
        m = types.ModuleType(name)
        m.__dict__.setdefault("__path__", [])
        m.__getattr__ = lambda attr, _n=name: mock.MagicMock(name=_n + "." + attr)      # any attribute access works
        stubs[name] = m
    with mock.patch.dict(sys.modules, stubs):
        spec = importlib.util.spec_from_file_location("ref_parse_darknet_yolo2", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    rs = np.random.RandomState(11)
    out = {"numpy_version": np.__version__}
    for tag, (k, cin, anchors, classes) in {"voc": (1, 48, 5, 20), "coco": (1, 32, 5, 80), "small": (3, 8, 3, 2)}.items():
        w = rs.normal(0, 1, size=(k, k, cin, anchors * (5 + classes))).astype(np.float32)
        b = rs.normal(0, 1, size=(anchors * (5 + classes),)).astype(np.float32)
        out[tag + "_w_in"], out[tag + "_b_in"] = w, b
        out[tag + "_anchors"] = np.int32(anchors)
        out[tag + "_w_out"] = ref.transpose_weights(w, anchors)
        out[tag + "_b_out"] = ref.transpose_biases(b, anchors)
    np.savez_compressed(os.path.join(HERE, "darknet_transpose.npz"), **out)
    print("wrote darknet_transpose.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
