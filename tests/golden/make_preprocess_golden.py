"""Generates tests/golden/standardize.npz with the REFERENCE'S OWN per_image_standardization
(/root/reference/utils/preprocess.py:23-25), loaded by file path with `tensorflow` replaced by an inert stub (the module
imports it at the top; the function called is pure numpy).  Run once, here:  python tests/golden/make_preprocess_golden.py"""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference/utils/preprocess.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    tf = types.ModuleType("tensorflow")
    tf.__getattr__ = lambda attr: mock.MagicMock(name="tensorflow." + attr)
    with mock.patch.dict(sys.modules, {"tensorflow": tf}):
        spec = importlib.util.spec_from_file_location("ref_preprocess", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    rs = np.random.RandomState(17)
    out = {"numpy_version": np.__version__}
    cases = {"rand32": rs.randint(0, 256, size=(32, 32, 3)), "rand64x48": rs.randint(0, 256, size=(64, 48, 3)),
             "dark": rs.randint(0, 4, size=(40, 40, 3)), "constant": np.full((16, 16, 3), 77), "one_hot": np.eye(24 * 24 * 3)[5].reshape(24, 24, 3) * 255}
    for name, img in cases.items():
        u8 = img.astype(np.uint8)
        out[name + "_u8"] = u8
        res = ref.per_image_standardization(u8.astype(np.float32))       # detect.py:62: uint8 -> float32 -> preprocess
        out[name + "_out"] = np.asarray(res)
        out[name + "_out_dtype"] = str(np.asarray(res).dtype)
    np.savez_compressed(os.path.join(HERE, "standardize.npz"), **out)
    print({k: (getattr(v, "shape", v), getattr(v, "dtype", "")) for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
