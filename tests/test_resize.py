"""Image resize of the detection path (detect.py:65 `_image.resize((width, height))`): the arithmetic lives in
Pillow, so the fixtures are Pillow's own outputs (tests/golden/resize.npz, tests/golden/make_resize_golden.py).  CPU: the numpy
oracle and the C++ core the CUDA kernels call (compiled for the host) reproduce them bit for bit; where Pillow is importable
the oracle is also compared live over random size pairs.  The GPU test lives in tests/test_gpu_unverified.py (opt-in: the
kernels have not run on a GPU yet)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle.resize_oracle import BICUBIC, NEAREST, resize_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "resize.npz")
CASES = ("shrink", "enlarge", "mixed", "same_w", "tiny", "strong")


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_pillow_bit_for_bit(name):
    g = np.load(GOLD)
    oh, ow = (int(v) for v in g[name + "_size"])
    assert np.array_equal(resize_oracle(g[name + "_in"], ow, oh, BICUBIC), g[name + "_bicubic"])
    assert np.array_equal(resize_oracle(g[name + "_in"], ow, oh, NEAREST), g[name + "_nearest"])


def test_oracle_against_live_pillow_over_random_sizes():
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(5)
    for _ in range(25):
        h, w, oh, ow = rs.randint(1, 300), rs.randint(1, 300), 32 * rs.randint(1, 8), 32 * rs.randint(1, 8)
        img = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        for flt in (BICUBIC, NEAREST):
            assert np.array_equal(resize_oracle(img, ow, oh, flt), np.asarray(Image.fromarray(img).resize((ow, oh), flt))), (h, w, oh, ow, flt)


def test_device_core_compiled_for_the_host_reproduces_pillow(tmp_path):
    """yolo_tf_b200/csrc/y2_resize_core.cuh (table builders in double + the per-element functions every CUDA thread runs), driven
    in the kernels' element order by tests/host/resize_harness.cu."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "resize_harness")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tests", "host", "resize_harness.cu")])
    g = np.load(GOLD)
    for name in CASES:
        img = g[name + "_in"]
        oh, ow = (int(v) for v in g[name + "_size"])
        src, dst = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
        img.tofile(src)
        for flt, key in ((BICUBIC, "_bicubic"), (NEAREST, "_nearest")):
            subprocess.check_call([exe, src, str(img.shape[0]), str(img.shape[1]), "3", str(oh), str(ow), str(flt), dst])
            assert np.array_equal(np.fromfile(dst, dtype=np.uint8).reshape(oh, ow, 3), g[name + key]), (name, flt)


def test_resize_abi_validation_and_no_cpu_path():
    import ctypes
    import torch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.utils import preprocess
    L = _lib.lib()
    P = ctypes.c_void_p(0x1000)
    assert L.y2_resize_workspace_bytes(480, 640, 416, 416, 3, 3) > 480 * 416 * 3
    assert 0 < L.y2_resize_workspace_bytes(480, 640, 416, 416, 3, 0) < 8192
    assert L.y2_resize_workspace_bytes(480, 640, 416, 416, 3, 1) == 0                     # BILINEAR etc. are not built
    assert L.y2_resize_u8(None, 4, 4, 3, P, 8, 8, 3, P, 1 << 20, None) == -1
    assert L.y2_resize_u8(P, 4, 4, 3, P, 8, 8, 2, P, 1 << 20, None) == -1 and b"resample" in L.y2_last_error()
    assert L.y2_resize_u8(P, 0, 4, 3, P, 8, 8, 3, P, 1 << 20, None) == -1
    assert L.y2_resize_u8(ctypes.c_void_p(0x1000), 4, 4, 3, P, 8, 8, 3, ctypes.c_void_p(0x1000), 16, None) == -1 and b"workspace" in L.y2_last_error()
    with pytest.raises(_lib.Y2Error):
        preprocess.resize(torch.zeros(4, 4, 3, dtype=torch.uint8), 8, 8)
