"""CPU (-m "not gpu"): the float32 IoU arithmetic the NMS kernels use -- iou_ref / iou_hit / box_area / ford in
yolo_tf_b200/csrc/y2_nms_iou.cuh, the SAME source text the device code is compiled from -- built for the host
(tests/host/nms_iou_harness.cu) and compared with the numpy oracle (oracle/nms_oracle.py:pair_iou = utils/postprocess.py:21-36)
bit for bit on a fuzz that sits on the edges: nested boxes with area ratios within ulps of the threshold, identical and disjoint
boxes, zero-area and denormal boxes, areas that overflow float32 (np.maximum propagates the NaN of inf + inf - inf: no hit), and
thresholds <= 0 and = 1.  Also: the conservative filter of iou_hit never changes an answer, and ford() orders like float '<'."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from oracle.nms_oracle import pair_iou
from test_nms_cull_bound import f32, pairs                  # tests/ is on sys.path (conftest.py)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("nmsiou") / "nms_iou_harness")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "nms_iou_harness.cu")])
    return exe


def run(exe, tmp_path, thr, k_lo, k_hi, o_lo, o_hi):
    n = len(k_lo)
    src, dst = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(src, "wb") as f:
        f.write(struct.pack("<if", n, float(f32(thr))))
        f.write(np.concatenate([k_lo, k_hi, o_lo, o_hi], axis=1).astype(f32).tobytes())
    out = subprocess.run([exe, src, dst], capture_output=True, text=True)
    assert out.returncode == 0 and "NMS IOU HARNESS OK" in out.stdout, out.stdout + out.stderr
    return np.fromfile(dst, dtype=np.uint32).reshape(n, 4)


@pytest.mark.parametrize("thr", [0.4, 0.5, 1.0, 0.05, 0.0, -0.5])
def test_kernel_iou_source_matches_the_numpy_oracle(harness, tmp_path, thr):
    rs = np.random.RandomState(int(abs(thr) * 1000) + 7)
    t = thr if thr > 0 else 0.4
    k_lo, k_hi, o_lo, o_hi = pairs(rs, 60000, t)
    # add pairs of boxes whose areas overflow float32, nested and overlapping, and a block of disjoint pairs
    big = np.exp(rs.uniform(55, 60, size=(2000, 2))).astype(f32)
    c = rs.uniform(-1, 1, size=(2000, 2)).astype(f32)
    k_lo[:2000], k_hi[:2000] = c - big, c + big
    o_lo[:2000], o_hi[:2000] = c - big * f32(0.5), c + big * f32(2.0)
    o_lo[2000:3000] = k_hi[2000:3000] + f32(1.0)
    o_hi[2000:3000] = o_lo[2000:3000] + f32(1.0)
    got = run(harness, tmp_path, thr, k_lo, k_hi, o_lo, o_hi)
    with np.errstate(all="ignore"):
        want = np.array([pair_iou(k_lo[i], k_hi[i], o_lo[i:i + 1], o_hi[i:i + 1])[0] for i in range(len(k_lo))], f32)
        hit = want >= f32(thr)                                         # NaN >= thr is False
    nan = np.isnan(want)
    assert nan[:2000].sum() > 500                                      # the overflow block really produces inf + inf - inf
    assert np.array_equal(np.isnan(got[:, 0].view(f32)), nan)
    assert np.array_equal(got[~nan, 0], want[~nan].view(np.uint32))
    assert np.array_equal(got[:, 1] != 0, hit)                         # with the conservative filter
    assert np.array_equal(got[:, 2] != 0, hit)                         # without it
    # ford(): same order as '<' on the areas (non-NaN), -0.0 == +0.0
    with np.errstate(all="ignore"):
        area = ((k_hi[:, 0] - k_lo[:, 0]).astype(f32) * (k_hi[:, 1] - k_lo[:, 1]).astype(f32)).astype(f32)
    ok = ~np.isnan(area)
    order = np.argsort(area[ok], kind="stable")
    fo = got[ok, 3][order]
    assert np.all(fo[1:] >= fo[:-1])
    with np.errstate(all="ignore"):                                  # inf - inf between neighbours of the sorted areas
        assert np.array_equal(np.diff(area[ok][order]) > 0, np.diff(fo.astype(np.int64)) > 0)
