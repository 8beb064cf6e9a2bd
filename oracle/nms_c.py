"""ctypes loader for the C twin of the NMS oracle (oracle/nms_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnms_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        try:
            build()                                   # make: a no-op when libnms_oracle.so is newer than nms_oracle.c
        except Exception:
            if not os.path.exists(_SO):
                raise
        _lib = ctypes.CDLL(_SO)
        _lib.y2o_nms_batch.restype = ctypes.c_int
        _lib.y2o_nms_batch.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_float] * 2 + [ctypes.c_void_p]
    return _lib


def nms_c_batch(conf, xy_min, xy_max, threshold, threshold_iou):
    """conf [B,N,C] float32 C-contiguous (mutated in place); returns order [B,N] int32.
    Raises AssertionError where the reference's asserts would fire."""
    assert conf.dtype == np.float32 and conf.flags.c_contiguous
    xy_min = np.ascontiguousarray(xy_min, dtype=np.float32)
    xy_max = np.ascontiguousarray(xy_max, dtype=np.float32)
    b, n, c = conf.shape
    order = np.empty((b, n), dtype=np.int32)
    rc = lib().y2o_nms_batch(conf.ctypes.data, xy_min.ctypes.data, xy_max.ctypes.data, b, n, c,
                             float(np.float32(threshold)), float(np.float32(threshold_iou)), order.ctypes.data)
    if rc != 0:
        raise AssertionError("reference assert (NaN or xy_min > xy_max) would fire")
    return order
