"""ctypes binding of libyolo2_b200.so (include/yolo2_b200.h).  Fails loudly when the CUDA
library is missing -- there is no eager/CPU fallback anywhere in this package."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libyolo2_b200.so")
_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_f = ctypes.c_float
c_sz = ctypes.c_size_t


class HeadOutputs(ctypes.Structure):
    """y2_head_outputs (include/yolo2_b200.h)."""
    FIELDS = ("conf", "xy_min", "xy_max", "iou", "prob", "wh", "areas", "xy", "offset_xy",
              "offset_xy_min", "offset_xy_max", "coords", "wh01")
    _fields_ = [(n, c_p) for n in FIELDS]


_SIGNATURES = {
    "y2_last_error": (ctypes.c_char_p, []),
    "y2_version": (c_i, []),
    "y2_create": (c_i, [ctypes.POINTER(c_p), c_i, c_i, c_i]),
    "y2_create_net": (c_i, [ctypes.POINTER(c_p), c_i, c_i, c_i, c_i]),
    "y2_destroy": (None, [c_p]),
    "y2_num_layers": (c_i, [c_p]),
    "y2_layer_info": (c_i, [c_p, c_i] + [ctypes.POINTER(c_i)] * 4),
    "y2_load_weights": (c_i, [c_p, c_i] + [c_p] * 6 + [c_p]),
    "y2_workspace_bytes": (c_sz, [c_p, c_i, c_i, c_i]),
    "y2_darknet_forward": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_sz, c_i, c_p]),
    "y2_get_activation": (c_i, [c_p, c_i, c_i, c_p, c_p]),
    "y2_conv2d": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_p]),
    "y2_leaky_relu": (c_i, [c_p, c_sz, c_f, c_p, c_p]),
    "y2_reorg": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "y2_head_decode": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_p, ctypes.POINTER(HeadOutputs), c_p]),
    "y2_loss_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "y2_loss_fwd_bwd": (c_i, [c_p] + [c_i] * 5 + [c_p] * 7 + [ctypes.POINTER(c_f), c_p, c_p, c_p, c_sz, c_p]),
    "y2_nms_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "y2_nms": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_sz, c_p]),
    "y2_check_async_errors": (c_i, []),
    "y2_set_option": (c_i, [c_p, ctypes.c_char_p, c_i]),
    "y2_conv2d_wgrad": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_i, c_p]),
    "y2_train_workspace_bytes": (c_sz, [c_p, c_i, c_i, c_i]),
    "y2_darknet_forward_train": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_sz, c_p]),
    "y2_darknet_backward": (c_i, [c_p, c_p, c_p, c_p]),
    "y2_param_count": (c_sz, [c_p]),
    "y2_param_offsets": (c_i, [c_p, c_i] + [ctypes.POINTER(c_sz)] * 3),
    "y2_get_bn_state": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p, c_p]),
    "y2_train_probe": (c_i, [c_p, c_i, c_p, c_p]),
    "y2_train_get_tensor": (c_i, [c_p, c_i, c_i, c_p, c_p]),
    "y2_set_profiling": (c_i, [c_p, c_i]),
    "y2_get_layer_ms": (c_i, [c_p, ctypes.POINTER(c_f), ctypes.POINTER(c_f)]),
    "y2_launch_count": (ctypes.c_ulonglong, []),
    "y2_num_param_tensors": (c_i, [c_p]),
    "y2_adam_workspace_bytes": (c_sz, [c_p]),
    "y2_standardize_workspace_bytes": (c_sz, [c_i, c_sz]),
    "y2_per_image_standardization": (c_i, [c_p, c_i, c_i, c_sz, c_p, c_p, c_sz, c_p]),
    "y2_resize_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i, c_i]),
    "y2_resize_u8": (c_i, [c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_sz, c_p]),
    "y2_detections": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p]),
    "y2_transform_labels": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "y2_adam_step": (c_i, [c_p, c_p, c_p, c_p, ctypes.POINTER(c_p), c_i, c_f, c_f, c_f, c_f, ctypes.c_longlong, c_f, c_p, c_sz, c_p]),
}
EXPORTS = tuple(_SIGNATURES)


class Y2Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Y2Error(
                "yolo_tf_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C yolo_tf_b200/csrc` (needs nvcc, targets sm_100a). There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)        # AttributeError if the export is missing: loud by design
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise Y2Error(lib().y2_last_error().decode("utf-8", "replace") + " (rc=%d)" % rc)


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise Y2Error("expected a CUDA tensor, got device %s (no CPU path exists)" % t.device)
    if not t.is_contiguous():
        raise Y2Error("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise Y2Error("expected dtype %s, got %s" % (dtype, t.dtype))
    return c_p(t.data_ptr())


def current_stream():
    import torch
    return c_p(torch.cuda.current_stream().cuda_stream)
