"""CPU (-m "not gpu"): error behaviour of the C ABI (include/yolo2_b200.h) on a machine WITHOUT a GPU.

The reference signals bad input with Python asserts / exceptions (utils/__init__.py:52-56, utils/postprocess.py:22-27,
utils/data/__init__.py:119,142); the library's contract is "return < 0 and leave a message in y2_last_error()", never a
crash and never a silent CPU fallback.  Every check below is argument validation that runs BEFORE the first CUDA call, so
it can be exercised here; the pointers passed are never dereferenced on the host.  No compute call succeeds in this file.
"""
import ctypes

import pytest

from yolo_tf_b200 import _lib

P = ctypes.c_void_p(0x1000)          # an opaque non-NULL "device pointer"; validation must reject before touching it
BIG = 1 << 30


def _err():
    return _lib.lib().y2_last_error().decode()


def _hp():
    return (ctypes.c_float * 4)(1, 5, 1, 1)


def test_create_rejects_bad_arguments_and_missing_device():
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.y2_create_net(ctypes.byref(h), 0, 20, 5, 7) == -1 and "unknown architecture" in _err()
    assert L.y2_create_net(ctypes.byref(h), 0, 0, 5, 0) == -1 and "positive" in _err()
    assert L.y2_create_net(ctypes.byref(h), 0, 20, 0, 1) == -1 and "positive" in _err()
    import torch
    if not torch.cuda.is_available():
        # no device: creation fails loudly with the CUDA runtime's own message, it does not fall back to anything
        assert L.y2_create(ctypes.byref(h), 0, 20, 5) < 0 and "cuda" in _err().lower()
        assert not h.value


def test_null_handle_is_rejected_everywhere():
    L = _lib.lib()
    i = ctypes.c_int()
    z = ctypes.c_size_t()
    assert L.y2_num_layers(None) == -1 and "null handle" in _err()
    assert L.y2_layer_info(None, 0, ctypes.byref(i), ctypes.byref(i), ctypes.byref(i), ctypes.byref(i)) == -1
    assert L.y2_load_weights(None, 0, P, P, P, P, P, None, None) == -1
    assert L.y2_workspace_bytes(None, 1, 32, 32) == 0
    assert L.y2_darknet_forward(None, P, 1, 32, 32, P, P, 0, 0, None) == -1 and "null" in _err()
    assert L.y2_set_option(None, b"halo", 1) == -1
    assert L.y2_set_profiling(None, 1) == -1
    assert L.y2_get_layer_ms(None, None, None) == -1
    assert L.y2_get_activation(None, 0, 0, P, None) == -1
    assert L.y2_train_workspace_bytes(None, 1, 32, 32) == 0
    assert L.y2_darknet_forward_train(None, P, 1, 32, 32, P, P, 0, None) == -1
    assert L.y2_darknet_backward(None, P, P, None) == -1
    assert L.y2_param_count(None) == 0
    assert L.y2_param_offsets(None, 0, ctypes.byref(z), ctypes.byref(z), ctypes.byref(z)) == -1
    assert L.y2_get_bn_state(None, 0, None, None, None, None, None) == -1
    assert L.y2_num_param_tensors(None) == -1
    assert L.y2_adam_workspace_bytes(None) == 0
    assert L.y2_adam_step(None, P, P, P, None, 0, 1e-3, 0.9, 0.999, 1e-8, 1, 0.0, P, 0, None) == -1
    assert L.y2_train_probe(None, 0, None, None) == -1
    assert L.y2_train_get_tensor(None, 0, 0, P, None) == -1
    L.y2_destroy(None)               # like free(NULL)


def test_reorg_validation():
    """model/yolo2/function.py:22-29 reshapes [B, H/stride, stride, W/stride, stride, C]: H, W must divide."""
    L = _lib.lib()
    assert L.y2_leaky_relu(None, 16, 0.1, P, None) == -1 and "null" in _err()
    assert L.y2_leaky_relu(P, 16, 0.1, None, None) == -1
    assert L.y2_leaky_relu(P, 0, 0.1, P, None) == 0                  # nothing to do, nothing launched
    assert L.y2_reorg(None, 1, 4, 4, 4, 2, P, None) == -1 and "null" in _err()
    assert L.y2_reorg(P, 1, 4, 4, 4, 2, None, None) == -1
    assert L.y2_reorg(P, 1, 5, 4, 4, 2, P, None) == -1 and "divisible" in _err()
    assert L.y2_reorg(P, 1, 4, 6, 4, 4, P, None) == -1
    assert L.y2_reorg(P, 1, 4, 4, 4, 0, P, None) == -1
    assert L.y2_reorg(P, 0, 4, 4, 4, 2, P, None) == 0          # empty batch: nothing to do, nothing launched


def test_head_decode_and_loss_validation():
    L = _lib.lib()
    outs = _lib.HeadOutputs()
    assert L.y2_head_decode(None, 1, 13, 13, 5, 20, P, ctypes.byref(outs), None) == -1 and "null" in _err()
    assert L.y2_head_decode(P, 1, 13, 13, 5, 20, None, ctypes.byref(outs), None) == -1
    assert L.y2_head_decode(P, 1, 13, 13, 5, 20, P, None, None) == -1
    assert L.y2_head_decode(P, 1, 13, 13, 0, 20, P, ctypes.byref(outs), None) == -1 and "shape" in _err()
    assert L.y2_head_decode(P, 1, 13, 13, 5, 0, P, ctypes.byref(outs), None) == -1
    assert L.y2_head_decode(P, 1, 0, 13, 5, 20, P, ctypes.byref(outs), None) == -1
    need = L.y2_loss_workspace_bytes(64, 13, 13)
    assert 0 < need < (1 << 20)
    args = (P, 64, 13, 13, 5, 20, P, P, P, P, P, P, P)
    assert L.y2_loss_fwd_bwd(*args, _hp(), P, P, P, need - 1, None) == -1 and "workspace too small" in _err()
    assert L.y2_loss_fwd_bwd(*args, None, P, P, P, need, None) == -1 and "null" in _err()
    assert L.y2_loss_fwd_bwd(*args, _hp(), None, P, P, need, None) == -1             # objectives are mandatory ...
    assert L.y2_loss_fwd_bwd(None, *args[1:], _hp(), P, None, P, need, None) == -1  # ... dnet is the nullable one


def test_nms_validation():
    L = _lib.lib()
    assert L.y2_nms_workspace_bytes(0, 0, 0) == 0
    need = L.y2_nms_workspace_bytes(1, 845, 20)
    assert need > 0
    # BASELINE config 5, largest cell: B = 512, N = 1805 (19 x 19 x 5), C = 80 -- the query must not overflow
    big = L.y2_nms_workspace_bytes(512, 1805, 80)
    assert need < big < (4 << 30)
    assert L.y2_nms(None, P, P, 1, 845, 20, 0.3, 0.4, None, None, P, BIG, None) == -1 and "null" in _err()
    assert L.y2_nms(P, None, P, 1, 845, 20, 0.3, 0.4, None, None, P, BIG, None) == -1
    assert L.y2_nms(P, P, P, 1, 845, 20, 0.3, 0.4, None, None, None, BIG, None) == -1
    assert L.y2_nms(P, P, P, 1, 845, 20, 0.3, 0.4, None, None, P, need - 1, None) == -1 and "workspace too small" in _err()
    # the reference returns an empty list for zero boxes and is never called for zero images: no-ops here
    assert L.y2_nms(P, P, P, 0, 845, 20, 0.3, 0.4, None, None, P, BIG, None) == 0
    assert L.y2_nms(P, P, P, 1, 0, 20, 0.3, 0.4, None, None, P, BIG, None) == 0


def test_prepost_and_label_validation():
    L = _lib.lib()
    assert L.y2_per_image_standardization(P, 3, 1, 100, P, P, BIG, None) == -1 and "elem_bytes" in _err()
    assert L.y2_per_image_standardization(P, 4, 1, 0, P, P, BIG, None) == -1 and "empty" in _err()
    assert L.y2_per_image_standardization(None, 4, 1, 100, P, P, BIG, None) == -1
    need = L.y2_standardize_workspace_bytes(32, 416 * 416 * 3)
    assert need > 0
    assert L.y2_per_image_standardization(P, 1, 32, 416 * 416 * 3, P, P, need - 1, None) == -1
    assert L.y2_detections(None, P, P, 1, 845, 20, 0.3, 32, 32, P, P, P, P, P, None) == -1 and "null" in _err()
    assert L.y2_detections(P, P, P, 1, 845, 20, 0.3, 32, 32, None, P, P, P, P, None) == -1
    assert L.y2_transform_labels(P, P, P, 0, 20, 13, 13, P, P, P, P, P, P, None, None) == -1
    assert L.y2_transform_labels(P, P, P, 1, 0, 13, 13, P, P, P, P, P, P, None, None) == -1
    assert L.y2_transform_labels(P, P, None, 1, 20, 13, 13, P, P, P, P, P, P, None, None) == -1


def test_single_conv_entry_points_validate_shapes_first():
    """slim.layers.conv2d is only ever called with kernel_size 1 or 3 on this path (model/yolo2/inference.py:73-118)."""
    L = _lib.lib()
    assert L.y2_conv2d(None, 1, 8, 8, 32, P, 3, 32, None, None, 0, P, 0, 0, 0, None) == -1 and "null" in _err()
    assert L.y2_conv2d(P, 1, 8, 8, 32, P, 2, 32, None, None, 0, P, 0, 0, 0, None) == -1 and "ksize" in _err()
    assert L.y2_conv2d(P, 1, 8, 8, 24, P, 3, 32, None, None, 0, P, 0, 0, 0, None) == -1 and "multiple of 32" in _err()
    assert L.y2_conv2d(P, 0, 8, 8, 32, P, 3, 32, None, None, 0, P, 0, 0, 0, None) == -1 and "bad shape" in _err()
    assert L.y2_conv2d(P, 1, 8, 8, 32, P, 3, 32, None, None, 0, P, 0, 48, 0, None) == -1 and "block_n" in _err()
    assert L.y2_conv2d_wgrad(P, 1, 8, 8, 32, P, 2, 32, P, 0, None) == -1 and "ksize" in _err()
    assert L.y2_conv2d_wgrad(P, 1, 8, 8, 32, None, 3, 32, P, 0, None) == -1 and "null" in _err()
    assert L.y2_conv2d_wgrad(P, 1, 0, 8, 32, P, 3, 32, P, 0, None) == -1 and "bad shape" in _err()


def test_check_raises_with_the_library_message():
    L = _lib.lib()
    rc = L.y2_reorg(P, 1, 5, 4, 4, 2, P, None)
    with pytest.raises(_lib.Y2Error, match="divisible by stride"):
        _lib.check(rc)
    import torch
    if not torch.cuda.is_available():
        assert L.y2_launch_count() == 0          # nothing above launched a kernel
