"""utils/postprocess.py of the reference on the GPU: greedy per-class IoU NMS, bit-exact in which
scores it zeroes and in the order it returns (csrc/y2_nms.cu).

``non_max_suppress`` keeps the reference signature and side effect (utils/postprocess.py:39-51):
numpy in, caller's ``conf`` mutated in place, list of ``(conf_row, xy_min, xy_max)`` views out.
``non_max_suppress_device`` is the batched no-host-copy form the detection pipeline uses.
"""
import numpy as np

from .. import _lib


def non_max_suppress_device(conf, xy_min, xy_max, threshold, threshold_iou, want_order=False, check=True):
    """conf [B,N,C] float32 CUDA tensor (zeroed in place), xy_min/xy_max [B,N,2].
    Returns (order [B,N] int32 or None).  Raises AssertionError where the reference's asserts
    (NaN boxes / xy_min > xy_max, postprocess.py:22-27) would."""
    import torch
    L = _lib.lib()
    b, n, c = conf.shape
    order = torch.empty((b, n), dtype=torch.int32, device=conf.device) if want_order else None
    status = torch.empty((b,), dtype=torch.int32, device=conf.device) if check else None
    nbytes = L.y2_nms_workspace_bytes(b, n, c)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=conf.device)
    _lib.check(L.y2_nms(_lib.ptr(conf, torch.float32), _lib.ptr(xy_min, torch.float32), _lib.ptr(xy_max, torch.float32),
                        b, n, c, float(threshold), float(threshold_iou), _lib.ptr(order), _lib.ptr(status),
                        _lib.ptr(ws), nbytes, _lib.current_stream()))
    if check and bool(status.any().item()):
        raise AssertionError("non_max_suppress: NaN box or xy_min > xy_max (utils/postprocess.py:22-27)")
    return order


def non_max_suppress(conf, xy_min, xy_max, threshold, threshold_iou):
    """Reference-shaped entry point: one image, numpy arrays conf [cells, A, C], xy_min/xy_max
    [cells, A, 2]; ``conf`` is modified in place; returns the list the reference returns."""
    import torch
    if not (isinstance(conf, np.ndarray) and conf.dtype == np.float32 and conf.flags.c_contiguous):
        raise TypeError("conf must be a C-contiguous float32 ndarray (it is modified in place)")
    _, _, classes = conf.shape
    score = conf.reshape(-1, classes)
    lo = np.ascontiguousarray(xy_min, dtype=np.float32).reshape(-1, 2)
    hi = np.ascontiguousarray(xy_max, dtype=np.float32).reshape(-1, 2)
    d_conf = torch.from_numpy(score).cuda().unsqueeze(0).contiguous()
    d_lo = torch.from_numpy(lo).cuda().unsqueeze(0).contiguous()
    d_hi = torch.from_numpy(hi).cuda().unsqueeze(0).contiguous()
    order = non_max_suppress_device(d_conf, d_lo, d_hi, threshold, threshold_iou, want_order=True)
    score[...] = d_conf[0].cpu().numpy()                    # the reference's in-place side effect
    return [(score[i], lo[i], hi[i]) for i in order[0].cpu().numpy()]


def detections_device(conf, xy_min, xy_max, threshold, scale):
    """The selection loop of detect.py:72-87 for a batch, on the device, after ``non_max_suppress_device``:
    per box ``index = argmax(conf_row)`` (first maximum), kept iff ``conf_row[index] > threshold``; boxes scaled from cell
    units to pixels with ``scale = [image_width / cell_width, image_height / cell_height]`` (detect.py:72).
    conf [B,N,C], xy_min/xy_max [B,N,2] float32 CUDA tensors.  Returns (count [B] int32, box [B,N] int32, cls [B,N] int32,
    score [B,N] float32, xywh [B,N,4] float32 = (x_min, y_min, width, height) in pixels); entries [b, :count[b]] are valid,
    in box-index order (the reference walks them in its NMS list order; the SET is the same)."""
    import torch
    L = _lib.lib()
    b, n, c = conf.shape
    dev = conf.device
    count = torch.empty((b,), dtype=torch.int32, device=dev)
    box = torch.empty((b, n), dtype=torch.int32, device=dev)
    cls = torch.empty((b, n), dtype=torch.int32, device=dev)
    score = torch.empty((b, n), dtype=torch.float32, device=dev)
    xywh = torch.empty((b, n, 4), dtype=torch.float32, device=dev)
    _lib.check(L.y2_detections(_lib.ptr(conf, torch.float32), _lib.ptr(xy_min, torch.float32), _lib.ptr(xy_max, torch.float32), b, n, c,
                               float(threshold), float(scale[0]), float(scale[1]), _lib.ptr(count), _lib.ptr(box), _lib.ptr(cls),
                               _lib.ptr(score), _lib.ptr(xywh), _lib.current_stream()))
    return count, box, cls, score, xywh
