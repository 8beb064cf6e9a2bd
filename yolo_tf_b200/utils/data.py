"""utils/data/__init__.py of the reference, the part on the training path: ``transform_labels`` (:112-145), the label encoder
train.py runs per image through ``tf.py_func`` (``decode_labels`` :148-160).  Here it runs batched on the device
(csrc/y2_prepost.cu, ``y2_transform_labels``): ragged per-image object lists in, the six label tensors
``Builder.create_objectives`` consumes out.  No CPU path exists.
"""
import numpy as np

from .. import _lib


def transform_labels_batch(objects_class, objects_coord, classes, cell_width, cell_height, device=None):
    """objects_class: list (one entry per image) of int arrays [n_b]; objects_coord: list of float arrays [n_b, 4] =
    (xmin, ymin, xmax, ymax) normalised to the image.  Returns CUDA float32 tensors
    ``(mask [B,cells,1], prob [B,cells,1,classes], coords [B,cells,1,4], offset_xy_min [B,cells,1,2],
    offset_xy_max [B,cells,1,2], areas [B,cells,1])`` -- the reference's per-image outputs stacked by the batch queue
    (train.py:107).  Raises IndexError / AssertionError where the reference does (cell or class index out of range; negative
    width or height)."""
    import torch
    if len(objects_class) != len(objects_coord):
        raise AssertionError("objects_class and objects_coord differ in length")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _lib.Y2Error("transform_labels: device must be a CUDA device (no CPU path exists)")
    b = len(objects_class)
    counts = []
    for c, xy in zip(objects_class, objects_coord):
        c, xy = np.asarray(c), np.asarray(xy)
        assert len(c) == len(xy)                                      # utils/data/__init__.py:119
        counts.append(len(c))
    offsets = np.zeros(b + 1, np.int32)
    np.cumsum(counts, out=offsets[1:])
    total = int(offsets[-1])
    cls = np.concatenate([np.asarray(c, np.int32).reshape(-1) for c in objects_class]) if total else np.zeros(1, np.int32)
    xy = np.concatenate([np.asarray(c, np.float32).reshape(-1, 4) for c in objects_coord]) if total else np.zeros((1, 4), np.float32)
    # one pinned staging buffer -> one H2D copy for the ragged lists
    host = torch.empty(cls.size + xy.size + offsets.size, dtype=torch.int32).pin_memory()
    hv = host.numpy()
    hv[:cls.size] = cls
    hv[cls.size:cls.size + xy.size] = xy.reshape(-1).view(np.int32)
    hv[cls.size + xy.size:] = offsets
    d = host.to(dev, non_blocking=True)
    cls_d, xy_d, off_d = d[:cls.size], d[cls.size:cls.size + xy.size].view(torch.float32), d[cls.size + xy.size:]
    cells = cell_width * cell_height
    f32 = dict(dtype=torch.float32, device=dev)
    outs = (torch.empty((b, cells, 1), **f32), torch.empty((b, cells, 1, classes), **f32), torch.empty((b, cells, 1, 4), **f32),
            torch.empty((b, cells, 1, 2), **f32), torch.empty((b, cells, 1, 2), **f32), torch.empty((b, cells, 1), **f32))
    status = torch.empty(b, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().y2_transform_labels(_lib.ptr(cls_d.contiguous()), _lib.ptr(xy_d.contiguous()), _lib.ptr(off_d.contiguous()), b,
                                              int(classes), int(cell_width), int(cell_height), *[_lib.ptr(t) for t in outs],
                                              _lib.ptr(status), _lib.current_stream()))
    st = status.cpu().numpy()
    if (st & 1).any():
        raise IndexError("transform_labels: object cell / class index out of range in image(s) %s" % np.nonzero(st & 1)[0].tolist())
    if (st & 2).any():
        raise AssertionError("transform_labels: negative object width/height in image(s) %s" % np.nonzero(st & 2)[0].tolist())
    return outs


def transform_labels(objects_class, objects_coord, classes, cell_width, cell_height, dtype=np.float32):
    """Reference signature and return value (utils/data/__init__.py:112): one image, numpy in, numpy out
    ``(mask [cells,1], prob [cells,1,classes], coords [cells,1,4], offset_xy_min, offset_xy_max [cells,1,2], areas [cells,1])``,
    computed on the current CUDA device."""
    if np.dtype(dtype) != np.float32:
        raise TypeError("transform_labels: the device encoder computes in float32 (the dtype train.py uses)")
    outs = transform_labels_batch([objects_class], [objects_coord], classes, cell_width, cell_height)
    return tuple(t[0].cpu().numpy() for t in outs)
