"""utils/preprocess.py of the reference, the part on the detection path: ``per_image_standardization``
(utils/preprocess.py:23-25; detect.py:60-62 applies it -- via ``--preprocess std`` -- to the resized uint8 image cast to
float32 before feeding the network).  Batched on the device (csrc/y2_prepost.cu); uint8 input is cast inside the kernel, so
the host->device copy of a batch is a quarter of the float32 one.  No CPU path exists.
"""
from .. import _lib


def per_image_standardization(image):
    """image: CUDA tensor [H, W, 3] or [B, H, W, 3], uint8 or float32.  Returns float32 of the same shape:
    ``(image - mean) / max(std, 1/sqrt(n))`` per image, mean / population std over all n elements of that image."""
    import ctypes
    import torch
    if not image.is_cuda:
        raise _lib.Y2Error("per_image_standardization: input must be a CUDA tensor (no CPU path exists)")
    if image.dtype not in (torch.uint8, torch.float32):
        raise TypeError("per_image_standardization: uint8 or float32 expected, got %s" % image.dtype)
    x = image.contiguous()
    batched = x.dim() == 4
    b = x.shape[0] if batched else 1
    n = x.numel() // b
    L = _lib.lib()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    need = L.y2_standardize_workspace_bytes(b, n)
    ws = torch.empty(need + 256, dtype=torch.uint8, device=x.device)
    off = (-ws.data_ptr()) % 256
    _lib.check(L.y2_per_image_standardization(ctypes.c_void_p(x.data_ptr()), 1 if x.dtype == torch.uint8 else 4, b, n, _lib.ptr(out),
                                              ctypes.c_void_p(ws.data_ptr() + off), need, _lib.current_stream()))
    return out


NEAREST, BICUBIC = 0, 3          # PIL.Image.Resampling codes


def resize(image, width, height, resample=BICUBIC):
    """detect.py:65 `_image.resize((width, height))` on the device: image uint8 CUDA tensor [H, W, C] -> uint8 [height, width, C],
    bit for bit what Pillow's `Image.resize` returns with `resample` (BICUBIC = Pillow's default since 7.0 and what this
    container's Pillow applies to the reference's call; NEAREST = the default of the Pillow of the reference's time).
    Not yet run on a GPU (csrc/y2_resize.cu); the algorithm itself is verified against Pillow on the CPU."""
    import ctypes
    import torch
    if not image.is_cuda:
        raise _lib.Y2Error("resize: input must be a CUDA tensor (no CPU path exists)")
    if image.dtype != torch.uint8 or image.dim() != 3:
        raise TypeError("resize: uint8 [H, W, C] expected, got %s %s" % (image.dtype, tuple(image.shape)))
    x = image.contiguous()
    h, w, c = x.shape
    L = _lib.lib()
    need = L.y2_resize_workspace_bytes(h, w, int(height), int(width), c, int(resample))
    if need == 0:
        raise ValueError("resize: bad size %dx%d -> %dx%d or unknown filter %s" % (w, h, width, height, resample))
    ws = torch.empty(need + 256, dtype=torch.uint8, device=x.device)
    off = (-ws.data_ptr()) % 256
    out = torch.empty((int(height), int(width), c), dtype=torch.uint8, device=x.device)
    _lib.check(L.y2_resize_u8(ctypes.c_void_p(x.data_ptr()), h, w, c, ctypes.c_void_p(out.data_ptr()), int(height), int(width), int(resample),
                              ctypes.c_void_p(ws.data_ptr() + off), need, _lib.current_stream()))
    return out
