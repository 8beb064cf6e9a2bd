"""model/yolo2/function.py:22-29 -- reorg(net, stride=2): space-to-depth in TF ordering,
out[b, y, x, (dy*stride+dx)*C + c] = in[b, stride*y+dy, stride*x+dx, c]."""
from ... import _lib


def reorg(net, stride=2, name='reorg'):
    """net: float32 CUDA tensor [B, H, W, C] (static shape, like function.py:23)."""
    import torch
    b, h, w, c = net.shape
    if h % stride or w % stride:
        raise ValueError("reorg: spatial size %dx%d not divisible by stride %d" % (h, w, stride))
    out = torch.empty((b, h // stride, w // stride, c * stride * stride), dtype=torch.float32, device=net.device)
    _lib.check(_lib.lib().y2_reorg(_lib.ptr(net, torch.float32), b, h, w, c, stride, _lib.ptr(out), _lib.current_stream()))
    return out
