"""model/yolo/function.py:21-24 -- leaky_relu(inputs, alpha=.1) = max(x, alpha*x).

Inside the network the activation never exists as a separate op: it is fused into the conv epilogues (csrc/y2_conv_tc.cu,
csrc/y2_conv_simt.cu, csrc/y2_conv0_tc.cu; LEAKY_ALPHA is the constant those kernels hard-code).  The standalone function the
reference exports is kept as an op of its own (csrc/y2_layout.cu:leaky_relu_kernel, HBM-bound, bit-exact float32)."""
from ... import _lib

LEAKY_ALPHA = 0.1


def leaky_relu(inputs, alpha=LEAKY_ALPHA, name='leaky_relu'):
    """inputs: float32 CUDA tensor of any shape; returns a new tensor (the reference returns a new TF tensor)."""
    import torch
    x = inputs.contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.lib().y2_leaky_relu(_lib.ptr(x, torch.float32), x.numel(), float(alpha), _lib.ptr(out), _lib.current_stream()))
    return out
