"""model/yolo/function.py:21-24 -- leaky_relu(inputs, alpha=.1) = max(x, alpha*x).

In the B200 path the activation never exists as a separate op: it is fused into the conv epilogue
(csrc/y2_conv_tc.cu, csrc/y2_conv_simt.cu).  LEAKY_ALPHA is the constant those kernels hard-code."""
LEAKY_ALPHA = 0.1


def leaky_relu(inputs, alpha=LEAKY_ALPHA):
    raise NotImplementedError(
        "leaky_relu is fused into the tcgen05 conv epilogue on this backend; it is not available as a "
        "standalone op (and there is deliberately no eager fallback).")
