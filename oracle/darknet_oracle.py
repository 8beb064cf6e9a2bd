"""CPU oracle for the Darknet-19 + passthrough backbone.  TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by
the product package).

Restates, with torch-CPU float32 (or float64 "truth") tensors:

* ``darknet_oracle``  <- ``darknet()``     model/yolo2/inference.py:61-120
* ``reorg_oracle``    <- ``reorg()``       model/yolo2/function.py:22-29
* ``leaky_oracle``    <- ``leaky_relu()``  model/yolo/function.py:21-24
* BN arithmetic: TF-1.0 ``tf.nn.batch_normalization`` as used by
  ``slim.batch_norm(center=True, scale=True, epsilon=1e-5)`` (inference.py:62-66):
  ``inv = rsqrt(var + eps) * gamma ; y = x * inv + (beta - mean * inv)``;
  training mode uses the batch mean and the *population* variance over (B,H,W).
* ``max_pool2d`` 2x2 stride 2 SAME (inference.py:69); every pooled extent is even,
  so SAME never pads.
* ``tiny_oracle``     <- ``tiny()``        model/yolo2/inference.py:25-50, including its
  2x2 **stride-1** SAME max-pool (:42): TF SAME with k=2, s=1 pads one row/column at
  the bottom/right only (pad_total = 1, pad_before = 0) and max-pool ignores padding.

PINNED TO THE REFERENCE'S SOURCE, NOT TO TENSORFLOW'S ARITHMETIC: tests/golden/backbone_reference.npz
holds the outputs of the reference's own graph builders ``darknet()`` / ``_darknet()`` / ``tiny()`` /
``_tiny()``, ``reorg()`` and ``leaky_relu()``, compiled from the reference files and run with a torch
float64 stand-in for the slim / tf calls they make (tests/golden/make_backbone_golden.py); this
oracle matches them to 1e-9 incl. training-mode statistics, and the variable names / shapes the
graph creates are the ones the product looks up (tests/test_backbone_reference_golden.py).  The
reorg self-test vector (function.py:32-50) and the TF definitions written out as loops are checked
in tests/test_oracle_golden.py.  TensorFlow 1.0 itself is not installable in the authoring
container, so its conv / BN kernel ARITHMETIC stays PARITY UNPINNED.  The
float64 twin is used to apportion error between "fp32 summation order" and "kernel error".

Variable naming follows the TF checkpoint scope the reference builds
(inference.py:67,73,118): ``conv{i}/weights`` (HWIO), ``conv{i}/BatchNorm/{gamma,
beta,moving_mean,moving_variance}``, ``conv/weights``, ``conv/biases``.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LEAKY_ALPHA = 0.1


def layer_table(classes, num_anchors):
    """(name, kernel, cin, cout, then) in graph order; ``then`` in
    {None,'pool','passthrough+pool'}.  Derived from inference.py:70-118."""
    t = []
    cin, ch, idx = 3, 32, 0

    def add(k, cout, then=None):
        nonlocal cin, idx
        t.append(("conv%d" % idx, k, cin, cout, then))
        cin = cout
        idx += 1

    for _ in range(2):                       # :72-76
        add(3, ch, "pool")
        ch *= 2
    for _ in range(2):                       # :77-85
        add(3, ch)
        add(1, ch // 2)
        add(3, ch, "pool")
        ch *= 2
    add(3, ch)                               # :86-94
    add(1, ch // 2)
    add(3, ch)
    add(1, ch // 2)
    add(3, ch, "passthrough+pool")           # :95-96
    ch *= 2
    add(3, ch)                               # :100-113
    add(1, ch // 2)
    add(3, ch)
    add(1, ch // 2)
    add(3, ch)
    add(3, ch)
    add(3, ch)
    pt_c = 512 * 4
    t.append(("conv%d" % idx, 3, pt_c + ch, ch, "after_concat"))   # :115-117
    t.append(("conv", 1, ch, num_anchors * (5 + classes), "linear"))  # :118
    return t


def tiny_layer_table(classes, num_anchors):
    """Same tuple format as layer_table for ``tiny()`` (inference.py:33-48); ``then`` in
    {None,'pool','pool_s1','linear'}."""
    t, cin, ch = [], 3, 16
    for _ in range(5):                       # :35-39
        t.append(("conv%d" % len(t), 3, cin, ch, "pool"))
        cin, ch = ch, ch * 2
    t.append(("conv%d" % len(t), 3, cin, ch, "pool_s1"))   # :40-42
    cin, ch = ch, ch * 2
    for _ in range(2):                       # :45-47
        t.append(("conv%d" % len(t), 3, cin, ch, None))
        cin = ch
    t.append(("conv", 1, ch, num_anchors * (5 + classes), "linear"))   # :48
    return t


def init_params(classes, num_anchors, seed=1, mode="conditioned", table=None):
    """Synthetic "random-init checkpoint" (there is no network for real weights).

    mode='xavier'      : what slim would create: Xavier-uniform weights, BN gamma=1,
                         beta=0, moving_mean=0, moving_variance=1, final bias 0.
    mode='conditioned' : He-style weights and non-trivial BN statistics so that
                         22 layers of activations stay O(1).
    Returns dict name -> float32 ndarray.
    """
    rs = np.random.RandomState(seed)
    p = {}
    for name, k, cin, cout, then in (table or layer_table(classes, num_anchors)):
        fan_in, fan_out = k * k * cin, k * k * cout
        if mode == "xavier":
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            w = rs.uniform(-lim, lim, size=(k, k, cin, cout))
        else:
            std = math.sqrt(2.0 / (1.01 * fan_in))
            if then == "linear":
                std *= 0.25                      # keep logits O(1): exp(w), exp(h) stay finite
            w = rs.normal(0.0, std, size=(k, k, cin, cout))
        p[name + "/weights"] = w.astype(np.float32)
        if then == "linear":
            b = np.zeros(cout) if mode == "xavier" else rs.normal(0, 0.1, size=cout)
            p[name + "/biases"] = b.astype(np.float32)
        else:
            if mode == "xavier":
                g, b, m, v = np.ones(cout), np.zeros(cout), np.zeros(cout), np.ones(cout)
            else:
                g = rs.uniform(0.7, 1.2, size=cout)
                b = rs.normal(0, 0.1, size=cout)
                m = rs.normal(0, 0.1, size=cout)
                v = rs.uniform(0.8, 1.3, size=cout)
            p[name + "/BatchNorm/gamma"] = g.astype(np.float32)
            p[name + "/BatchNorm/beta"] = b.astype(np.float32)
            p[name + "/BatchNorm/moving_mean"] = m.astype(np.float32)
            p[name + "/BatchNorm/moving_variance"] = v.astype(np.float32)
    return p


def leaky_oracle(x, alpha=LEAKY_ALPHA):
    return torch.maximum(x, alpha * x)


def reorg_oracle(x_nhwc, stride=2):
    """function.py:22-29: reshape [B,H/s,s,W/s,s,C] -> transpose [0,1,3,2,4,5] -> reshape."""
    is_np = isinstance(x_nhwc, np.ndarray)
    x = torch.as_tensor(x_nhwc)
    b, h, w, c = x.shape
    y = x.reshape(b, h // stride, stride, w // stride, stride, c).permute(0, 1, 3, 2, 4, 5)
    y = y.reshape(b, h // stride, w // stride, stride * stride * c).contiguous()
    return y.numpy() if is_np else y


def _conv_same(x_nchw, w_hwio):
    k = w_hwio.shape[0]
    w = w_hwio.permute(3, 2, 0, 1).contiguous()        # HWIO -> OIHW
    return F.conv2d(x_nchw, w, padding=k // 2)


def max_pool_s1_same_oracle(x_nchw):
    """slim.max_pool2d(kernel 2, stride 1, padding SAME) (tiny, inference.py:42): out[y,x] = max over rows y..min(y+1,H-1),
    columns x..min(x+1,W-1) -- the single SAME pad row/column sits at the bottom/right and never wins."""
    return F.max_pool2d(F.pad(x_nchw, (0, 1, 0, 1), value=float("-inf")), 2, 1)


def tiny_oracle(x_nhwc, params, classes, num_anchors, dtype=torch.float32, taps=None, threads=None):
    """``tiny()`` forward (inference.py:25-50), inference-mode BN.  Same conventions as darknet_oracle."""
    return darknet_oracle(x_nhwc, params, classes, num_anchors, False, dtype, taps, threads,
                          table=tiny_layer_table(classes, num_anchors))


def darknet_oracle(x_nhwc, params, classes, num_anchors, training=False, dtype=torch.float32,
                   taps=None, threads=None, table=None):
    """Forward pass.  x_nhwc [B,H,W,3]; returns [B,H/32,W/32,A*(5+C)] ndarray of `dtype`.

    taps: optional dict that receives every layer's post-activation NHWC output
    (name -> ndarray) for per-layer parity tests.
    """
    if threads:
        torch.set_num_threads(threads)
    x = torch.as_tensor(np.ascontiguousarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2).contiguous()
    P = {k: torch.as_tensor(v).to(dtype) for k, v in params.items()}
    passthrough = None
    with torch.no_grad():
        for name, k, cin, cout, then in (table or layer_table(classes, num_anchors)):
            if then == "after_concat":
                r = reorg_oracle(passthrough.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
                x = torch.cat([r, x], dim=1)                       # inference.py:116
            x = _conv_same(x, P[name + "/weights"])
            if then == "linear":
                x = x + P[name + "/biases"].view(1, -1, 1, 1)
            else:
                if training:
                    mean = x.mean(dim=(0, 2, 3))
                    var = ((x - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
                else:
                    mean = P[name + "/BatchNorm/moving_mean"]
                    var = P[name + "/BatchNorm/moving_variance"]
                inv = torch.rsqrt(var + BN_EPS) * P[name + "/BatchNorm/gamma"]
                x = x * inv.view(1, -1, 1, 1) + (P[name + "/BatchNorm/beta"] - mean * inv).view(1, -1, 1, 1)
                x = leaky_oracle(x)
            if taps is not None:
                taps[name] = x.permute(0, 2, 3, 1).contiguous().numpy()
            if then == "passthrough+pool":
                passthrough = x
            if then in ("pool", "passthrough+pool"):
                x = F.max_pool2d(x, 2, 2)
                if taps is not None:
                    taps[name + "/pool"] = x.permute(0, 2, 3, 1).contiguous().numpy()
            if then == "pool_s1":
                x = max_pool_s1_same_oracle(x)
                if taps is not None:
                    taps[name + "/pool"] = x.permute(0, 2, 3, 1).contiguous().numpy()
    return x.permute(0, 2, 3, 1).contiguous().numpy()


def conv_bn_leaky_oracle(x_nhwc, w_hwio, scale, bias, leaky=True, dtype=torch.float32):
    """One layer in "folded" form (y = conv(x)*scale + bias, optional leaky) for
    per-layer kernel tests."""
    x = torch.as_tensor(np.ascontiguousarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2)
    y = _conv_same(x, torch.as_tensor(w_hwio).to(dtype))
    y = y * torch.as_tensor(scale).to(dtype).view(1, -1, 1, 1) + torch.as_tensor(bias).to(dtype).view(1, -1, 1, 1)
    if leaky:
        y = leaky_oracle(y)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def flops_per_image(h, w, classes, num_anchors, table=None):
    """2*MAC of the 22 convs (SURVEY.md section 8a layer table)."""
    total = 0
    hh, ww = h, w
    for name, k, cin, cout, then in (table or layer_table(classes, num_anchors)):
        total += 2 * hh * ww * k * k * cin * cout
        if then in ("pool", "passthrough+pool"):
            hh //= 2
            ww //= 2
    return total
