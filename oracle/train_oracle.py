"""CPU oracle for the TRAINING step (BASELINE config 3): forward with batch-statistics BN, the 4-part
loss, and backward through everything.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs).

Restates with torch autograd (float64 by default = "truth"; float32 available):

* forward      <- ``darknet(..., training=True)``   model/yolo2/inference.py:61-120 with
                  ``slim.batch_norm(is_training=True, decay=0.999, epsilon=1e-5)``: batch mean and
                  POPULATION variance over (B,H,W), gradients flowing through both
* loss         <- ``Objectives`` model/yolo2/__init__.py:62-94 + hparam weighting :114-119
* backward     <- what ``slim.learning.create_train_op`` (train.py:127-129) gets from tf.gradients
* moving stats <- UPDATE_OPS of slim.batch_norm: m <- m*decay + batch*(1-decay)

PINNED TO THE REFERENCE'S SOURCE, NOT TO TENSORFLOW'S ARITHMETIC: tests/golden/train_reference.npz holds one
training step of the reference's own ``darknet(training=True)`` + ``Model`` + ``Objectives`` source run with
torch float64 stand-ins for the slim / tf calls and autograd for tf.gradients (tests/golden/make_train_golden.py);
loss, objectives, output, d(total)/d(net) and the gradients of all 65 trainable variables match
(tests/test_train_reference_golden.py).  TensorFlow itself is not installable here: its kernel arithmetic stays
PARITY UNPINNED.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .darknet_oracle import BN_EPS, LEAKY_ALPHA, layer_table

BN_DECAY = 0.999


def train_step_oracle(x_nhwc, params, classes, anchors, labels, hparam, dtype=torch.float64, taps=None, device="cpu", taps_numpy=True, table=None):
    """Returns dict(objectives, total, grads {variable name -> ndarray}, net, new_moving {name -> ndarray},
    dnet = d total / d net).  Variable names as in oracle/darknet_oracle.py (no scope prefix).
    ``device``: where torch evaluates this same restatement.  "cpu" is the oracle of record; "cuda" (float64,
    TF32 off) is used by the -m gpu tests only for BASELINE config 3 at its full size (B = 64, 416 x 416), where the
    float64 step is ~7 TFLOP and tens of GB -- minutes on the host cores, seconds on the device.
    ``table``: layer table (default: darknet's ``layer_table``); pass ``tiny_layer_table(classes, A)`` for ``tiny()``
    (model/yolo2/inference.py:25-50), whose 2x2 stride-1 SAME max-pool (:42) pads bottom/right only and ignores the padding."""
    anchors = np.asarray(anchors, dtype=np.float64)
    A = len(anchors)
    P = {k: torch.tensor(np.asarray(v), dtype=dtype, device=device, requires_grad=("moving" not in k)) for k, v in params.items()}
    x = torch.tensor(np.ascontiguousarray(x_nhwc), dtype=dtype, device=device).permute(0, 3, 1, 2)
    new_moving = {}
    passthrough = None
    for name, k, cin, cout, then in (table if table is not None else layer_table(classes, A)):
        if then == "after_concat":
            b, c, h, w = passthrough.shape
            r = passthrough.permute(0, 2, 3, 1).reshape(b, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 2, 4, 5)
            r = r.reshape(b, h // 2, w // 2, 4 * c).permute(0, 3, 1, 2)
            x = torch.cat([r, x], dim=1)
        wt = P[name + "/weights"].permute(3, 2, 0, 1)
        x = F.conv2d(x, wt, padding=k // 2)
        if then == "linear":
            x = x + P[name + "/biases"].view(1, -1, 1, 1)
        else:
            mean = x.mean(dim=(0, 2, 3))
            var = ((x - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
            inv = torch.rsqrt(var + BN_EPS) * P[name + "/BatchNorm/gamma"]
            x = x * inv.view(1, -1, 1, 1) + (P[name + "/BatchNorm/beta"] - mean * inv).view(1, -1, 1, 1)
            x = torch.maximum(x, LEAKY_ALPHA * x)
            new_moving[name + "/BatchNorm/moving_mean"] = (P[name + "/BatchNorm/moving_mean"] * BN_DECAY + mean.detach() * (1 - BN_DECAY)).cpu().numpy()
            new_moving[name + "/BatchNorm/moving_variance"] = (P[name + "/BatchNorm/moving_variance"] * BN_DECAY + var.detach() * (1 - BN_DECAY)).cpu().numpy()
        if taps is not None:
            taps[name] = x.detach().permute(0, 2, 3, 1).contiguous().cpu().numpy() if taps_numpy else x.detach().permute(0, 2, 3, 1)
        if then == "passthrough+pool":
            passthrough = x
        if then in ("pool", "passthrough+pool"):
            x = F.max_pool2d(x, 2, 2)
        elif then == "pool_s1":
            x = F.max_pool2d(F.pad(x, (0, 1, 0, 1), value=float("-inf")), 2, 1)
    net = x.permute(0, 2, 3, 1).contiguous()
    net.retain_grad()
    b, hc, wc, _ = net.shape
    cells = hc * wc
    inp = net.reshape(b, cells, A, 5 + classes)
    anc = torch.tensor(anchors, dtype=dtype, device=device)
    sig = torch.sigmoid(inp[..., :3])
    iou_p, oxy = sig[..., 0], sig[..., 1:3]
    wh = torch.exp(inp[..., 3:5]) * anc.reshape(1, 1, A, 2)
    prob_p = torch.softmax(inp[..., 5:], -1)
    areas_p = wh[..., 0] * wh[..., 1]
    omin, omax = oxy - wh / 2, oxy + wh / 2
    wh01s = torch.sqrt(wh / torch.tensor([wc, hc], dtype=dtype, device=device).reshape(1, 1, 1, 2))
    coords_p = torch.cat([oxy, wh01s], -1)
    mask, prob, coords, tmin, tmax, areas = [torch.tensor(np.asarray(t), dtype=dtype, device=device) for t in labels]
    lo, hi = torch.maximum(omin, tmin), torch.minimum(omax, tmax)
    iwh = torch.clamp(hi - lo, min=0.0)
    inter = iwh[..., 0] * iwh[..., 1]
    iou = inter / torch.clamp(areas + areas_p - inter, min=1e-10)
    best = (iou == iou.max(2, keepdim=True).values).to(dtype).detach()
    mb = mask * best
    mn = 1 - mb
    cnt = float(b * cells * A)
    obj = {
        "iou_best": (mb * (iou_p - mb) ** 2).sum() / cnt,
        "iou_normal": (mn * (iou_p - mb) ** 2).sum() / cnt,
        "coords": (mb[..., None] * (coords_p - coords) ** 2).sum() / cnt,
        "prob": (mb[..., None] * (prob_p - prob) ** 2).sum() / cnt,
    }
    total = sum(obj[k] * float(hparam[k]) for k in obj)
    total.backward()
    grads = {k: v.grad.cpu().numpy() for k, v in P.items() if v.requires_grad and v.grad is not None}
    return {"objectives": {k: float(v.detach()) for k, v in obj.items()}, "total": float(total.detach()),
            "grads": grads, "net": net.detach().cpu().numpy(), "dnet": net.grad.cpu().numpy(), "new_moving": new_moving}
