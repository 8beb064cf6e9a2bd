"""Weight-gradient GEMM (wgrad_tc_kernel) at the TRAINING batch size (B=64) against a float64 reference on the same device:
how much do the long tensor-core accumulation chains (K = pixels) cost here?  Shapes of conv2, conv8, conv13, conv18.
Inputs: x ~ N(0,1); dy ~ N(0,1) (zero-mean products) and |dy| (same-sign sums: the worst case for a truncating accumulator).
Diagnostic tool; writes gpurun_out/diag_wgrad.json."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402

out = []
for (b, hw, cin, k, cout) in [(64, 104, 64, 3, 128), (64, 26, 256, 3, 512), (64, 13, 512, 3, 1024), (64, 13, 1024, 3, 1024), (4, 104, 64, 3, 128)]:
    g = torch.Generator(device="cuda").manual_seed(hw + cin)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g)
    for kind in ("normal", "abs"):
        dy = torch.randn(b, hw, hw, cout, device="cuda", generator=g)
        xx = x
        if kind == "abs":
            dy, xx = dy.abs(), x.abs()
        dw = torch.full((k, k, cin, cout), float("nan"), device="cuda")
        _lib.check(_lib.lib().y2_conv2d_wgrad(_lib.ptr(xx.contiguous()), b, hw, hw, cin, _lib.ptr(dy), k, cout, _lib.ptr(dw), 0, None))
        torch.cuda.synchronize()
        ref = torch.zeros(k, k, cin, cout, dtype=torch.float64, device="cuda")
        for i0 in range(0, b, 8):                      # float64 reference in batch slices (memory)
            xd = xx[i0:i0 + 8].double().permute(0, 3, 1, 2)
            w0 = torch.zeros(cout, cin, k, k, dtype=torch.float64, device="cuda", requires_grad=True)
            F.conv2d(xd, w0, padding=k // 2).backward(dy[i0:i0 + 8].double().permute(0, 3, 1, 2))
            ref += w0.grad.permute(2, 3, 1, 0)
        err = float((dw.double() - ref).abs().max() / ref.abs().max())
        bias = float(((dw.double() - ref) * ref.sign()).mean() / ref.abs().mean())
        r = {"shape": [b, hw, cin, k, cout], "inputs": kind, "rel_err_max": err, "mean_signed_rel_err": bias}
        out.append(r)
        print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/diag_wgrad.json", "w"), indent=1)
