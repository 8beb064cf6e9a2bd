"""Is the mid-layer conv limited by L2->SM operand traffic, and does K-alignment of concurrent CTAs (identical operand
tiles requested at the same time, deduplicated by L2) relieve it?  conv13 shape (3x3, 512 -> 1024, 13x13) at batches
that give exactly 1 or 2 full waves of 148 tiles, data-parallel (all CTAs at the same k-block) vs stream-K (every CTA at
a different K phase), parity mode and single-pass mode.  Diagnostic tool; writes gpurun_out/probe_align.json."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402

L = _lib.lib()
L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
L.y2_debug_last_conv_ms.restype = ctypes.c_float
out = []
for (B, hw, cin, cout, k) in [(28, 13, 512, 1024, 3), (56, 13, 512, 1024, 3), (28, 13, 1024, 1024, 3), (7, 26, 256, 512, 3)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, hw, hw, cin, device="cuda", generator=g)
    w = torch.randn(k, k, cin, cout, device="cuda", generator=g) / (k * k * cin) ** 0.5
    y = torch.empty(B, hw, hw, cout, device="cuda")
    for prec in (0, 1):
        for mode, name in ((1, "dp"), (2, "sk")):
            L.y2_debug_set(0, float(mode))
            ts = []
            for rep in range(3):
                rc = L.y2_conv2d(_lib.ptr(x), B, hw, hw, cin, _lib.ptr(w), k, cout, None, None, 0, _lib.ptr(y), prec, 0, 0, None)
                torch.cuda.synchronize()
                assert rc == 0, L.y2_last_error()
                ts.append(float(L.y2_debug_last_conv_ms()))
            L.y2_debug_set(0, 0.0)
            flops = 2.0 * B * hw * hw * k * k * cin * cout
            r = {"shape": [B, hw, cin, cout, k], "precision": prec, "mode": name, "ms": min(ts), "algorithmic_tflops": flops / min(ts) / 1e9,
                 "mma_tflops": flops * (3 if prec == 0 else 1) / min(ts) / 1e9}
            out.append(r)
            print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_align.json", "w"), indent=1)
