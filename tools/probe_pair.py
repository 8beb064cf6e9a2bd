"""CTA-pair (cta_group::2) vs single-CTA conv main loop at exactly one tile per SM / one pair tile per TPC: 3x3, 512 -> 512,
13x13, batch 56 (74 x 2 tiles of 128 x 256 = 37 x 2 pair tiles of 256 x 256, K = 4608 = 72 k-blocks), data-parallel
schedule, parity (3 MMAs per product) and single-pass modes; plus the conv18 and conv20 shapes of the bench batch.
Diagnostic tool; writes gpurun_out/probe_pair.json."""
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
for (B, hw, cin, cout, k, sched) in [(56, 13, 512, 512, 3, 1), (32, 13, 1024, 1024, 3, 0), (32, 13, 3072, 1024, 3, 0), (32, 26, 256, 512, 3, 0),
                                     (32, 13, 1024, 512, 1, 0), (32, 52, 128, 256, 3, 0)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, hw, hw, cin, device="cuda", generator=g)
    w = torch.randn(k, k, cin, cout, device="cuda", generator=g) / (k * k * cin) ** 0.5
    y = torch.empty(B, hw, hw, cout, device="cuda")
    for prec in (0, 1):
        for pair in (0, 1):
            L.y2_debug_set(0, float(sched))
            L.y2_debug_set(7, float(pair))
            ts = []
            for rep in range(5):
                rc = L.y2_conv2d(_lib.ptr(x), B, hw, hw, cin, _lib.ptr(w), k, cout, None, None, 0, _lib.ptr(y), prec, 0, 0, None)
                torch.cuda.synchronize()
                assert rc == 0, L.y2_last_error()
                ts.append(float(L.y2_debug_last_conv_ms()))
            L.y2_debug_set(0, 0.0)
            L.y2_debug_set(7, 0.0)
            flops = 2.0 * B * hw * hw * k * k * cin * cout
            r = {"shape": [B, hw, cin, cout, k], "precision": prec, "pair": pair, "sched": sched, "ms": min(ts),
                 "algorithmic_tflops": flops / min(ts) / 1e9, "mma_tflops": flops * (3 if prec == 0 else 1) / min(ts) / 1e9}
            out.append(r)
            print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_pair.json", "w"), indent=1)
