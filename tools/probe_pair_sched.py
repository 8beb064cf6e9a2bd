"""Per-CTA phase timing (globaltimer stamps) of the stream-K layers in CTA-pair mode: where do conv8 / conv13 / conv18 lose the
~25 % of the tensor pipe ncu reports?  Diagnostic; writes gpurun_out/probe_pair_sched.json."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402

L = _lib.lib()
raw = ctypes.CDLL(_lib.LIB_PATH)
raw.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
raw.y2_debug_last_conv_ms.restype = ctypes.c_float
raw.y2_debug_cta_times.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
SHAPES = {"conv8": (32, 26, 256, 3, 512), "conv13": (32, 13, 512, 3, 1024), "conv18": (32, 13, 1024, 3, 1024), "conv5": (32, 52, 128, 3, 256)}
if os.environ.get("PROBE_SHAPES"):
    SHAPES = {k: v for k, v in SHAPES.items() if k in os.environ["PROBE_SHAPES"].split(",")}
out = []
raw.y2_debug_set(2, 1.0)
for kv in sys.argv[1:]:                      # extra y2_debug_set settings, e.g. 6=0 (no TMA-store epilogue)
    k_, v_ = kv.split("=")
    raw.y2_debug_set(int(k_), float(v_))
for name, (b, hw, cin, k, cout) in SHAPES.items():
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g)
    w = torch.randn(k, k, cin, cout, device="cuda", generator=g) * 0.02
    y = torch.empty(b, hw, hw, cout, device="cuda")
    for pair in (0, 1):
        for mode, ov in {"default": 0, "dp": 1}.items():
            for kcap in (32, 0):
                raw.y2_debug_set(0, float(ov))
                raw.y2_debug_set(7, float(pair))
                raw.y2_debug_set(8, float(kcap))
                best = 1e9
                for rep in range(3):
                    _lib.check(L.y2_conv2d(_lib.ptr(x), b, hw, hw, cin, _lib.ptr(w), k, cout, None, None, 1, _lib.ptr(y), 0, 0, 0, None))
                    best = min(best, raw.y2_debug_last_conv_ms())
                buf = (ctypes.c_ulonglong * (1024 * 4))()
                sched = (ctypes.c_int * 4)()
                n = raw.y2_debug_cta_times(buf, 1024, sched)
                t = np.array(buf[:n * 4], dtype=np.float64).reshape(n, 4)
                t0 = t[:, 0].min()
                flops = 2.0 * b * hw * hw * k * k * cin * cout
                rec = {"layer": name, "pair": pair, "mode": mode, "kcap": kcap, "ms": best, "algorithmic_tflops": flops / best / 1e9,
                       "dp_tiles": sched[0], "sk_workers": sched[1], "grid": sched[2], "KB": sched[3],
                       "start_us_max": float((t[:, 0] - t0).max() / 1e3),
                       "end_us_min": float((t[:, 3] - t0).min() / 1e3), "end_us_mean": float((t[:, 3] - t0).mean() / 1e3),
                       "end_us_max": float((t[:, 3] - t0).max() / 1e3)}
                fin = t[:, 2] > 0
                if fin.any():
                    rec["heads"] = int(fin.sum())
                    rec["head_wait_start_us_mean"] = float((t[fin, 2] - t0).mean() / 1e3)
                    rec["head_wait_plus_epilogue_us_mean"] = float((t[fin, 3] - t[fin, 2]).mean() / 1e3)
                    rec["head_wait_plus_epilogue_us_max"] = float((t[fin, 3] - t[fin, 2]).max() / 1e3)
                out.append(rec)
                print(json.dumps(rec), flush=True)
raw.y2_debug_set(0, 0.0)
raw.y2_debug_set(7, 0.0)
raw.y2_debug_set(8, 32.0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_pair_sched.json", "w"), indent=1)
