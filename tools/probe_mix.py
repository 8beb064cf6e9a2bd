"""EXPERIMENTAL mixed-kind conv (csrc/y2_conv_mix.cu, y2_conv2d_mix) against the shipped kernel, one layer shape at a time:
  * does kind::f16 + kind::f8f6f4 accumulate correctly into one TMEM tile (error vs float64 for terms = 7, 1, 6),
  * what the main loop costs with 1 fp16 + 2 fp8 products (terms 7; with the 32-k-block chain cap and without) vs the fp16
    product alone (1) vs the two fp8 products alone (6)
    vs the shipped bf16x3 kernel in its data-parallel single-CTA configuration and in its default configuration.
Written without GPU time (round 1 budget spent); first thing to run in round 2.  Writes gpurun_out/probe_mix.json."""
import ctypes
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402

DRY = os.environ.get("Y2_PROBE_DRYRUN") == "1"       # CPU dry run of this script's own control flow: no library, no CUDA
if DRY:
    class _Fake(object):
        def __getattr__(self, name):
            return (lambda *a: 1.0) if name.endswith("_ms") else (lambda *a: 0)
    L = _Fake()
    _lib.ptr = lambda t, dtype=None: None
    _lib.check = lambda rc: None
    _cuda_sync = torch.cuda.synchronize
    torch.cuda.synchronize = lambda: None
    DEV = "cpu"
else:
    L = _lib.lib()
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    L.y2_debug_last_conv_ms.restype = ctypes.c_float
    DEV = "cuda"
out = []
# one tile per SM (conv13's shape at batch 28 -> 148 tiles of 128 x 256), then conv8 / conv18 / conv20 / conv14 at the bench batch
SHAPES = [(28, 13, 512, 1024, 3), (32, 26, 256, 512, 3), (32, 13, 1024, 1024, 3), (32, 13, 3072, 1024, 3), (32, 13, 1024, 512, 1)]
if DRY:
    SHAPES = [(1, 13, 64, 64, 3), (1, 13, 64, 32, 1)]
for (B, hw, cin, cout, k) in SHAPES:
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(B, hw, hw, cin, device=DEV, generator=g)
    x = torch.maximum(x, 0.1 * x)
    w = torch.randn(k, k, cin, cout, device=DEV, generator=g) * (2.0 / (k * k * cin)) ** 0.5
    y = torch.empty(B, hw, hw, cout, device=DEV)
    ref = torch.zeros(B, hw, hw, cout, dtype=torch.float64, device=DEV)
    for i0 in range(0, B, 4):
        ref[i0:i0 + 4] = F.conv2d(x[i0:i0 + 4].double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
    flops = 2.0 * B * hw * hw * k * k * cin * cout
    row = {"shape": [B, hw, cin, cout, k], "kblocks": k * k * cin // 64}
    for terms, kcap in ((7, 32), (7, 16), (7, 0), (1, 0), (6, 0)):
        ts = []
        for rep in range(5):
            y.fill_(float("nan"))
            rc = L.y2_conv2d_mix(_lib.ptr(x), B, hw, hw, cin, _lib.ptr(w), k, cout, None, None, 0, _lib.ptr(y), terms, kcap, 0, None)
            torch.cuda.synchronize()
            assert rc == 0, L.y2_last_error()
            ts.append(float(L.y2_debug_last_mix_ms()))
        err = float((y.double() - ref).abs().max() / ref.abs().max())
        row["mix_terms%d_kcap%d" % (terms, kcap)] = {"ms": min(ts), "algorithmic_tflops": flops / min(ts) / 1e9, "rel_err_vs_fp64": err}
    for name, sched, pair in (("bf16x3_dp_single", 1, 0), ("bf16x3_default", 0, 1)):
        L.y2_debug_set(0, float(sched))
        L.y2_debug_set(7, float(pair))
        ts = []
        for rep in range(5):
            rc = L.y2_conv2d(_lib.ptr(x), B, hw, hw, cin, _lib.ptr(w), k, cout, None, None, 0, _lib.ptr(y), 0, 0, 0, None)
            torch.cuda.synchronize()
            assert rc == 0, L.y2_last_error()
            ts.append(float(L.y2_debug_last_conv_ms()))
        L.y2_debug_set(0, 0.0)
        L.y2_debug_set(7, 0.0)
        row[name] = {"ms": min(ts), "algorithmic_tflops": flops / min(ts) / 1e9,
                     "rel_err_vs_fp64": float((y.double() - ref).abs().max() / ref.abs().max())}
    out.append(row)
    print(json.dumps(row), flush=True)
# ---- epilogue cost of the storage format: float32 output only vs float32 + split triple + amax vs split triple only
epi = []
for (B, hw, cin, cout, k) in (SHAPES[:1] if DRY else [(28, 13, 512, 1024, 3), (32, 26, 256, 512, 3), (32, 13, 1024, 512, 1)]):
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(B, hw, hw, cin, device=DEV, generator=g)
    w = torch.randn(k, k, cin, cout, device=DEV, generator=g) * (2.0 / (k * k * cin)) ** 0.5
    n0, n1 = x.numel(), B * hw * hw * cout
    amax_x = float(x.abs().max())
    bound = float(w.abs().sum(dim=(0, 1, 2)).max()) * amax_x
    x16 = torch.empty(n0, dtype=torch.float16, device=DEV)
    x8, rx8 = torch.empty(n0, dtype=torch.uint8, device=DEV), torch.empty(n0, dtype=torch.uint8, device=DEV)
    _lib.check(L.y2_mix_split(_lib.ptr(x), n0, amax_x, _lib.ptr(x16), _lib.ptr(x8), _lib.ptr(rx8), None))
    y = torch.empty(B, hw, hw, cout, device=DEV)
    o16 = torch.empty(n1, dtype=torch.float16, device=DEV)
    o8, or8 = torch.empty(n1, dtype=torch.uint8, device=DEV), torch.empty(n1, dtype=torch.uint8, device=DEV)
    amax = torch.zeros(1, dtype=torch.int32, device=DEV)
    row = {"shape": [B, hw, cin, cout, k]}
    for name, yy, split in (("f32_only", y, False), ("f32_and_split", y, True), ("split_only", None, True)):
        ts = []
        for rep in range(5):
            rc = L.y2_conv2d_mix_pre(_lib.ptr(x16), _lib.ptr(x8), _lib.ptr(rx8), amax_x, B, hw, hw, cin, _lib.ptr(w), k, cout, None, None, 1,
                                     _lib.ptr(yy), _lib.ptr(o16) if split else None, _lib.ptr(o8) if split else None, _lib.ptr(or8) if split else None,
                                     bound if split else 0.0, _lib.ptr(amax) if split else None, 7, 0, 0, None)
            torch.cuda.synchronize()
            assert rc == 0, L.y2_last_error()
            ts.append(float(L.y2_debug_last_mix_ms()))
        row[name + "_ms"] = min(ts)
    row["bound_over_amax"] = bound / float(y.abs().max())
    epi.append(row)
    print(json.dumps(row), flush=True)
_lib.check(L.y2_check_async_errors())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(epi, open("gpurun_out/probe_mix_epilogue.json", "w"), indent=1)
json.dump(out, open("gpurun_out/probe_mix.json", "w"), indent=1)
