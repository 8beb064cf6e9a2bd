"""Probe of the per-operand plane formats of the tcgen05 GEMMs (kind::f16 with A / B each f16 or bf16): error against float64
and time of one conv / one weight gradient with bf16 planes (precision 0), fp16 planes (2: activations as they are, weights
pre-scaled by a power of two) and the two mixed combinations the training step needs (3: bf16 activations x fp16 weights =
dgrad; 4: fp16 activations x bf16 weights; wgrad: fp16 x planes x bf16 dx planes).  Prints JSON lines.
    python tools/probe_fmt.py"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


def rel_after_channel_scale(a, b):
    """max-norm error left after the best per-output-channel scale is taken out (what a batch-statistics BN after the conv
    removes: a uniform shrink of a channel, e.g. the truncation bias of the accumulator, does not survive it)."""
    a = a.double().reshape(-1, a.shape[-1])
    b = b.reshape(-1, b.shape[-1])
    alpha = (a * b).sum(0) / (b * b).sum(0).clamp_min(1e-300)
    return float((a - alpha * b).abs().max() / b.abs().max()), float((alpha - 1).abs().max()), float((alpha - 1).mean())


def one(shape, mode):
    """One mode in this process (an illegal-instruction fault poisons the CUDA context)."""
    L = _lib.lib()
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    L.y2_debug_last_conv_ms.restype = ctypes.c_float
    torch.backends.cudnn.allow_tf32 = False
    b, hw, cin, k, cout = shape
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g)
    x = torch.maximum(x, 0.1 * x)
    w = torch.randn(k, k, cin, cout, device="cuda", generator=g) * (2.0 / (1.01 * k * k * cin)) ** 0.5
    if mode.startswith("precision_"):
        prec = int(mode.split("_")[1])
        ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
        y = torch.full((b, hw, hw, cout), float("nan"), device="cuda")
        rc = L.y2_conv2d(_lib.ptr(x), b, hw, hw, cin, _lib.ptr(w), k, cout, None, None, 0, _lib.ptr(y), prec, 0, 0, None)
        if rc != 0:
            return {"rc": rc, "error": L.y2_last_error().decode()[:200]}
        torch.cuda.synchronize()
        noise, shrink_max, shrink_mean = rel_after_channel_scale(y, ref)
        return {"rc": rc, "rel_err_vs_fp64": rel(y, ref), "noise_after_channel_scale": noise, "channel_scale_minus_1_max": shrink_max,
                "channel_scale_minus_1_mean": shrink_mean, "ms": float(L.y2_debug_last_conv_ms())}
    fmt = int(mode.split("_")[2])
    dy = torch.randn(b, hw, hw, cout, device="cuda", generator=g) * 1e-4
    w0 = torch.zeros(cout, cin, k, k, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(x.double().permute(0, 3, 1, 2), w0, padding=k // 2).backward(dy.double().permute(0, 3, 1, 2))
    gref = w0.grad.permute(2, 3, 1, 0)
    L.y2_debug_set(10, float(fmt))
    dw = torch.full((k, k, cin, cout), float("nan"), device="cuda")
    rc = L.y2_conv2d_wgrad(_lib.ptr(x), b, hw, hw, cin, _lib.ptr(dy), k, cout, _lib.ptr(dw), 0, None)
    if rc != 0:
        return {"rc": rc, "error": L.y2_last_error().decode()[:200]}
    torch.cuda.synchronize()
    return {"rc": rc, "rel_err_vs_fp64": rel(dw, gref)}


SHAPES = ((8, 26, 256, 3, 512), (4, 13, 1024, 1, 512), (32, 13, 1024, 3, 1024), (2, 13, 3072, 3, 1024), (2, 104, 64, 3, 128))
MODES = ("precision_0", "precision_2") + (("precision_3", "precision_4", "wgrad_fmt_0", "wgrad_fmt_3") if os.environ.get("Y2_PROBE_MIXED") else ())


def main():
    import subprocess
    if len(sys.argv) > 2:                                   # child: shape index, mode
        print("RESULT " + json.dumps(one(SHAPES[int(sys.argv[1])], sys.argv[2])), flush=True)
        return
    out = []
    dead = set()
    for i, shape in enumerate(SHAPES):
        row = {"conv": list(shape)}
        for mode in MODES:
            if mode in dead:
                row[mode] = {"skipped": "faulted on an earlier shape"}
                continue
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i), mode], capture_output=True, text=True, timeout=300)
            hit = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if hit:
                row[mode] = json.loads(hit[-1][7:])
            else:
                row[mode] = {"fault": (r.stderr.strip().splitlines() or ["?"])[-1][:200]}
                dead.add(mode)
        print(json.dumps(row), flush=True)
        out.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_fmt.json", "w"), indent=1)


if __name__ == "__main__":
    main()
