"""EXPERIMENTAL: the whole Darknet-19 (inference) driven layer by layer from Python through the probe kernels, to measure the END-TO-END
error of the candidate precision mode on the hardware before any integration work:
    mix    : y2_mix_split (scales from the one-layer bound S * amax_in + T) + y2_conv2d_mix_pre per layer (fp16 + 2 x e4m3, kcap 32 / 16)
    bf16x3 : the shipped kernel through y2_conv2d per layer (same harness, same glue)
against the same network in float64 (torch).  conv0 / conv1 (3 and 32 input channels: below the probe kernel's 64-channel k-block)
run in float64 in all three; pools / reorg / concat are torch ops on the float32 layer outputs.  Model: model/yolo2/inference.py:61-120
with conditioned random weights.  Written without GPU time; part of the first call of the next round.  Writes gpurun_out/probe_mix_network.json."""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402
from yolo_tf_b200.model.yolo2.inference import layer_geometry  # noqa: E402

DRY = os.environ.get("Y2_PROBE_DRYRUN") == "1"       # CPU dry run of the harness glue (shapes, bounds, concat): no library calls
L = None if DRY else _lib.lib()
dev = "cpu" if DRY else "cuda"
B, SIZE, C, A = (1, 64, 20, 5) if DRY else (2, 416, 80, 5)
rs = np.random.RandomState(1)
layers = []
for name, k, cin, cout, has_bn, pool in layer_geometry(C, A):
    std = math.sqrt(2.0 / (1.01 * k * k * cin)) * (1.0 if has_bn else 0.25)
    w = torch.from_numpy(rs.normal(0.0, std, size=(k, k, cin, cout)).astype(np.float32)).to(dev)
    if has_bn:
        g, b = rs.uniform(0.7, 1.2, size=cout), rs.normal(0, 0.1, size=cout)
        m, v = rs.normal(0, 0.1, size=cout), rs.uniform(0.8, 1.3, size=cout)
        inv = g / np.sqrt(v + 1e-5)
        scale, bias = inv, b - m * inv
    else:
        scale, bias = np.ones(cout), rs.normal(0, 0.1, size=cout)
    layers.append((name, k, cin, cout, has_bn, pool, w, torch.from_numpy(scale.astype(np.float32)).to(dev), torch.from_numpy(bias.astype(np.float32)).to(dev)))
x0 = torch.from_numpy(np.random.RandomState(2).normal(0, 1, size=(B, SIZE, SIZE, 3)).astype(np.float32)).to(dev)


def reorg(t):                                   # model/yolo2/function.py:22-29
    b, h, w, c = t.shape
    return t.reshape(b, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h // 2, w // 2, 4 * c)


def conv_ref(x, w, scale, bias, leaky, dtype):
    y = F.conv2d(x.to(dtype).permute(0, 3, 1, 2), w.to(dtype).permute(3, 2, 0, 1), padding=w.shape[0] // 2).permute(0, 2, 3, 1)
    y = y * scale.to(dtype) + bias.to(dtype)
    return torch.maximum(y, 0.1 * y) if leaky else y


def conv_bf16x3(x, w, scale, bias, leaky):
    if DRY:
        return conv_ref(x, w, scale, bias, leaky, torch.float32)
    b, h, wd, cin = x.shape
    y = torch.empty(b, h, wd, w.shape[3], device=dev)
    _lib.check(L.y2_conv2d(_lib.ptr(x.contiguous()), b, h, wd, cin, _lib.ptr(w), w.shape[0], w.shape[3], _lib.ptr(scale), _lib.ptr(bias), int(leaky),
                           _lib.ptr(y), 0, 0, 0, None))
    return y


def make_conv_mix(kcap, stats):
    def conv_mix(x, w, scale, bias, leaky):
        b, h, wd, cin = x.shape
        x = x.contiguous()
        n = x.numel()
        amax_in = float(x.abs().max())
        x16 = torch.empty(n, dtype=torch.float16, device=dev)
        x8, rx8 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
        # the scales a producer would have used for this tensor: from ITS one-layer bound, carried along by the caller
        bound = stats.get("bound", amax_in)
        assert bound >= amax_in * (1 - 1e-6), (bound, amax_in)                     # the bound is a bound
        if DRY:
            y = conv_ref(x, w, scale, bias, leaky, torch.float32)
            stats["bound"] = float((scale.abs() * w.abs().sum(dim=(0, 1, 2))).max()) * amax_in + float(bias.abs().max())
            stats.setdefault("looseness", []).append(stats["bound"] / float(y.abs().max()))
            return y
        _lib.check(L.y2_mix_split(_lib.ptr(x), n, bound, _lib.ptr(x16), _lib.ptr(x8), _lib.ptr(rx8), None))
        y = torch.empty(b, h, wd, w.shape[3], device=dev)
        _lib.check(L.y2_conv2d_mix_pre(_lib.ptr(x16), _lib.ptr(x8), _lib.ptr(rx8), bound, b, h, wd, cin, _lib.ptr(w), w.shape[0], w.shape[3], _lib.ptr(scale),
                                       _lib.ptr(bias), int(leaky), _lib.ptr(y), None, None, None, 0.0, None, 7, kcap, 0, None))
        # bound of what was just produced (S * amax_in + T), used when the NEXT layer splits it
        stats["bound"] = float((scale.abs() * w.abs().sum(dim=(0, 1, 2))).max()) * amax_in + float(bias.abs().max())
        stats.setdefault("looseness", []).append(stats["bound"] / float(y.abs().max()))
        return y
    return conv_mix


def run(conv, dtype_glue=torch.float32, stats=None):
    x = x0.to(torch.float64)
    tap = None
    for i, (name, k, cin, cout, has_bn, pool, w, scale, bias) in enumerate(layers):
        if name == "conv20":
            if stats is not None:
                stats["bound"] = max(stats["bound"], stats["tap_bound"])           # one pair of scales for the concat buffer
            x = torch.cat([reorg(tap), x], dim=3)
        if i < 2:
            x = conv_ref(x, w, scale, bias, has_bn, torch.float64).to(dtype_glue)  # conv0 / conv1: outside the probe kernel's shapes
            if stats is not None:
                stats["bound"] = float(x.abs().max())
        else:
            x = conv(x, w, scale, bias, has_bn)
        if name == "conv12":
            tap = x
            if stats is not None:
                stats["tap_bound"] = stats["bound"]
        if pool:
            x = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    if not DRY:
        torch.cuda.synchronize()
    return x


ref = run(lambda x, w, s, b, l: conv_ref(x, w, s, b, l, torch.float64), dtype_glue=torch.float64)
rel = lambda y: float((y.double() - ref).abs().max() / ref.abs().max())
out = {"config": {"batch": B, "size": SIZE, "classes": C}}
out["torch_fp32"] = rel(run(lambda x, w, s, b, l: conv_ref(x, w, s, b, l, torch.float32)))
out["bf16x3_kcap32"] = rel(run(conv_bf16x3))
for kcap in (32, 16, 0):
    st = {}
    out["mix_kcap%d" % kcap] = rel(run(make_conv_mix(kcap, st), stats=st))
    out["mix_kcap%d_looseness_log2_max" % kcap] = math.log2(max(st["looseness"]))
if not DRY:
    _lib.check(L.y2_check_async_errors())
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_mix_network.json", "w"), indent=1)
