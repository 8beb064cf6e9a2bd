"""CPU (-m "not gpu"): feasibility study of the precision mode PLANNED for the tcgen05 convs ("next") --
NOT a test of shipped kernel code.  The shipped parity mode spends 3 bf16 MMAs per product (hi*hi + hi*lo + lo*hi); the
tensor pipe is 89 % busy with them, so the only step change left is fewer MMA cycles per product.  Candidate:

    x*w  ~=  X16*W16  +  2^-s * ( X8*RW8 + RX8*W8 )          1 kind::f16 MMA + 2 kind::f8f6f4 MMAs (each half the cycles)
    X16 = fp16(x*2^e16)      RX8 = e4m3((x - X16/2^e16) * 2^(e8+12))      X8 = e4m3(x*2^e8)       (weights likewise, static)

i.e. 2.0 MMA-equivalents and 4 + 4 bytes of shared-memory reads per product instead of 3.0 and 6 + 6, same 4 bytes per
element in HBM.  fp16 / e4m3 lack bf16's range, so every tensor carries two power-of-two scales.  They are derived on the
fly, without a host round trip or a calibration pass: the producer of a tensor publishes its true amax (an atomicMax in
its epilogue), and a layer scales its OUTPUT by the one-layer bound  S * amax_in + T  with the static
S = max_n |bn_scale_n| * sum_k |w_nk|,  T = max_n |bn_bias_n|  (typically 2^5.5, at most 2^8.7 above the true amax).

This file emulates that arithmetic exactly (quantisers = torch's fp16 / float8_e4m3fn casts, products exact, sums in
float64) over the whole 22-layer network and pins the claim the plan rests on: end to end it stays inside the 1e-4 parity
bar with a factor 2 to spare -- for O(1) activations AND for the random-init checkpoint of BASELINE config 1 whose
activations decay by 10^4 through the network, at input scales 10^-3 .. 10^3 -- while fp16 alone, or fixed scales, do not.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import darknet_oracle as D


def q16(x):
    return x.to(torch.float32).to(torch.float16).to(torch.float64)


def q8(x):
    return x.to(torch.float32).clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float64)      # cvt.rn.satfinite.e4m3


def conv64(x, w_hwio):
    k = w_hwio.shape[0]
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1).contiguous().double(), padding=k // 2)


def pow2_floor(v):
    return 2.0 ** math.floor(math.log2(v))


class Emulator(object):
    def __init__(self, params, classes=20, anchors=5, dynamic=True, corrections=True):
        self.P = {k: torch.as_tensor(v).float() for k, v in params.items()}
        self.table = D.layer_table(classes, anchors)
        self.dynamic, self.corrections = dynamic, corrections
        self.looseness = []

    def store(self, x, bound):
        """What an epilogue leaves in HBM for the next layer; bound >= amax(x) fixes the two scales."""
        if not self.dynamic:
            bound = 128.0                                                           # fixed scales: fine for O(1) tensors only
        e16, e8 = pow2_floor(2.0 ** 15 / bound), pow2_floor(2.0 ** 8 / bound)
        x16 = q16(x * e16)
        r = x.double() - x16 / e16
        return dict(x=x, x16=x16, x8=q8(x * e8), rx8=q8(r * (e8 * 4096.0)), e16=e16, e8=e8, amax=float(x.abs().max()), bound=bound)

    def conv(self, s, w):
        wmax = float(w.abs().max())
        f16, f8 = pow2_floor(2.0 ** 14 / wmax), pow2_floor(2.0 ** 8 / wmax)
        w16 = q16(w * f16)
        main = conv64(s["x16"], w16) / (s["e16"] * f16)
        if not self.corrections:
            return main.float()
        rw = w.double() - w16 / f16
        corr = conv64(s["x8"], q8(rw * (f8 * 4096.0))) + conv64(s["rx8"], q8(w * f8))
        return (main + corr / (s["e8"] * f8 * 4096.0)).float()                      # two accumulators, joined in the epilogue

    def run(self, x_nhwc):
        P = self.P
        x = torch.as_tensor(x_nhwc).float().permute(0, 3, 1, 2).contiguous()
        s = self.store(x, float(x.abs().max()))             # the image: the standardisation pass knows its amax
        tap = None
        for name, k, cin, cout, then in self.table:
            w = P[name + "/weights"]
            if then == "after_concat":                       # both halves of the concat buffer share one pair of scales
                r = D.reorg_oracle(tap["x"].permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
                s = self.store(torch.cat([r, s["x"]], dim=1), max(tap["bound"], s["bound"]))
            z = self.conv(s, w)
            if then == "linear":
                return (z + P[name + "/biases"].view(1, -1, 1, 1)).permute(0, 2, 3, 1).contiguous().numpy()
            inv = torch.rsqrt(P[name + "/BatchNorm/moving_variance"] + D.BN_EPS) * P[name + "/BatchNorm/gamma"]
            b = P[name + "/BatchNorm/beta"] - P[name + "/BatchNorm/moving_mean"] * inv
            y = D.leaky_oracle(z * inv.view(1, -1, 1, 1) + b.view(1, -1, 1, 1))
            bound = float((inv.abs() * w.abs().sum(dim=(0, 1, 2))).max()) * s["amax"] + float(b.abs().max())
            if then == "passthrough+pool":
                tap = self.store(y, bound)
            if then in ("pool", "passthrough+pool"):
                y = F.max_pool2d(y, 2, 2)
            s = self.store(y, bound)
            self.looseness.append(bound / s["amax"])


def _case(mode, scale, size=32):
    params = D.init_params(20, 5, seed=1, mode=mode)
    x = (np.random.RandomState(0).normal(0, 1, size=(2, size, size, 3)) * scale).astype(np.float32)
    ref = D.darknet_oracle(x, params, 20, 5, dtype=torch.float64)
    return params, x, (lambda y: float(np.abs(y - ref).max() / np.abs(ref).max()))


@pytest.mark.parametrize("mode,scale", [("conditioned", 1.0), ("conditioned", 1e3), ("xavier", 1.0), ("xavier", 1e-3)])
def test_candidate_mode_stays_inside_the_parity_bar_with_dynamic_scales(mode, scale):
    params, x, rel = _case(mode, scale)
    em = Emulator(params)
    err = rel(em.run(x))
    assert err < 6e-5, err                                     # bar 1e-4; the shipped bf16x3 mode emulates to 1.6e-5 - 2e-5
    assert max(em.looseness) < 2.0 ** 11                       # the one-layer bound wastes at most ~10 binades (2^8.7 at 64x64 input,
                                                               # 2^10.2 on this 32x32 input whose 1x1 maps are mostly SAME padding)


def test_fp16_alone_and_fixed_scales_do_not():
    params, x, rel = _case("conditioned", 1.0)
    assert rel(Emulator(params, corrections=False).run(x)) > 3e-4           # 11 significand bits are not enough
    assert rel(Emulator(params, dynamic=False).run(x)) < 6e-5               # O(1) activations: fixed scales would do ...
    params, x, rel = _case("xavier", 1 / 30.0)
    assert rel(Emulator(params, dynamic=False).run(x)) > 3e-4               # ... the decaying random-init network is why they cannot be fixed
    assert rel(Emulator(params).run(x)) < 6e-5
