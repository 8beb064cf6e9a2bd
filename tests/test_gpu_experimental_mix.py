"""GPU, opt-in (Y2_EXPERIMENTAL=1): the EXPERIMENTAL mixed-kind conv (csrc/y2_conv_mix.cu, y2_conv2d_mix) against float64.

This kernel is not on any product path and was written after the round's GPU budget was spent, so it has never run on a
B200; the test is skipped unless Y2_EXPERIMENTAL=1 so that the suite the driver runs only contains verified code.  What it
checks once enabled: one fp16 product + two e4m3 correction products accumulated into ONE TMEM tile reproduce the conv to
the error the CPU emulation predicts (tests/test_numerics_candidate_fp16_fp8.py); the fp16 product alone does not.
The first tests keep chains <= 36 k-blocks; the last one runs conv20's K (432 k-blocks) with and without the chain cap."""
import os

import pytest
import torch
import torch.nn.functional as F

from yolo_tf_b200 import _lib

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("Y2_EXPERIMENTAL") != "1", reason="experimental kernel, never run on a GPU yet: set Y2_EXPERIMENTAL=1")]


def _run(b, hw, cin, cout, k, terms, leaky=1, block_n=0, seed=0, kcap=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g)
    x = torch.maximum(x, 0.1 * x)
    w = torch.randn(k, k, cin, cout, device="cuda", generator=g) * (2.0 / (k * k * cin)) ** 0.5
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    bi = torch.randn(cout, device="cuda", generator=g) * 0.1
    y = torch.full((b, hw, hw, cout), float("nan"), device="cuda")
    _lib.check(_lib.lib().y2_conv2d_mix(_lib.ptr(x), b, hw, hw, cin, _lib.ptr(w), k, cout, _lib.ptr(sc), _lib.ptr(bi), leaky, _lib.ptr(y),
                                        terms, kcap, block_n, None))
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
    ref = ref * sc.double() + bi.double()
    if leaky:
        ref = torch.maximum(ref, 0.1 * ref)
    assert torch.isfinite(y).all()
    return float((y.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("shape", [(2, 26, 256, 512, 3), (2, 13, 128, 1024, 3), (3, 13, 1024, 425, 1), (1, 52, 128, 64, 1), (2, 26, 64, 96, 3)])
def test_mixed_kind_conv_matches_fp64(shape):
    err = _run(*shape, terms=7)
    # exact-accumulation emulation of these shapes: 0.9e-5 .. 1.4e-5 (shipped bf16x3 kernel: 4.5e-6 measured); the fp16 term alone
    # gives 2.8e-4 and a wrong descriptor / format garbage, so 3e-5 separates "works" from "does not" with room for the accumulator
    assert err <= 3e-5, err


def test_fp16_product_alone_is_not_enough_and_tile_widths_agree():
    assert 5e-5 < _run(2, 26, 256, 512, 3, terms=1) < 3e-3
    a = _run(2, 26, 256, 512, 3, terms=7, block_n=128)
    assert a <= 3e-5, a


def test_partial_last_tile_and_linear_output():
    assert _run(1, 13, 64, 32, 3, terms=7, leaky=0) <= 3e-5        # 169 pixels: second M tile mostly out of range


def test_long_chains_need_the_cap_here_too():
    """conv20's K = 3 x 3 x 3072 (432 k-blocks of 64, 8 MMAs each): one uncapped chain carries the truncation bias of DESIGN 4.9;
    chains of <= 32 k-blocks summed in fp32 by the epilogue bring the error back to the short-chain level."""
    capped = _run(2, 13, 3072, 256, 3, terms=7, kcap=32)
    uncapped = _run(2, 13, 3072, 256, 3, terms=7, kcap=0)
    assert capped <= 3e-5, capped                 # ~1e-5 from the operand formats + ~0.6e-5 of truncation bias per 32-k-block chain
    assert uncapped > capped, (uncapped, capped)


def test_two_layers_chained_in_the_storage_format():
    """Layer 1 writes its result in the format layer 2 reads (fp16 + e4m3 + e4m3 residual, power-of-two scales from the
    one-layer bound S * amax_in + T) and publishes its amax; layer 2 consumes that directly.  Checks the stored triple against
    the float32 output it encodes, the published amax, and the two-layer result against float64."""
    import math
    import torch.nn.functional as F
    L = _lib.lib()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    b, hw, c0, c1, c2 = 2, 26, 128, 256, 128
    x = torch.randn(b, hw, hw, c0, device=dev, generator=g)
    x = torch.maximum(x, 0.1 * x)
    w1 = torch.randn(3, 3, c0, c1, device=dev, generator=g) * (2.0 / (9 * c0)) ** 0.5
    w2 = torch.randn(1, 1, c1, c2, device=dev, generator=g) * (2.0 / c1) ** 0.5
    s1, b1 = torch.rand(c1, device=dev, generator=g) + 0.5, torch.randn(c1, device=dev, generator=g) * 0.1
    s2, b2 = torch.rand(c2, device=dev, generator=g) + 0.5, torch.randn(c2, device=dev, generator=g) * 0.1
    amax_x = float(x.abs().max())
    bound1 = float((s1.abs() * w1.abs().sum(dim=(0, 1, 2))).max()) * amax_x + float(b1.abs().max())       # S * amax_in + T
    n0, n1 = x.numel(), b * hw * hw * c1
    x16 = torch.empty(n0, dtype=torch.float16, device=dev)
    x8, rx8 = torch.empty(n0, dtype=torch.uint8, device=dev), torch.empty(n0, dtype=torch.uint8, device=dev)
    _lib.check(L.y2_mix_split(_lib.ptr(x), n0, amax_x, _lib.ptr(x16), _lib.ptr(x8), _lib.ptr(rx8), None))
    y1 = torch.full((b, hw, hw, c1), float("nan"), device=dev)
    o16 = torch.empty(n1, dtype=torch.float16, device=dev)
    o8, or8 = torch.empty(n1, dtype=torch.uint8, device=dev), torch.empty(n1, dtype=torch.uint8, device=dev)
    amax1 = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.y2_conv2d_mix_pre(_lib.ptr(x16), _lib.ptr(x8), _lib.ptr(rx8), amax_x, b, hw, hw, c0, _lib.ptr(w1), 3, c1, _lib.ptr(s1), _lib.ptr(b1), 1,
                                   _lib.ptr(y1), _lib.ptr(o16), _lib.ptr(o8), _lib.ptr(or8), bound1, _lib.ptr(amax1), 7, 32, 0, None))
    torch.cuda.synchronize()
    ref1 = F.conv2d(x.double().permute(0, 3, 1, 2), w1.double().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1) * s1.double() + b1.double()
    ref1 = torch.maximum(ref1, 0.1 * ref1)
    assert float((y1.double() - ref1).abs().max() / ref1.abs().max()) <= 3e-5        # emulation: 1.0e-5
    assert float(y1.abs().max()) <= bound1                                                # the bound is a bound
    assert amax1.view(torch.float32).item() == float(y1.abs().max())                     # published amax = what was written
    ba = 2.0 ** math.ceil(math.log2(bound1))
    e16, e8 = 32768.0 / ba, 256.0 / ba
    dec = o16.double() / e16 + or8.view(torch.float8_e4m3fn).double() / (e8 * 4096.0)
    assert float((dec.view_as(y1) - y1.double()).abs().max()) <= 2.0 ** -15 * ba          # fp16 + 4 residual bits of the scaled range
    assert float((o8.view(torch.float8_e4m3fn).double().view_as(y1) / e8 - y1.double()).abs().max()) <= 2.0 ** -4 * float(y1.abs().max()) + 2.0 ** -10 / e8
    y2 = torch.full((b, hw, hw, c2), float("nan"), device=dev)
    _lib.check(L.y2_conv2d_mix_pre(_lib.ptr(o16), _lib.ptr(o8), _lib.ptr(or8), bound1, b, hw, hw, c1, _lib.ptr(w2), 1, c2, _lib.ptr(s2), _lib.ptr(b2), 1,
                                   _lib.ptr(y2), None, None, None, 0.0, None, 7, 32, 0, None))
    torch.cuda.synchronize()
    _lib.check(L.y2_check_async_errors())
    ref2 = F.conv2d(ref1.permute(0, 3, 1, 2), w2.double().permute(3, 2, 0, 1)).permute(0, 2, 3, 1) * s2.double() + b2.double()
    ref2 = torch.maximum(ref2, 0.1 * ref2)
    assert float((y2.double() - ref2).abs().max() / ref2.abs().max()) <= 4e-5        # emulation: 1.1e-5
