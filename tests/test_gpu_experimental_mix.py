"""GPU, opt-in (Y2_EXPERIMENTAL=1): the EXPERIMENTAL mixed-kind conv (csrc/y2_conv_mix.cu, y2_conv2d_mix) against float64.

This kernel is not on any product path and was written after the round's GPU budget was spent, so it has never run on a
B200; the test is skipped unless Y2_EXPERIMENTAL=1 so that the suite the driver runs only contains verified code.  What it
checks once enabled: one fp16 product + two e4m3 correction products accumulated into ONE TMEM tile reproduce the conv to
the error the CPU emulation predicts (tests/test_numerics_candidate_fp16_fp8.py); the fp16 product alone does not.
The first tests keep chains <= 36 k-blocks; the last one runs conv20's K (432 k-blocks) with and without the chain cap."""
import os

import numpy as np
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
    assert err <= 2e-5, err                      # emulation: 3e-6 .. 6e-6 per layer; shipped bf16x3 kernel: 4.5e-6


def test_fp16_product_alone_is_not_enough_and_tile_widths_agree():
    assert 5e-5 < _run(2, 26, 256, 512, 3, terms=1) < 3e-3
    a = _run(2, 26, 256, 512, 3, terms=7, block_n=128)
    assert a <= 2e-5, a


def test_partial_last_tile_and_linear_output():
    assert _run(1, 13, 64, 32, 3, terms=7, leaky=0) <= 2e-5        # 169 pixels: second M tile mostly out of range


def test_long_chains_need_the_cap_here_too():
    """conv20's K = 3 x 3 x 3072 (432 k-blocks of 64, 8 MMAs each): one uncapped chain carries the truncation bias of DESIGN 4.9;
    chains of <= 32 k-blocks summed in fp32 by the epilogue bring the error back to the short-chain level."""
    capped = _run(2, 13, 3072, 256, 3, terms=7, kcap=32)
    uncapped = _run(2, 13, 3072, 256, 3, terms=7, kcap=0)
    assert capped <= 2e-5, capped
    assert uncapped > capped, (uncapped, capped)
