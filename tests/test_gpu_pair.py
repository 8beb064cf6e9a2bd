"""-m gpu parity tests for the CTA-pair (tcgen05 cta_group::2) mode of the conv kernel: two SMs of one TPC share every MMA
(M = 256 rows per pair, each CTA stages its own A rows and half of the B tile).  Same oracle and tolerance as the
single-CTA kernel, and the two modes must agree to fp32 summation-order level on every shape."""
import ctypes

import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import conv_bn_leaky_oracle, darknet_oracle, init_params

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


def _conv(cuda, x, w, scale, bias, leaky, pair, precision=0, block_n=0, max_ctas=0):
    import torch
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    xs, ws = torch.as_tensor(x).to(cuda), torch.as_tensor(w).to(cuda)
    sc, bi = torch.as_tensor(scale).to(cuda), torch.as_tensor(bias).to(cuda)
    b, h, wd, cin = x.shape
    k, _, _, cout = w.shape
    y = torch.full((b, h, wd, cout), float("nan"), device=cuda)
    L.y2_debug_set(7, float(pair))
    try:
        _lib.check(L.y2_conv2d(_lib.ptr(xs), b, h, wd, cin, _lib.ptr(ws), k, cout, _lib.ptr(sc), _lib.ptr(bi), int(leaky), _lib.ptr(y),
                               precision, block_n, max_ctas, None))
        torch.cuda.synchronize()
    finally:
        L.y2_debug_set(7, 0.0)
    return y.cpu().numpy()


# (batch, hw, cin, k, cout): odd and even numbers of 128-row tiles, one and several N tiles, 1x1 and 3x3, the concat layer
SHAPES = [(3, 13, 512, 3, 1024), (3, 13, 1024, 1, 512), (2, 26, 256, 3, 512), (2, 52, 128, 3, 256), (3, 13, 3072, 3, 1024),
          (3, 13, 1024, 1, 425), (2, 104, 64, 3, 128), (1, 19, 1024, 3, 1024), (2, 104, 128, 1, 64)]


@pytest.mark.parametrize("b,hw,cin,k,cout", SHAPES)
def test_pair_conv_matches_oracle_and_single_cta(cuda, b, hw, cin, k, cout):
    rs = np.random.RandomState(cin + cout + hw)
    x = rs.normal(0, 1, size=(b, hw, hw, cin)).astype(np.float32)
    w = (rs.normal(0, 1, size=(k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
    scale = rs.uniform(0.5, 1.5, size=cout).astype(np.float32)
    bias = rs.normal(0, 0.1, size=cout).astype(np.float32)
    ref = conv_bn_leaky_oracle(x, w, scale, bias, dtype=__import__("torch").float64)
    got = _conv(cuda, x, w, scale, bias, True, pair=1)
    assert not np.isnan(got).any()
    assert _rel(got, ref) <= TOL
    single = _conv(cuda, x, w, scale, bias, True, pair=0)
    assert _rel(got, single.astype(np.float64)) <= 2e-5             # same products, different fp32 summation order
    again = _conv(cuda, x, w, scale, bias, True, pair=1)
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32))           # fixed summation order


@pytest.mark.parametrize("max_ctas,block_n", [(0, 0), (1, 0), (2, 0), (-3, 0), (-5, 128), (-37, 64), (-74, 32), (37, 64)])
def test_pair_streamk_and_tile_variants(cuda, max_ctas, block_n):
    """3 m-tiles (2 pair tiles, the second half empty) x 72 k-blocks: capped grids, forced stream-K with up to ~36 partials
    per tile (staged and overflow hand-off paths), narrow N tiles."""
    rs = np.random.RandomState(4)
    x = rs.normal(0, 1, size=(2, 13, 13, 512)).astype(np.float32)
    w = (rs.normal(0, 1, size=(3, 3, 512, 256)) * 0.02).astype(np.float32)
    scale = rs.uniform(0.5, 1.5, size=256).astype(np.float32)
    bias = rs.normal(0, 0.1, size=256).astype(np.float32)
    ref = conv_bn_leaky_oracle(x, w, scale, bias)
    got = _conv(cuda, x, w, scale, bias, True, pair=1, max_ctas=max_ctas, block_n=block_n)
    assert _rel(got, ref.astype(np.float64)) <= TOL
    p1 = _conv(cuda, x, w, scale, bias, True, pair=1, max_ctas=max_ctas, block_n=block_n, precision=1)
    assert _rel(p1, ref.astype(np.float64)) <= 2e-2


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("classes,size,batch", [(20, 416, 2), (80, 416, 32), (80, 608, 1), (20, 64, 32)])
def test_darknet_forward_pair_mode_vs_oracle(cuda, mode, classes, size, batch):
    """Whole network with the pair option: 1 = the 3x3 layers with 256-wide N tiles, 2 = every eligible layer (incl. the
    fused max-pool layers' spatial tiles and the 1x1 layers)."""
    import torch
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import inference
    params = init_params(classes, 5, seed=1)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    rs = np.random.RandomState(2)
    nref = min(batch, 2)                                    # the CPU oracle on the first images only (batch 32 is slow on CPU)
    x = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    ref = darknet_oracle(x[:nref], params, classes, 5)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    xd = torch.from_numpy(x).to(cuda)
    _lib.check(_lib.lib().y2_set_option(eng.h, b"pair", 0))
    try:
        _, base = inference.darknet(xd, classes, 5)
        base = base.cpu().numpy()
        _lib.check(_lib.lib().y2_set_option(eng.h, b"pair", mode))
        _, out = inference.darknet(xd, classes, 5)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().y2_check_async_errors())
        out = out.cpu().numpy()
    finally:
        _lib.check(_lib.lib().y2_set_option(eng.h, b"pair", 1))      # library default
    print("pair mode %d: rel err vs oracle %.2e (single-CTA plan %.2e), pair vs single %.2e" % (
        mode, _rel(out[:nref], ref.astype(np.float64)), _rel(base[:nref], ref.astype(np.float64)), _rel(out, base.astype(np.float64))))
    assert _rel(out[:nref], ref.astype(np.float64)) <= TOL
    assert _rel(base[:nref], ref.astype(np.float64)) <= TOL       # (batch 32 x 416 x 80 classes is the bench configuration)
    assert _rel(out, base.astype(np.float64)) <= 5e-5           # two fp32-grade evaluations, different summation orders


@pytest.mark.parametrize("classes,size,batch,pick", [(80, 416, 32, (0, 13, 31)), (80, 608, 16, (0, 7, 15))])
def test_full_size_batch_independence(cuda, classes, size, batch, pick):
    """Size-independent property at BASELINE's full sizes (no oracle needed): an image's output must not depend on the batch
    it travels in.  The plans differ completely (tile schedule, stream-K splits, fused pools, chain lengths), so this is
    exactly the check that exposes a batch-size-dependent error such as the truncating accumulator (1.9e-4 at B = 32 before
    the chain cap); both evaluations are fp32-grade, so they must agree to the 1e-4 bar."""
    import torch
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import inference
    params = init_params(classes, 5, seed=1)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(batch, size, size, 3, device="cuda", generator=g)
    _, full = inference.darknet(x, classes, 5)
    full = full.cpu().numpy()
    for i in pick:
        _, one = inference.darknet(x[i:i + 1].contiguous(), classes, 5)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().y2_check_async_errors())
        one = one.cpu().numpy()
        assert _rel(full[i:i + 1], one.astype(np.float64)) <= TOL, i
