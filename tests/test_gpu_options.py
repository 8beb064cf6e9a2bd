"""-m gpu regression guards for the conv kernel's execution options: every option changes HOW the result is
produced (halo tile vs im2col fetches, TMA-store vs per-thread stores, programmatic dependent launch, fused vs separate
max-pool), never WHAT -- outputs must be bit-identical between settings and within 1e-4 of the CPU oracle."""
import ctypes

import numpy as np
import pytest

from oracle.darknet_oracle import darknet_oracle, init_params

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _debug_set(key, value):
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    L.y2_debug_set.restype = ctypes.c_int
    assert L.y2_debug_set(key, float(value)) == 0


def _store(params):
    from yolo_tf_b200 import variables
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})


@pytest.mark.parametrize("classes,size,batch", [(20, 64, 3), (80, 416, 2)])
def test_forward_identical_across_kernel_options(cuda, classes, size, batch):
    import torch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import inference
    params = init_params(classes, 5, seed=11)
    _store(params)
    x = np.random.RandomState(4).normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    ref = darknet_oracle(x, params, classes, 5)
    xd = torch.from_numpy(x).to(cuda)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    L = _lib.lib()
    outs = {}
    try:
        for name, pdl, tma, halo in [("default", 1, 1, 1), ("no_pdl", 0, 1, 1), ("no_tma_store", 1, 0, 1), ("no_halo", 1, 1, 0),
                                     ("plain", 0, 0, 0)]:
            _debug_set(5, pdl)
            _debug_set(6, tma)
            _lib.check(L.y2_set_option(eng.h, b"halo", halo))          # also drops the cached plan
            _, out = inference.darknet(xd, classes, 5)
            torch.cuda.synchronize()
            _lib.check(L.y2_check_async_errors())
            outs[name] = out.cpu().numpy()
    finally:
        _debug_set(5, 1)
        _debug_set(6, 1)
        _lib.check(L.y2_set_option(eng.h, b"halo", 1))
    for name, o in outs.items():
        assert not np.isnan(o).any(), name
        assert np.array_equal(o, outs["default"]), name + " differs from the default configuration"
    err = float(np.abs(outs["default"].astype(np.float64) - ref).max() / np.abs(ref).max())
    assert err <= TOL, err


@pytest.mark.parametrize("b,h,w,cout", [(1, 16, 8, 64), (3, 48, 40, 32), (2, 208, 208, 64)])
def test_halo_conv_matches_im2col_and_fp64(cuda, b, h, w, cout):
    """3x3, 32 input channels through y2_conv2d: halo tile (shifted descriptor windows) vs im2col fetches."""
    import torch
    import torch.nn.functional as F
    from yolo_tf_b200 import _lib
    rs = np.random.RandomState(h + w + cout)
    x = torch.as_tensor(rs.normal(0, 1, size=(b, h, w, 32)).astype(np.float32)).to(cuda)
    wt = torch.as_tensor((rs.normal(0, 1, size=(3, 3, 32, cout)) / 17.0).astype(np.float32)).to(cuda)
    scale = torch.as_tensor(rs.uniform(0.5, 1.5, size=cout).astype(np.float32)).to(cuda)
    bias = torch.as_tensor(rs.normal(0, 0.1, size=cout).astype(np.float32)).to(cuda)
    got = {}
    try:
        for halo in (0, 1):
            _debug_set(4, halo)
            y = torch.full((b, h, w, cout), float("nan"), device=cuda)
            _lib.check(_lib.lib().y2_conv2d(_lib.ptr(x), b, h, w, 32, _lib.ptr(wt), 3, cout, _lib.ptr(scale), _lib.ptr(bias), 1,
                                            _lib.ptr(y), 0, 0, 0, None))
            torch.cuda.synchronize()
            got[halo] = y.cpu().numpy()
    finally:
        _debug_set(4, 0)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), wt.double().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    ref = ref * scale.double() + bias.double()
    ref = torch.maximum(ref, 0.1 * ref).cpu().numpy()
    assert np.array_equal(got[0], got[1])
    assert float(np.abs(got[1] - ref).max() / np.abs(ref).max()) <= TOL
