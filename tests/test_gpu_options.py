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


@pytest.mark.parametrize("arch,classes,size,batch", [("darknet", 80, 416, 4), ("darknet", 20, 96, 32), ("tiny", 20, 128, 8)])
def test_alternating_workspace_arenas_give_the_same_output_in_less_memory(cuda, arch, classes, size, batch):
    """Product default: a layer's output lives until the next layer has read it, so the outputs alternate between two arenas
    (the test suite otherwise runs with one slot per layer, keep_activations = 1, to read every tap back).  Same bits out,
    >= 2x less workspace, and the per-layer getter says why it cannot serve."""
    import torch
    from oracle.darknet_oracle import tiny_layer_table
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import inference
    table = tiny_layer_table(classes, 5) if arch == "tiny" else None
    params = init_params(classes, 5, seed=21, table=table)
    store = variables.reset_default_store()
    store.assign({"yolo2_%s/%s" % (arch, k): v for k, v in params.items()})
    xd = torch.from_numpy(np.random.RandomState(6).normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)).to(cuda)
    fn = getattr(inference, arch)
    L = _lib.lib()
    keep = inference._Engine.KEEP_ACTIVATIONS
    try:
        outs, ws = {}, {}
        for k in (True, False):
            inference._Engine.KEEP_ACTIVATIONS = k
            _, out = fn(xd, classes, 5)
            torch.cuda.synchronize()
            _lib.check(L.y2_check_async_errors())
            eng = inference._Engine.get(torch.device("cuda:0"), classes, 5, inference.ARCH_TINY if arch == "tiny" else inference.ARCH_DARKNET)
            outs[k], ws[k] = out.clone(), L.y2_workspace_bytes(eng.h, batch, size, size)
            if not k:
                with pytest.raises(_lib.Y2Error, match="keep_activations"):
                    eng.activation(2, False, (batch, size // 4, size // 4, eng.layers[2][2]))
        assert torch.equal(outs[True], outs[False])
        assert ws[False] < ws[True], ws
        if arch == "darknet":           # the saving at the bench sizes (size query only): 1.8 -> 0.46 GB at B = 32 / 416^2, 30 -> 7 GB at B = 256 / 608^2
            big = {}
            for k in (True, False):
                inference._Engine.KEEP_ACTIVATIONS = k
                e = inference._Engine.get(torch.device("cuda:0"), classes, 5)
                big[k] = (L.y2_workspace_bytes(e.h, 32, 416, 416), L.y2_workspace_bytes(e.h, 256, 608, 608))
            assert big[False][0] * 3 <= big[True][0] and big[False][1] * 3 <= big[True][1], big
            assert big[False][1] <= 8 * 2 ** 30, big
    finally:
        inference._Engine.KEEP_ACTIVATIONS = keep
