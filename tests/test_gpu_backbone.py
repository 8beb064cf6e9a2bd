"""-m gpu parity tests for the conv stack: the tcgen05 implicit-GEMM kernel per layer shape
(vs torch fp64 conv on the same device and the CPU oracle), and the whole Darknet-19 forward
through the reference-shaped API vs the CPU oracle, layer by layer.

Tolerance (north_star): 1e-4 relative, measured as max|a-b| / max|b| per output tensor.
"""
import os

import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import conv_bn_leaky_oracle, darknet_oracle, init_params, layer_table

pytestmark = pytest.mark.gpu
TOL = 1e-4
CONV0_TC_DEFAULT = 2          # library default of y2_set_option(h, "conv0_tc", ...)


def _conv(cuda, x, w, scale, bias, leaky, precision=0, block_n=0, max_ctas=0):
    import torch
    from yolo_tf_b200 import _lib
    xs, ws = torch.as_tensor(x).to(cuda), torch.as_tensor(w).to(cuda)
    sc = torch.as_tensor(scale).to(cuda) if scale is not None else None
    bi = torch.as_tensor(bias).to(cuda) if bias is not None else None
    b, h, wd, cin = x.shape
    k, _, _, cout = w.shape
    y = torch.full((b, h, wd, cout), float("nan"), device=cuda)
    _lib.check(_lib.lib().y2_conv2d(_lib.ptr(xs), b, h, wd, cin, _lib.ptr(ws), k, cout, _lib.ptr(sc), _lib.ptr(bi),
                                    int(leaky), _lib.ptr(y), precision, block_n, max_ctas, None))
    torch.cuda.synchronize()
    return y.cpu().numpy()


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


# every distinct (k, cin, cout, spatial) of the 416 network, batch 2 (3 for the 13x13 layers: M tail)
SHAPES = [(2, 208, 32, 3, 64), (2, 104, 64, 3, 128), (2, 104, 128, 1, 64), (2, 52, 128, 3, 256), (2, 52, 256, 1, 128),
          (2, 26, 256, 3, 512), (2, 26, 512, 1, 256), (3, 13, 512, 3, 1024), (3, 13, 1024, 1, 512),
          (3, 13, 1024, 3, 1024), (3, 13, 3072, 3, 1024), (3, 13, 1024, 1, 425), (3, 13, 1024, 1, 125),
          (2, 19, 1024, 3, 1024), (1, 38, 256, 3, 512)]


@pytest.mark.parametrize("b,hw,cin,k,cout", SHAPES)
def test_conv_layer_shapes_vs_fp64(cuda, b, hw, cin, k, cout):
    import torch
    import torch.nn.functional as F
    rs = np.random.RandomState(cin + cout + hw)
    x = rs.normal(0, 1, size=(b, hw, hw, cin)).astype(np.float32)
    w = (rs.normal(0, 1, size=(k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
    scale = rs.uniform(0.5, 1.5, size=cout).astype(np.float32)
    bias = rs.normal(0, 0.1, size=cout).astype(np.float32)
    got = _conv(cuda, x, w, scale, bias, True)
    xd = torch.as_tensor(x).to(cuda).double().permute(0, 3, 1, 2)
    wd = torch.as_tensor(w).to(cuda).double().permute(3, 2, 0, 1)
    ref = F.conv2d(xd, wd, padding=k // 2).permute(0, 2, 3, 1) * torch.as_tensor(scale).to(cuda).double() \
        + torch.as_tensor(bias).to(cuda).double()
    ref = torch.maximum(ref, 0.1 * ref).cpu().numpy()
    assert not np.isnan(got).any()
    assert _rel(got, ref) <= TOL


@pytest.mark.parametrize("max_ctas,block_n", [(0, 0), (1, 0), (2, 0), (-3, 0), (-5, 128), (-37, 64), (-148, 32), (37, 64)])
def test_conv_streamk_and_tile_variants_vs_oracle(cuda, max_ctas, block_n):
    """Hybrid schedule and stream-K hand-off: 3 m-tiles x K=72 k-blocks; max_ctas > 0 caps the grid (data-parallel
    waves + cost-model remainder), max_ctas < 0 forces stream-K over |max_ctas| CTAs (up to ~24 partials per tile,
    more than the 6 the cp.async staging holds, so the direct-load overflow path runs too)."""
    rs = np.random.RandomState(4)
    x = rs.normal(0, 1, size=(2, 13, 13, 512)).astype(np.float32)
    w = (rs.normal(0, 1, size=(3, 3, 512, 256)) * 0.02).astype(np.float32)
    scale = rs.uniform(0.5, 1.5, size=256).astype(np.float32)
    bias = rs.normal(0, 0.1, size=256).astype(np.float32)
    ref = conv_bn_leaky_oracle(x, w, scale, bias)
    got = _conv(cuda, x, w, scale, bias, True, max_ctas=max_ctas, block_n=block_n)
    assert _rel(got, ref.astype(np.float64)) <= TOL
    again = _conv(cuda, x, w, scale, bias, True, max_ctas=max_ctas, block_n=block_n)
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32))      # fixed summation order: bit-reproducible


def test_conv_streamk_many_contributors(cuda):
    """2 tiles x 144 k-blocks over 72 CTAs: every tile is assembled from ~36 partials."""
    rs = np.random.RandomState(8)
    x = rs.normal(0, 1, size=(1, 13, 13, 1024)).astype(np.float32)
    w = (rs.normal(0, 1, size=(3, 3, 1024, 256)) * 0.015).astype(np.float32)
    ref = conv_bn_leaky_oracle(x, w, np.ones(256, np.float32), np.zeros(256, np.float32))
    got = _conv(cuda, x, w, None, None, True)
    assert _rel(got, ref.astype(np.float64)) <= TOL


def test_conv_single_pass_bf16_is_reduced_precision(cuda):
    """precision=1 (hi planes only) must be ~bf16-grade: proves the x3 split is what buys fp32 parity."""
    rs = np.random.RandomState(5)
    x = rs.normal(0, 1, size=(1, 26, 26, 256)).astype(np.float32)
    w = (rs.normal(0, 1, size=(3, 3, 256, 128)) * 0.03).astype(np.float32)
    ref = conv_bn_leaky_oracle(x, w, np.ones(128, np.float32), np.zeros(128, np.float32), leaky=False)
    e1 = _rel(_conv(cuda, x, w, None, None, False, precision=1), ref.astype(np.float64))
    e0 = _rel(_conv(cuda, x, w, None, None, False, precision=0), ref.astype(np.float64))
    assert e0 <= TOL and 1e-4 < e1 < 2e-2


def _setup_store(params):
    from yolo_tf_b200 import variables
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    return store


@pytest.mark.parametrize("classes,size,batch,anchors", [(20, 416, 2, ho.ANCHORS_VOC), (80, 416, 1, ho.ANCHORS_COCO),
                                                        (80, 608, 1, ho.ANCHORS_COCO), (20, 64, 3, ho.ANCHORS_VOC)])
def test_darknet_forward_layer_by_layer_vs_oracle(cuda, classes, size, batch, anchors):
    import torch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import inference
    params = init_params(classes, 5, seed=1)
    _setup_store(params)
    rs = np.random.RandomState(2)
    x = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    taps = {}
    ref = darknet_oracle(x, params, classes, 5, taps=taps)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    xd = torch.from_numpy(x).to(cuda)
    worst = {}
    # Pass 1 keeps every un-pooled tensor (separate pool kernels); pass 2 is the default plan, where the pool of a
    # layer may live in the conv epilogue and only the pooled tensor exists.
    for fuse in (0, 1):
        _lib.check(_lib.lib().y2_set_option(eng.h, b"fuse_pool", fuse))
        scope, out = inference.darknet(xd, classes, 5)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().y2_check_async_errors())
        assert scope == "yolo2_darknet"
        tag = "" if fuse == 0 else "[fused]"
        for i, (name, k, cin, cout, then) in enumerate(layer_table(classes, 5)[:-1]):
            if i == 0:
                got = eng.activation(0, True, taps["conv0/pool"].shape).cpu().numpy()
                worst["conv0/pool" + tag] = _rel(got, taps["conv0/pool"].astype(np.float64))
                continue
            has_pool = then in ("pool", "passthrough+pool")
            if fuse == 0:
                got = eng.activation(i, False, taps[name].shape).cpu().numpy()
                worst[name] = _rel(got, taps[name].astype(np.float64))
            if has_pool:
                gp = eng.activation(i, True, taps[name + "/pool"].shape).cpu().numpy()
                worst[name + "/pool" + tag] = _rel(gp, taps[name + "/pool"].astype(np.float64))
        worst["output" + tag] = _rel(out.cpu().numpy(), ref.astype(np.float64))
    print("per-layer rel err:", {k: "%.1e" % v for k, v in worst.items()})
    assert max(worst.values()) <= TOL, worst


def test_darknet_forward_vs_the_reference_graph_golden(cuda):
    """tests/golden/backbone_reference.npz = the reference's own darknet() graph builder run with a torch float64 stand-in for
    slim / tf (tests/golden/make_backbone_golden.py): the device forward against THOSE numbers directly (the CPU suite pins the
    oracle to them at 1e-9), incl. the taps that feed the passthrough / concat."""
    import torch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import inference
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "backbone_reference.npz"))
    _setup_store(init_params(20, 5, seed=1))
    x = torch.from_numpy(d["x64"]).to(cuda)
    eng = inference._Engine.get(torch.device("cuda:0"), 20, 5)
    _lib.check(_lib.lib().y2_set_option(eng.h, b"fuse_pool", 0))          # keep the un-pooled tensors readable
    try:
        scope, out = inference.darknet(x, 20, 5)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().y2_check_async_errors())
        assert scope == str(d["darknet_scope"]) == "yolo2_darknet"
        assert _rel(out.cpu().numpy(), d["darknet_out"]) <= TOL
        for i, hw, c in ((12, 4, 512), (19, 2, 1024), (20, 2, 1024)):
            got = eng.activation(i, False, (3, hw, hw, c)).cpu().numpy()[:, :4, :4, :16]
            # the fixture keeps a 4 x 4 x 16 corner of each tap, so the error is normalised by the corner's own (smaller) maximum
            assert _rel(got, d["darknet_tap_conv%d" % i]) <= 3 * TOL, i
    finally:
        _lib.check(_lib.lib().y2_set_option(eng.h, b"fuse_pool", 1))


def test_center_false_variants_read_biases_and_resync(cuda):
    """`_darknet` (inference.py:125-126): BN without beta + a separate `<conv>/biases` variable.  The engine must read the shift
    from `biases` (not from a zero-initialised beta), re-sync when the same store is used through the other variant, and name the
    training gradients after the variables that graph has (round-1 advisor finding)."""
    import torch
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder, inference
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "backbone_reference.npz"))
    params = init_params(20, 5, seed=1)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k.replace("BatchNorm/beta", "biases"): v for k, v in params.items()})
    x = torch.from_numpy(d["x64"]).to(cuda)
    _, out = inference._darknet(x, 20, 5)
    assert _rel(out.cpu().numpy(), d["darknet_nocenter_out"]) <= TOL
    # the same store through `darknet` (center=True): beta does not exist -> created as zeros -> a different function
    _, other = inference.darknet(x, 20, 5)
    assert _rel(other.cpu().numpy(), d["darknet_nocenter_out"]) > 1e-3
    _, again = inference._darknet(x, 20, 5)                     # ... and back: the freshness key includes (scope, center)
    assert torch.equal(again, out)
    # training through the center=False graph: gradients are named after `biases`
    builder = Builder.from_values([str(i) for i in range(20)], 64, 64, ho.ANCHORS_VOC, inference_name="_darknet")
    builder(x, training=True)
    builder.create_objectives(ho.synthetic_labels(x.shape[0], 20, 2, 2, seed=1))
    flat, views = builder.backward(allreduce=False)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    assert "yolo2_darknet/conv3/biases" in views and not any(k.endswith("BatchNorm/beta") for k in views)
    from yolo_tf_b200.optimizer import AdamOptimizer, create_train_op
    create_train_op(builder, AdamOptimizer(1e-3)).apply_gradients(flat, views)       # round 1: KeyError here
    torch.cuda.synchronize()


def test_fused_maxpool_epilogue_matches_separate_pool(cuda):
    """Batch 32 admits the spatial tiling (16x8x1, 8x8x2, 4x4x8, 2x2x32 pixel blocks) that lets the 2x2 max-pool run
    inside the conv epilogue (warp shuffles).  Fused and separate paths must agree bit for bit, and with the oracle."""
    import torch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import inference
    classes, size, batch = 20, 64, 32
    params = init_params(classes, 5, seed=5)
    _setup_store(params)
    rs = np.random.RandomState(3)
    x = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    xd = torch.from_numpy(x).to(cuda)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    outs, pooled = {}, {}
    for fuse in (1, 0):
        _lib.check(_lib.lib().y2_set_option(eng.h, b"fuse_pool", fuse))
        _, out = inference.darknet(xd, classes, 5)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().y2_check_async_errors())
        outs[fuse] = out.cpu().numpy()
        pooled[fuse] = {i: eng.activation(i, True, (batch, size >> (k + 1), size >> (k + 1), c)).cpu().numpy()
                        for k, (i, c) in enumerate([(0, 32), (1, 64), (4, 128), (7, 256), (12, 512)])}
        if fuse:
            with pytest.raises(_lib.Y2Error):
                eng.activation(1, False, (batch, 32, 32, 64))       # un-pooled conv1 is never materialised when fused
    _lib.check(_lib.lib().y2_set_option(eng.h, b"fuse_pool", 1))
    assert np.array_equal(outs[1].view(np.uint32), outs[0].view(np.uint32))
    for i in pooled[1]:
        assert np.array_equal(pooled[1][i].view(np.uint32), pooled[0][i].view(np.uint32)), i
    ref = darknet_oracle(x, params, classes, 5)
    assert _rel(outs[1], ref.astype(np.float64)) <= TOL


def test_xavier_checkpoint_config1(cuda):
    """BASELINE config 1 weights: what slim creates (Xavier-uniform, BN gamma=1 beta=0 mean=0 var=1)."""
    import torch
    from yolo_tf_b200.model.yolo2 import inference
    params = init_params(20, 5, seed=1, mode="xavier")
    _setup_store(params)
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, size=(416, 416, 3)).astype(np.float32)
    x = ((img - img.mean()) / max(img.std(), 1.0 / np.sqrt(img.size)))[None]      # utils/preprocess.py:23-25
    ref = darknet_oracle(x, params, 20, 5)
    _, out = inference.darknet(torch.from_numpy(x.astype(np.float32)).to(cuda), 20, 5)
    assert _rel(out.cpu().numpy(), ref.astype(np.float64)) <= TOL


def test_builder_surface_and_detection_pipeline(cuda):
    """detect.py-shaped use: Builder -> model.conf/xy_min/xy_max -> non_max_suppress (numpy, in place)."""
    import torch
    from oracle.nms_oracle import nms_oracle
    from yolo_tf_b200.model.yolo2 import Builder
    from yolo_tf_b200.utils import postprocess
    classes = 20
    params = init_params(classes, 5, seed=3)
    _setup_store(params)
    rs = np.random.RandomState(1)
    x = rs.normal(0, 1, size=(1, 96, 96, 3)).astype(np.float32)
    builder = Builder.from_values([str(i) for i in range(classes)], 96, 96, ho.ANCHORS_VOC)
    builder(torch.from_numpy(x).to(cuda))
    m = builder.model
    assert (m.cell_height, m.cell_width) == (3, 3) and m.conf.shape == (1, 9, 5, classes)
    ref = ho.decode_oracle(darknet_oracle(x, params, classes, 5), classes, ho.ANCHORS_VOC)
    for k in ("conf", "xy_min", "xy_max", "iou", "prob", "wh", "coords", "offset_xy_min", "areas", "wh01_sqrt"):
        got = getattr(m, k).cpu().numpy()
        assert np.abs(got - ref[k]).max() <= 1e-4 * max(1.0, np.abs(ref[k]).max()), k
    conf = m.conf[0].cpu().numpy()
    lo, hi = m.xy_min[0].cpu().numpy(), m.xy_max[0].cpu().numpy()
    c_ref = conf.copy()
    order_ref = nms_oracle(c_ref, lo, hi, 0.02, 0.4)
    boxes = postprocess.non_max_suppress(conf, lo, hi, 0.02, 0.4)
    assert np.array_equal(conf.view(np.uint32), c_ref.view(np.uint32))          # in-place side effect, bit-exact
    assert len(boxes) == 45 and boxes[0][0].shape == (classes,)
    flat = conf.reshape(-1, classes)
    assert all(np.shares_memory(b[0], conf) for b in boxes[:3])
    got_order = [int((b[0].__array_interface__["data"][0] - flat.__array_interface__["data"][0]) // (4 * classes)) for b in boxes]
    assert got_order == order_ref.tolist()


def test_objectives_through_builder(cuda):
    import torch
    from yolo_tf_b200.model.yolo2 import Model, Objectives
    rs = np.random.RandomState(6)
    classes, hc, wc, B = 20, 13, 13, 8
    net = rs.normal(0, 1, size=(B, hc, wc, 5 * (5 + classes))).astype(np.float32)
    labels = ho.synthetic_labels(B, classes, wc, hc, seed=6)
    model = Model(torch.from_numpy(net).to(cuda), classes, ho.ANCHORS_VOC, training=True)
    obj = Objectives(model, *labels, hparam=ho.HPARAM_DEFAULT)
    ref_obj, ref_g = ho.loss_grad_oracle(net, classes, ho.ANCHORS_VOC, labels, dtype=np.float64)
    for k in ref_obj:
        assert abs(float(obj[k]) - ref_obj[k]) <= 1e-4 * abs(ref_obj[k]), k
    assert abs(float(obj.total_loss()) - ho.total_loss_oracle(ref_obj)) <= 1e-4 * ho.total_loss_oracle(ref_obj)
    assert np.abs(obj.grad_inputs.cpu().numpy() - ref_g).max() <= 1e-4 * np.abs(ref_g).max()
    with pytest.raises(AttributeError):
        model.conf                       # not built when training (model/yolo2/__init__.py:50)


@pytest.mark.parametrize("tc_mode", [1, 2])
@pytest.mark.parametrize("arch,classes,size,batch", [("darknet", 20, 64, 3), ("darknet", 80, 416, 2), ("darknet", 20, 608, 1),
                                                     ("darknet", 20, 96, 32), ("tiny", 20, 416, 2)])
def test_conv0_on_tensor_cores_vs_oracle_and_cuda_core_kernel(cuda, tc_mode, arch, classes, size, batch):
    """conv0 as SIMT-built im2col tiles + tcgen05 (option conv0_tc; 2 = unchecked gather for interior tiles): the pooled conv0
    activation against the oracle and against the exact-fp32 CUDA-core kernel (they differ only by the 2^-17 operand
    split), and the network output."""
    import torch
    from oracle.darknet_oracle import tiny_layer_table, tiny_oracle
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import inference
    table = tiny_layer_table(classes, 5) if arch == "tiny" else None
    params = init_params(classes, 5, seed=9, table=table)
    store = variables.reset_default_store()
    store.assign({"yolo2_%s/" % arch + k: v for k, v in params.items()})
    rs = np.random.RandomState(5)
    x = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    nref = min(batch, 2)
    taps = {}
    ref = (tiny_oracle if arch == "tiny" else darknet_oracle)(x[:nref], params, classes, 5, taps=taps)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5, inference.ARCH_TINY if arch == "tiny" else inference.ARCH_DARKNET)
    fn = inference.tiny if arch == "tiny" else inference.darknet
    xd = torch.from_numpy(x).to(cuda)
    c0 = taps["conv0/pool"].shape[-1]
    got = {}
    try:
        for mode in (0, tc_mode):
            _lib.check(_lib.lib().y2_set_option(eng.h, b"conv0_tc", mode))
            _, out = fn(xd, classes, 5)
            torch.cuda.synchronize()
            _lib.check(_lib.lib().y2_check_async_errors())
            got[mode] = (eng.activation(0, True, (batch, size // 2, size // 2, c0)).cpu().numpy(), out.cpu().numpy())
    finally:
        _lib.check(_lib.lib().y2_set_option(eng.h, b"conv0_tc", CONV0_TC_DEFAULT))
    assert _rel(got[tc_mode][0][:nref], taps["conv0/pool"].astype(np.float64)) <= 2e-5
    assert _rel(got[tc_mode][0], got[0][0].astype(np.float64)) <= 2e-5
    assert _rel(got[tc_mode][1][:nref], ref.astype(np.float64)) <= TOL
