"""-m gpu parity tests for the HBM-bound kernels (reorg, head decode, loss fwd+bwd, NMS):
CUDA through the C ABI vs the CPU oracle on identical seeded inputs, plus the committed goldens."""
import ctypes
import glob
import os

import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import reorg_oracle
from oracle.nms_c import nms_c_batch
from oracle.nms_oracle import nms_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _t(a, dev):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a)).to(dev)


# ------------------------------------------------------------------ reorg (bit-exact permutation)
def test_reorg_reference_selftest_vector(cuda):
    """The only golden vector in the reference: model/yolo2/function.py:32-50."""
    import torch
    from yolo_tf_b200 import _lib
    img = np.array([[0, 1, 0, 1], [2, 3, 2, 3], [0, 1, 0, 1], [2, 3, 2, 3]], np.float32).reshape(1, 4, 4, 1)
    x = _t(img, cuda)
    y = torch.empty(1, 2, 2, 4, device=cuda)
    _lib.check(_lib.lib().y2_reorg(_lib.ptr(x), 1, 4, 4, 1, 2, _lib.ptr(y), None))
    out = y.cpu().numpy()
    for i in range(4):
        assert np.all(out[0, :, :, i] == i)


@pytest.mark.parametrize("shape", [(2, 26, 26, 512), (1, 38, 38, 512), (3, 4, 6, 5), (2, 8, 8, 2)])
def test_reorg_random_bit_exact(cuda, shape):
    import torch
    from yolo_tf_b200 import _lib
    rs = np.random.RandomState(0)
    a = rs.normal(size=shape).astype(np.float32)
    x = _t(a, cuda)
    b, h, w, c = shape
    y = torch.empty(b, h // 2, w // 2, 4 * c, device=cuda)
    _lib.check(_lib.lib().y2_reorg(_lib.ptr(x), b, h, w, c, 2, _lib.ptr(y), None))
    assert np.array_equal(y.cpu().numpy().view(np.uint32), reorg_oracle(a).view(np.uint32))


# ------------------------------------------------------------------ leaky_relu (standalone op; fused everywhere inside the network)
@pytest.mark.parametrize("shape,alpha", [((2, 13, 13, 1024), 0.1), ((3, 7, 5, 3), 0.1), ((1001,), 0.25), ((4, 416, 416, 32), 0.1)])
def test_leaky_relu_standalone_bit_exact(cuda, shape, alpha):
    """model/yolo/function.py:21-24: max(x, alpha * x) in float32, incl. -0.0, +-inf, NaN, a size that is no multiple of 4 and an
    unaligned view (scalar path)."""
    import torch
    from yolo_tf_b200.model.yolo.function import leaky_relu
    rs = np.random.RandomState(1)
    a = rs.normal(0, 3, size=shape).astype(np.float32)
    flat = a.reshape(-1)
    flat[:6] = [0.0, -0.0, np.inf, -np.inf, np.nan, -1e-45]
    with np.errstate(invalid="ignore"):
        want = np.maximum(a, np.float32(alpha) * a)
    got = leaky_relu(_t(a, cuda), alpha).cpu().numpy()
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))
    if flat.size > 8:                                                  # 4-byte-aligned view: the scalar path
        x = _t(flat, cuda)[1:]
        got1 = leaky_relu(x, alpha).cpu().numpy()
        w1 = want.reshape(-1)[1:]
        ok = ~np.isnan(w1)
        assert np.array_equal(got1[ok].view(np.uint32), w1[ok].view(np.uint32))


# ------------------------------------------------------------------ decode
@pytest.mark.parametrize("B,hc,wc,C,anchors", [(2, 13, 13, 20, ho.ANCHORS_VOC), (3, 19, 19, 80, ho.ANCHORS_COCO),
                                               (1, 3, 5, 7, ho.ANCHORS_VOC[:3]), (5, 13, 13, 1, ho.ANCHORS_COCO)])
def test_head_decode_vs_oracle(cuda, B, hc, wc, C, anchors):
    import torch
    from yolo_tf_b200 import _lib
    rs = np.random.RandomState(7)
    A = len(anchors)
    net = rs.normal(0, 1.5, size=(B, hc, wc, A * (5 + C))).astype(np.float32)
    m = ho.decode_oracle(net, C, anchors)
    N = hc * wc * A
    dev = {k: torch.empty(*s, device=cuda) for k, s in {
        "conf": (B, N, C), "xy_min": (B, N, 2), "xy_max": (B, N, 2), "iou": (B, N), "prob": (B, N, C), "wh": (B, N, 2),
        "areas": (B, N), "xy": (B, N, 2), "offset_xy": (B, N, 2), "offset_xy_min": (B, N, 2),
        "offset_xy_max": (B, N, 2), "coords": (B, N, 4), "wh01": (B, N, 2)}.items()}
    outs = _lib.HeadOutputs(**{k: v.data_ptr() for k, v in dev.items()})
    x = _t(net, cuda)
    anc = _t(np.asarray(anchors, np.float32), cuda)
    _lib.check(_lib.lib().y2_head_decode(_lib.ptr(x), B, hc, wc, A, C, _lib.ptr(anc), ctypes.byref(outs), None))
    for k, v in dev.items():
        ref = np.asarray(m[k]).reshape(v.shape)
        got = v.cpu().numpy()
        tol = 1e-4 * max(1.0, float(np.abs(ref).max()))     # north_star: within 1e-4 relative
        assert np.abs(got - ref).max() <= tol, (k, float(np.abs(got - ref).max()))


# ------------------------------------------------------------------ loss forward + backward
@pytest.mark.parametrize("B,hc,wc,C,anchors,seed", [(4, 13, 13, 20, ho.ANCHORS_VOC, 3), (2, 19, 19, 80, ho.ANCHORS_COCO, 4),
                                                    (64, 13, 13, 20, ho.ANCHORS_VOC, 5)])
def test_loss_fwd_bwd_vs_oracle(cuda, B, hc, wc, C, anchors, seed):
    import torch
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(seed)
    A = len(anchors)
    net = rs.normal(0, 1.0, size=(B, hc, wc, A * (5 + C))).astype(np.float32)
    labels = ho.synthetic_labels(B, C, wc, hc, seed=seed)
    obj64, g64 = ho.loss_grad_oracle(net, C, anchors, labels, dtype=np.float64)
    x = _t(net, cuda)
    anc = _t(np.asarray(anchors, np.float32), cuda)
    lab = [_t(t.reshape(t.shape[0], t.shape[1], -1) if t.ndim > 2 else t, cuda) for t in labels]
    objs = torch.zeros(4, device=cuda)
    dnet = torch.full_like(x, float("nan"))
    ws_bytes = L.y2_loss_workspace_bytes(B, hc, wc)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    hp = (ctypes.c_float * 4)(1.0, 5.0, 1.0, 1.0)
    _lib.check(L.y2_loss_fwd_bwd(_lib.ptr(x), B, hc, wc, A, C, _lib.ptr(anc), *[_lib.ptr(t) for t in lab], hp,
                                 _lib.ptr(objs), _lib.ptr(dnet), _lib.ptr(ws), ws_bytes, None))
    got = objs.cpu().numpy()
    for i, k in enumerate(("prob", "iou_best", "iou_normal", "coords")):
        assert abs(got[i] - obj64[k]) <= 1e-4 * max(abs(obj64[k]), 1e-6), (k, got[i], obj64[k])
    g = dnet.cpu().numpy().astype(np.float64)
    assert not np.isnan(g).any()
    assert np.abs(g - g64).max() <= 1e-4 * np.abs(g64).max()


@pytest.mark.parametrize("name", ["voc", "coco", "nonsquare", "ties"])
def test_head_decode_and_loss_vs_the_reference_source_golden(cuda, name):
    """tests/golden/head_reference.npz = the reference's own Model / Objectives classes (model/yolo2/__init__.py:27-94) run with a
    torch stand-in for TF, float64, plus autograd through them (tests/golden/make_head_golden.py): the device decode, objectives
    and gradient against THOSE numbers directly (the CPU suite pins the oracle to them at 1e-12)."""
    import torch
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    d = np.load(os.path.join(GOLD, "head_reference.npz"))
    net, C, anchors = d[name + "_net"], int(d[name + "_meta"][0]), d[name + "_anchors"]
    labels = [d["%s_label_%s" % (name, n)] for n in ("mask", "prob", "coords", "offset_xy_min", "offset_xy_max", "areas")]
    B, hc, wc, _ = net.shape
    A, N = len(anchors), hc * wc * len(anchors)
    x = _t(net, cuda)
    anc = _t(np.asarray(anchors, np.float32), cuda)
    dev = {k: torch.empty(*s, device=cuda) for k, s in {"conf": (B, N, C), "xy_min": (B, N, 2), "xy_max": (B, N, 2), "iou": (B, N),
                                                        "prob": (B, N, C), "wh": (B, N, 2), "coords": (B, N, 4)}.items()}
    outs = _lib.HeadOutputs(**{k: v.data_ptr() for k, v in dev.items()})
    _lib.check(L.y2_head_decode(_lib.ptr(x), B, hc, wc, A, C, _lib.ptr(anc), ctypes.byref(outs), None))
    for k, v in dev.items():
        ref = d["%s_f64_model_%s" % (name, k)].reshape(v.shape)
        assert np.abs(v.cpu().numpy() - ref).max() <= 1e-4 * max(1.0, float(np.abs(ref).max())), k
    lab = [_t(t.reshape(t.shape[0], t.shape[1], -1) if t.ndim > 2 else t, cuda) for t in labels]
    objs = torch.zeros(4, device=cuda)
    dnet = torch.full_like(x, float("nan"))
    ws_bytes = L.y2_loss_workspace_bytes(B, hc, wc)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    hp = (ctypes.c_float * 4)(1.0, 5.0, 1.0, 1.0)
    _lib.check(L.y2_loss_fwd_bwd(_lib.ptr(x), B, hc, wc, A, C, _lib.ptr(anc), *[_lib.ptr(t) for t in lab], hp,
                                 _lib.ptr(objs), _lib.ptr(dnet), _lib.ptr(ws), ws_bytes, None))
    got = objs.cpu().numpy()
    for i, k in enumerate(("prob", "iou_best", "iou_normal", "coords")):
        want = float(d["%s_f64_obj_%s" % (name, k)])
        assert abs(got[i] - want) <= 1e-4 * max(abs(want), 1e-6), (k, got[i], want)
    g, g64 = dnet.cpu().numpy().astype(np.float64), d[name + "_f64_grad"]
    assert not np.isnan(g).any()
    assert np.abs(g - g64).max() <= 1e-4 * np.abs(g64).max()


# ------------------------------------------------------------------ NMS (bit-exact)
def _run_nms(cuda, conf, lo, hi, thr, thr_iou, want_order=True):
    import torch
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    B, N, C = conf.shape
    c, d_lo, d_hi = _t(conf, cuda), _t(lo, cuda), _t(hi, cuda)      # keep alive until the sync below
    order = torch.full((B, N), -1, dtype=torch.int32, device=cuda) if want_order else None
    status = torch.zeros(B, dtype=torch.int32, device=cuda)
    nbytes = L.y2_nms_workspace_bytes(B, N, C)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=cuda)
    _lib.check(L.y2_nms(_lib.ptr(c), _lib.ptr(d_lo), _lib.ptr(d_hi), B, N, C, thr, thr_iou,
                        _lib.ptr(order), _lib.ptr(status), _lib.ptr(ws), nbytes, None))
    torch.cuda.synchronize()
    return c.cpu().numpy(), (order.cpu().numpy() if want_order else None), status.cpu().numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "nms_*.npz"))))
def test_nms_matches_reference_goldens(cuda, path):
    d = np.load(path)
    conf = d["conf_in"]
    N, C = conf.shape[0] * conf.shape[1], conf.shape[2]
    got, order, status = _run_nms(cuda, conf.reshape(1, N, C), d["xy_min"].reshape(1, N, 2), d["xy_max"].reshape(1, N, 2),
                                  float(d["threshold"]), float(d["threshold_iou"]))
    assert status[0] == 0
    assert np.array_equal(got.reshape(conf.shape).view(np.uint32), d["conf_out"].view(np.uint32))
    assert np.array_equal(order[0], d["order"])


def _sweep_inputs(rs, B, hc, wc, C, K, quant=None):
    anchors = ho.ANCHORS_COCO
    A, cells = len(anchors), hc * wc
    gy, gx = np.meshgrid(np.arange(hc), np.arange(wc), indexing="ij")
    centre = np.stack([gx, gy], -1).reshape(1, cells, 1, 2) + rs.uniform(0, 1, size=(B, cells, A, 2))
    wh = anchors.reshape(1, 1, A, 2) * np.exp(rs.normal(0, 0.5, size=(B, cells, A, 2)))
    lo = (centre - wh / 2).astype(np.float32).reshape(B, cells * A, 2)
    hi = (centre + wh / 2).astype(np.float32).reshape(B, cells * A, 2)
    N = cells * A
    conf = rs.uniform(0, 0.29, size=(B, N * C))
    for b in range(B):
        pick = rs.choice(N * C, size=K, replace=False)
        conf[b, pick] = rs.uniform(0.3, 1.0, size=K)
    conf = conf.astype(np.float32)
    if quant:
        conf = (np.round(conf * quant) / quant).astype(np.float32)
    return conf.reshape(B, N, C), lo, hi


@pytest.mark.parametrize("B,hc,wc,C,K,quant", [(4, 13, 13, 80, 1000, None), (2, 19, 19, 80, 3000, None),
                                               (3, 13, 13, 20, 600, 32), (2, 13, 13, 80, 10000, 16), (2, 7, 7, 3, 400, 4),
                                               # 256 < candidates per class <= 1024: the whole-CTA select path with its own column scan, with ties
                                               (2, 13, 13, 4, 2500, 8),
                                               # > 1024 candidates per class: the general (global-memory) select path, with ties
                                               (1, 19, 19, 2, 3000, 8),
                                               # C % 4 != 0: scalar loads
                                               (2, 9, 9, 6, 700, None),
                                               # ~250 candidates per class with heavy ties: both select paths in one launch
                                               (1, 19, 19, 80, 20000, 64)])
def test_nms_vs_c_oracle(cuda, B, hc, wc, C, K, quant):
    rs = np.random.RandomState(11)
    conf, lo, hi = _sweep_inputs(rs, B, hc, wc, C, K, quant)
    ref = conf.copy()
    ref_order = nms_c_batch(ref, lo, hi, 0.3, 0.4)
    got, order, status = _run_nms(cuda, conf, lo, hi, 0.3, 0.4)
    assert not status.any()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(order, ref_order)


@pytest.mark.parametrize("thr_iou", [0.4, 1.0, 0.0, -0.5])
def test_nms_area_culling_at_the_boundary(cuda, thr_iou):
    """nms_apply_kernel drops IoU tests by an area window (CPU fuzz of the bound: tests/test_nms_cull_bound.py).  Inputs
    where the window is tight or degenerate: nested boxes whose area ratio is within ulps of threshold_iou, identical boxes,
    zero-area boxes, areas that overflow float32 / are denormal, and threshold_iou <= 0 (every pair "hits": no culling allowed)
    and = 1 (identical boxes only)."""
    rs = np.random.RandomState(31)
    B, N, C = 2, 1500, 8
    t = thr_iou if thr_iou > 0 else 0.4
    c = rs.uniform(0, 13, size=(B, N, 2))
    wh = np.exp(rs.uniform(-2, 2.5, size=(B, N, 2)))
    parent = rs.randint(0, 40, size=(B, N))                           # most boxes are nested copies of one of 40 "parents"
    nested = rs.rand(B, N) < 0.7
    ratio = t * (1.0 + rs.uniform(-2e-6, 2e-6, size=(B, N)))
    ratio = np.where(rs.rand(B, N) < 0.2, 1.0, ratio)                 # identical to the parent
    split = np.where(ratio == 1.0, 1.0, np.exp(rs.uniform(-0.2, 0.2, size=(B, N))))
    s = np.minimum(np.stack([np.sqrt(ratio) * split, np.sqrt(ratio) / split], -1), 1.0)
    bi = np.arange(B)[:, None]
    c = np.where(nested[..., None], c[bi, parent], c)
    wh = np.where(nested[..., None], wh[bi, parent] * s, wh)
    wh[:, 50:60, 1] = 0.0                                             # zero-area boxes
    wh[:, 60:64] *= 1e25                                              # area overflows to inf
    wh[:, 64:68] *= 1e-25                                             # denormal / zero area
    lo, hi = (c - wh / 2).astype(np.float32), (c + wh / 2).astype(np.float32)
    conf = rs.uniform(0, 0.29, size=(B, N, C)).astype(np.float32)
    conf[:, :40] = rs.uniform(0.3, 1.0, size=(B, 40, C))              # the parents are candidates of every class
    conf[:, 50:68:3] = 0.95
    ref = conf.copy()
    ref_order = nms_c_batch(ref, lo, hi, 0.3, thr_iou)
    got, order, status = _run_nms(cuda, conf, lo, hi, 0.3, thr_iou)
    assert not status.any()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(order, ref_order)
    assert (ref != conf).sum() > (1000 if thr_iou < 1 else 100)      # the case suppresses a lot


@pytest.mark.parametrize("C,quant", [(80, 16), (20, None), (6, 8)])
def test_nms_select_class_group_width(cuda, C, quant):
    """nms_select_kernel gives a CTA 8, 16 or 32 classes of an image (by regime; y2_debug_set key 13 forces it): all three must give
    the oracle's bits, incl. C % 4 != 0 (scalar loads) and a last group that is only partly filled."""
    import ctypes
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    rs = np.random.RandomState(17)
    conf, lo, hi = _sweep_inputs(rs, 3, 13, 13, C, 40 * C, quant)
    ref = conf.copy()
    ref_order = nms_c_batch(ref, lo, hi, 0.3, 0.4)
    try:
        for cg in (8, 16, 32):
            L.y2_debug_set(13, float(cg))
            got, order, status = _run_nms(cuda, conf, lo, hi, 0.3, 0.4)
            assert not status.any()
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), cg
            assert np.array_equal(order, ref_order), cg
    finally:
        L.y2_debug_set(13, 0.0)


def test_nms_many_images_takes_the_per_warp_path_for_heavy_classes(cuda):
    """With few CTAs in flight (every test above) a class with more than 64 candidates is handled by the whole CTA; a batch of
    hundreds of images (BASELINE configs[4]) keeps one warp per class.  B = 128, ~125 candidates per class with ties: both paths
    must give the bits of the C oracle."""
    rs = np.random.RandomState(23)
    conf, lo, hi = _sweep_inputs(rs, 128, 13, 13, 80, 10000, 64)
    got, _, status = _run_nms(cuda, conf, lo, hi, 0.3, 0.4, want_order=False)
    assert not status.any()
    for sl in (slice(0, 2), slice(126, 128)):
        ref = conf[sl].copy()
        nms_c_batch(ref, lo[sl], hi[sl], 0.3, 0.4)
        assert np.array_equal(got[sl].view(np.uint32), ref.view(np.uint32))
    small, _, _ = _run_nms(cuda, conf[:3], lo[:3], hi[:3], 0.3, 0.4, want_order=False)       # 9 CTAs: the cooperative path
    assert np.array_equal(small.view(np.uint32), got[:3].view(np.uint32))


def test_nms_python_oracle_small(cuda):
    rs = np.random.RandomState(5)
    conf, lo, hi = _sweep_inputs(rs, 1, 5, 5, 4, 60, 8)
    ref = conf[0].copy().reshape(25, 5, 4)
    ref_order = nms_oracle(ref, lo[0].reshape(25, 5, 2), hi[0].reshape(25, 5, 2), 0.3, 0.4)
    got, order, _ = _run_nms(cuda, conf, lo, hi, 0.3, 0.4)
    assert np.array_equal(got[0].reshape(ref.shape).view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(order[0], ref_order)


def test_nms_edge_cases(cuda):
    # empty batch / no boxes / single box / assert-equivalent status on an inverted box
    rs = np.random.RandomState(2)
    conf, lo, hi = _sweep_inputs(rs, 1, 2, 2, 3, 5)
    got, order, status = _run_nms(cuda, conf[:, :1], lo[:, :1], hi[:, :1], 0.3, 0.4)
    assert np.array_equal(got, conf[:, :1]) and order[0, 0] == 0 and status[0] == 0
    bad_lo = lo.copy()
    bad_lo[0, 3, 0] = hi[0, 3, 0] + 1.0                       # xy_min > xy_max: reference assert (postprocess.py:26-27)
    c2 = conf.copy()
    c2[0, 0, 0] = 0.9
    _, _, status = _run_nms(cuda, c2, bad_lo, hi, 0.3, 0.4)
    assert status[0] == 1
    with pytest.raises(AssertionError):
        nms_c_batch(c2.copy(), bad_lo, hi, 0.3, 0.4)


def test_nms_idempotent_and_full_size(cuda):
    """BASELINE config 5 sizes (B=512 would take the CPU oracle minutes): property checks at B=64, N=1805, C=80."""
    rs = np.random.RandomState(9)
    conf, lo, hi = _sweep_inputs(rs, 64, 19, 19, 80, 3000)
    once, _, _ = _run_nms(cuda, conf, lo, hi, 0.3, 0.4, want_order=False)
    twice, _, _ = _run_nms(cuda, once, lo, hi, 0.3, 0.4, want_order=False)
    assert np.array_equal(once.view(np.uint32), twice.view(np.uint32))        # NMS is idempotent
    changed = once != conf
    assert np.all(once[changed] == 0)                                          # only ever writes zeros
    ref = conf[:2].copy()
    nms_c_batch(ref, lo[:2], hi[:2], 0.3, 0.4)
    assert np.array_equal(once[:2].view(np.uint32), ref.view(np.uint32))
