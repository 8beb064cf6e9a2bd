"""CPU (-m "not gpu"): pins the oracle against everything the reference's own tests hold for this
path -- the reorg self-test vector (model/yolo2/function.py:32-50) and the NMS outputs produced by
the reference's own utils/postprocess.py (tests/golden/nms_*.npz, made by make_nms_golden.py) --
and checks the oracle's internal consistency (closed-form loss gradient vs autograd, fp32 vs fp64)."""
import glob
import os

import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import (darknet_oracle, flops_per_image, init_params, layer_table, reorg_oracle)
from oracle.nms_c import nms_c_batch
from oracle.nms_oracle import nms_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NMS_FILES = sorted(glob.glob(os.path.join(GOLD, "nms_*.npz")))


def test_goldens_present():
    assert len(NMS_FILES) >= 9


def test_reorg_selftest_vector():
    img = np.array([[0, 1, 0, 1], [2, 3, 2, 3], [0, 1, 0, 1], [2, 3, 2, 3]], np.float32).reshape(1, 4, 4, 1)
    out = reorg_oracle(img)
    assert out.shape == (1, 2, 2, 4)
    for i in range(4):
        assert np.unique(out[0, :, :, i]).tolist() == [i]


@pytest.mark.parametrize("path", NMS_FILES)
def test_nms_python_oracle_matches_reference(path):
    d = np.load(path)
    conf = d["conf_in"].copy()
    order = nms_oracle(conf, d["xy_min"], d["xy_max"], float(d["threshold"]), float(d["threshold_iou"]))
    assert np.array_equal(conf.view(np.uint32), d["conf_out"].view(np.uint32))
    assert np.array_equal(order, d["order"])


@pytest.mark.parametrize("path", NMS_FILES)
def test_nms_c_oracle_matches_reference(path):
    d = np.load(path)
    shp = d["conf_in"].shape
    n, c = shp[0] * shp[1], shp[2]
    conf = d["conf_in"].copy().reshape(1, n, c)
    order = nms_c_batch(conf, d["xy_min"].reshape(1, n, 2), d["xy_max"].reshape(1, n, 2), float(d["threshold"]),
                        float(d["threshold_iou"]))
    assert np.array_equal(conf.reshape(shp).view(np.uint32), d["conf_out"].view(np.uint32))
    assert np.array_equal(order[0], d["order"])


def test_layer_table_matches_survey_flops():
    assert len(layer_table(20, 5)) == 22
    assert abs(flops_per_image(416, 416, 20, 5) / 1e9 - 34.898) < 1e-3
    assert abs(flops_per_image(416, 416, 80, 5) / 1e9 - 35.002) < 1e-3
    assert abs(flops_per_image(608, 608, 80, 5) / 1e9 - 74.768) < 1e-3


def test_loss_closed_form_gradient_equals_autograd():
    rs = np.random.RandomState(0)
    B, hc, wc, C = 2, 5, 5, 20
    net = rs.normal(0, 1, size=(B, hc, wc, 5 * (5 + C))).astype(np.float32)
    lab = ho.synthetic_labels(B, C, wc, hc, seed=3)
    obj, g = ho.loss_grad_oracle(net, C, ho.ANCHORS_VOC, lab, dtype=np.float64)
    obj2, g2 = ho.loss_grad_autograd(net, C, ho.ANCHORS_VOC, lab)
    for k in obj:
        assert abs(float(obj[k]) - obj2[k]) < 1e-12
    assert np.abs(g - g2).max() < 1e-15
    obj32, g32 = ho.loss_grad_oracle(net, C, ho.ANCHORS_VOC, lab)
    assert np.abs(g32 - g2).max() <= 1e-5 * np.abs(g2).max()


def test_labels_self_consistent():
    """utils/visualize.py:45-48 asserts offset_xy_min/max are consistent with coords (rtol 1e-3)."""
    mask, prob, coords, lo, hi, areas = ho.synthetic_labels(4, 20, 13, 13, seed=1)
    sel = mask[..., 0] > 0
    wh = (hi - lo)[sel][:, 0]
    np.testing.assert_allclose(wh[:, 0] / 13, coords[sel][:, 0, 2] ** 2, rtol=1e-3)
    np.testing.assert_allclose(wh[:, 1] / 13, coords[sel][:, 0, 3] ** 2, rtol=1e-3)
    np.testing.assert_allclose(areas[sel][:, 0], wh[:, 0] * wh[:, 1], rtol=1e-5)


def test_darknet_oracle_fp32_vs_fp64_small():
    import torch
    rs = np.random.RandomState(2)
    p = init_params(20, 5, seed=1)
    x = rs.normal(0, 1, size=(1, 64, 64, 3)).astype(np.float32)
    y32 = darknet_oracle(x, p, 20, 5)
    y64 = darknet_oracle(x, p, 20, 5, dtype=torch.float64)
    assert y32.shape == (1, 2, 2, 125)
    assert np.abs(y32 - y64).max() <= 1e-4 * np.abs(y64).max()


def test_tiny_oracle_table_pool_semantics_and_fp64():
    """tiny() (model/yolo2/inference.py:25-50): 9 convs, 6.971 GFLOP per 416x416 image (C=20); the stride-1 SAME pool clips
    its window at the bottom/right edge (TF pads after, never before, and max-pool ignores padding)."""
    import torch
    from oracle.darknet_oracle import max_pool_s1_same_oracle, tiny_layer_table, tiny_oracle
    t = tiny_layer_table(20, 5)
    assert [(c[2], c[3]) for c in t] == [(3, 16), (16, 32), (32, 64), (64, 128), (128, 256), (256, 512), (512, 1024), (1024, 1024), (1024, 125)]
    assert [c[4] for c in t] == ["pool"] * 5 + ["pool_s1", None, None, "linear"]
    assert abs(flops_per_image(416, 416, 20, 5, table=t) / 1e9 - 6.971) < 1e-3
    a = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(2, 3, 4, 5)
    a = (a * 7919 % 101) - 50
    b = max_pool_s1_same_oracle(a)
    assert b.shape == a.shape
    for y in range(4):
        for x in range(5):
            assert torch.equal(b[:, :, y, x], a[:, :, y:y + 2, x:x + 2].amax(dim=(2, 3)))
    rs = np.random.RandomState(4)
    p = init_params(20, 5, seed=2, table=t)
    x = rs.normal(0, 1, size=(1, 96, 96, 3)).astype(np.float32)
    y32 = tiny_oracle(x, p, 20, 5)
    y64 = tiny_oracle(x, p, 20, 5, dtype=torch.float64)
    assert y32.shape == (1, 3, 3, 125)
    assert np.abs(y32 - y64).max() <= 1e-4 * np.abs(y64).max()


def test_oracle_conv_bn_pool_follow_the_tf_definitions_written_out_as_loops():
    """TensorFlow itself cannot run here (parity unpinned), so the oracle's primitives are at least pinned to the
    DEFINITIONS the survey restates, spelled out as explicit loops that share no code with torch's conv:
    tf.nn.conv2d (NHWC, HWIO, stride 1, SAME) = cross-correlation out[b,y,x,o] = sum_{dy,dx,c} in[b,y+dy-p,x+dx-p,c] w[dy,dx,c,o]
    with zero padding p = k//2; tf.nn.batch_normalization: inv = rsqrt(var+eps)*gamma, y = x*inv + (beta - mean*inv);
    leaky_relu = max(x, 0.1x) (model/yolo/function.py:21-24); max_pool2d 2x2 stride 2 (inference.py:69)."""
    import torch
    import torch.nn.functional as F
    from oracle.darknet_oracle import BN_EPS, _conv_same, leaky_oracle
    rs = np.random.RandomState(9)
    for k, cin, cout, h, w in ((3, 3, 4, 5, 6), (1, 5, 3, 4, 4), (3, 2, 2, 1, 1)):
        x = rs.normal(size=(2, h, w, cin))
        wt = rs.normal(size=(k, k, cin, cout))
        want = np.zeros((2, h, w, cout))
        p = k // 2
        for b in range(2):
            for y in range(h):
                for xx in range(w):
                    for dy in range(k):
                        for dx in range(k):
                            yy, xs = y + dy - p, xx + dx - p
                            if 0 <= yy < h and 0 <= xs < w:
                                want[b, y, xx] += x[b, yy, xs] @ wt[dy, dx]
        got = _conv_same(torch.as_tensor(x).permute(0, 3, 1, 2), torch.as_tensor(wt)).permute(0, 2, 3, 1).numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    # one whole BN-conv layer + pool through darknet_oracle's code path, against the written-out arithmetic
    table = [("conv0", 3, 3, 4, "pool"), ("conv", 1, 4, 6, "linear")]
    prm = init_params(1, 1, seed=5, table=table)
    x = rs.normal(size=(1, 4, 4, 3))
    got = darknet_oracle(x, prm, 1, 1, dtype=torch.float64, table=table)
    z = _conv_same(torch.as_tensor(x).permute(0, 3, 1, 2), torch.as_tensor(prm["conv0/weights"]).double()).permute(0, 2, 3, 1).numpy()
    g, b_, m, v = (prm["conv0/BatchNorm/" + n].astype(np.float64) for n in ("gamma", "beta", "moving_mean", "moving_variance"))
    inv = g / np.sqrt(v + BN_EPS)
    a = z * inv + (b_ - m * inv)
    a = np.maximum(a, 0.1 * a)
    pooled = a.reshape(1, 2, 2, 2, 2, 4).max(axis=(2, 4))
    want = pooled @ prm["conv/weights"].astype(np.float64)[0, 0] + prm["conv/biases"].astype(np.float64)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    # training mode: batch mean and POPULATION variance over (B, H, W) (slim.batch_norm(is_training=True))
    got_t = darknet_oracle(np.concatenate([x, 2 * x]), prm, 1, 1, training=True, dtype=torch.float64, table=table)
    z2 = np.concatenate([z, 2 * z])
    mu, var = z2.mean(axis=(0, 1, 2)), z2.var(axis=(0, 1, 2))
    inv = g / np.sqrt(var + BN_EPS)
    a = z2 * inv + (b_ - mu * inv)
    a = np.maximum(a, 0.1 * a)
    want_t = a.reshape(2, 2, 2, 2, 2, 4).max(axis=(2, 4)) @ prm["conv/weights"].astype(np.float64)[0, 0] + prm["conv/biases"].astype(np.float64)
    np.testing.assert_allclose(got_t, want_t, rtol=1e-11, atol=1e-11)
    assert float(leaky_oracle(torch.tensor(-2.0, dtype=torch.float64))) == -0.2 and float(leaky_oracle(torch.tensor(0.0))) == 0.0


REF_NMS = "/root/reference/utils/postprocess.py"


@pytest.mark.skipif(not os.path.exists(REF_NMS), reason="the reference checkout only exists in the authoring container")
def test_nms_oracles_against_the_live_reference_on_random_cases():
    """Beyond the 9 committed goldens: 40 fresh random cases per run against the reference's own non_max_suppress loaded by path
    (pure numpy / Python, so it runs here) -- heavy score ties (quantised scores), duplicate boxes, thresholds hit exactly,
    1 .. 6 classes -- for the numpy oracle and its C twin, bit for bit in the zeroed matrix and in the returned order."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_postprocess_live", REF_NMS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rs = np.random.RandomState(20260117)
    for case in range(40):
        cells, a, c = rs.randint(1, 10), rs.randint(1, 4), rs.randint(1, 7)
        n = cells * a
        centre = rs.uniform(0, 4, size=(cells, a, 2))
        wh = rs.uniform(0.5, 3, size=(cells, a, 2))
        lo, hi = (centre - wh / 2).astype(np.float32), (centre + wh / 2).astype(np.float32)
        if case % 4 == 0 and n > 2:                                   # duplicate boxes
            lo[-1, -1], hi[-1, -1] = lo[0, 0], hi[0, 0]
        conf = rs.uniform(0, 1, size=(cells, a, c)).astype(np.float32)
        if case % 2 == 0:
            conf = (np.round(conf * 4) / 4).astype(np.float32)        # ties, and values exactly on a threshold of 0.25 / 0.5
        thr, thr_iou = [(0.3, 0.4), (0.25, 0.5), (0.5, 0.25)][case % 3]
        want = conf.copy()
        boxes = ref.non_max_suppress(want, lo, hi, thr, thr_iou)
        flat_lo = lo.reshape(n, 2)
        want_order = []
        for row, b_lo, _ in boxes:                                    # rows are views into `want`: recover each one's box index
            off = (row.__array_interface__["data"][0] - want.__array_interface__["data"][0]) // (4 * c)
            want_order.append(off)
            assert np.array_equal(b_lo, flat_lo[off])
        got = conf.copy()
        order = nms_oracle(got, lo, hi, thr, thr_iou)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), case
        assert list(order) == want_order, case
        got_c = conf.copy().reshape(1, n, c)
        order_c = nms_c_batch(got_c, lo.reshape(1, n, 2), hi.reshape(1, n, 2), thr, thr_iou)
        assert np.array_equal(got_c.reshape(cells, a, c).view(np.uint32), want.view(np.uint32)), case
        assert list(order_c[0]) == want_order, case
