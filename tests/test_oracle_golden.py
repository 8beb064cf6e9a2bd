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
