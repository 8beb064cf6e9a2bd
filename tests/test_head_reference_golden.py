"""CPU (-m "not gpu"): pins the head oracle (oracle/head_oracle.py: decode, 4-part loss, closed-form gradient) to outputs of the
REFERENCE'S OWN `Model` / `Objectives` source (model/yolo2/__init__.py:27-94), executed by tests/golden/make_head_golden.py
with a torch stand-in for the TF ops it calls (TensorFlow 1.0 is not installable here) -> tests/golden/head_reference.npz.
Everything the reference's source decides is covered: channel order, slice -> op wiring, IoU / best-box ties / masks, cnt,
the objectives, and (autograd over that source) d(total_loss)/d(net), which the reference leaves to tf.gradients.
The float64 vectors are matched to 1e-12, the float32 oracle to the float32 run at the size of float32 summation noise."""
import os

import numpy as np
import pytest

from oracle import head_oracle as ho

GOLD = os.path.join(os.path.dirname(__file__), "golden", "head_reference.npz")
CASES = ("voc", "coco", "nonsquare", "ties")
ATTRS = ("iou", "offset_xy", "wh", "prob", "areas", "offset_xy_min", "offset_xy_max", "wh01", "wh01_sqrt", "coords", "xy", "xy_min",
         "xy_max", "conf")
HPARAM = {"prob": 1.0, "iou_best": 5.0, "iou_normal": 1.0, "coords": 1.0}


def _load(name):
    d = np.load(GOLD)
    labels = tuple(d["%s_label_%s" % (name, n)] for n in ("mask", "prob", "coords", "offset_xy_min", "offset_xy_max", "areas"))
    return d, d[name + "_net"], int(d[name + "_meta"][0]), d[name + "_anchors"], labels


@pytest.mark.parametrize("name", CASES)
def test_decode_oracle_matches_the_reference_source(name):
    d, net, classes, anchors, _ = _load(name)
    m = ho.decode_oracle(net, classes, anchors, dtype=np.float64)
    for k in ATTRS:
        want = d["%s_f64_model_%s" % (name, k)]
        assert m[k].shape == want.shape, k
        np.testing.assert_allclose(m[k], want, rtol=1e-12, atol=1e-14, err_msg=k)
    m32 = ho.decode_oracle(net, classes, anchors)                     # the float32 path the GPU tests compare against
    for k in ("conf", "xy_min", "xy_max", "coords", "prob"):
        want = d["%s_f64_model_%s" % (name, k)]
        assert np.abs(m32[k] - want).max() <= 2e-6 * max(1.0, np.abs(want).max()), k


@pytest.mark.parametrize("name", CASES)
def test_objectives_and_gradient_match_the_reference_source(name):
    d, net, classes, anchors, labels = _load(name)
    obj, g = ho.loss_grad_oracle(net, classes, anchors, labels, hparam=HPARAM, dtype=np.float64)
    for k in HPARAM:
        np.testing.assert_allclose(float(obj[k]), float(d["%s_f64_obj_%s" % (name, k)]), rtol=1e-12, err_msg=k)
    want = d[name + "_f64_grad"]
    assert g.shape == want.shape
    np.testing.assert_allclose(g, want, rtol=1e-10, atol=1e-16)         # closed form == autodiff through the reference's source
    obj32, g32 = ho.loss_grad_oracle(net, classes, anchors, labels, hparam=HPARAM)
    for k in HPARAM:
        assert abs(float(obj32[k]) - float(d["%s_f32_obj_%s" % (name, k)])) <= 2e-6 * abs(float(d["%s_f64_obj_%s" % (name, k)])) + 1e-9
        assert abs(float(obj32[k]) - float(d["%s_f64_obj_%s" % (name, k)])) <= 1e-5 * abs(float(d["%s_f64_obj_%s" % (name, k)])) + 1e-9
    assert np.abs(g32 - want).max() <= 1e-5 * np.abs(want).max()
    assert np.abs(d[name + "_f32_grad"] - want).max() <= 1e-5 * np.abs(want).max()   # float32 autodiff has the same noise


def test_ties_case_really_has_tied_best_boxes_and_the_empty_image_none():
    d, net, classes, anchors, labels = _load("ties")
    m = ho.decode_oracle(net, classes, anchors, training=True, dtype=np.float64)
    _, aux = ho.objectives_oracle(m, labels, dtype=np.float64)
    mb = aux["mask_best"]
    assert (mb.sum(-1) == 2).any()              # two anchors share the maximum IoU in some object cell: both count (tf.equal)
    d, net, classes, anchors, labels = _load("voc")
    assert labels[0][1].sum() == 0 and labels[0][0].sum() > 0          # second image of the batch has no object


@pytest.mark.skipif(not os.path.exists("/root/reference/model/yolo2/__init__.py"), reason="the reference checkout only exists in the authoring container")
def test_head_oracle_against_the_live_reference_source_on_random_cases():
    """Beyond the 4 committed cases: 12 fresh random heads per run through the reference's own Model / Objectives source (the
    generator's stand-in for TF, float64): grids 1 x 1 .. 9 x 9 incl. non-square, 1 .. 6 anchors, 1 .. 30 classes, 0 .. 12 objects."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    try:
        import make_head_golden as mh
        from make_labels_golden import load_reference_transform_labels
    finally:
        sys.path.pop(0)
    import torch
    tl = load_reference_transform_labels()
    rs = np.random.RandomState(123)
    for case in range(12):
        b, hc, wc, a, c = int(rs.randint(1, 4)), int(rs.randint(1, 10)), int(rs.randint(1, 10)), int(rs.randint(1, 7)), int(rs.randint(1, 31))
        anchors = rs.uniform(0.5, 8.0, size=(a, 2))
        net = rs.normal(0, 1.2, size=(b, hc, wc, a * (5 + c))).astype(np.float32)
        per = []
        for _ in range(b):
            n = int(rs.randint(0, 13))
            cx, cy, w, h = rs.uniform(0, 1, n), rs.uniform(0, 1, n), rs.uniform(0.05, 0.6, n), rs.uniform(0.05, 0.6, n)
            coord = np.stack([np.clip(cx - w / 2, 0, 1 - 1e-6), np.clip(cy - h / 2, 0, 1 - 1e-6), np.clip(cx + w / 2, 0, 1 - 1e-6),
                              np.clip(cy + h / 2, 0, 1 - 1e-6)], 1).astype(np.float32).reshape(n, 4)
            per.append(tl(rs.randint(0, c, n), coord, c, wc, hc))
        labels = tuple(np.stack([p[i] for p in per], 0) for i in range(6))
        attrs, obj, grad = mh.run_reference(net, c, anchors, labels, torch.float64)
        m = ho.decode_oracle(net, c, anchors, dtype=np.float64)
        for k in ATTRS:
            np.testing.assert_allclose(m[k], attrs[k], rtol=1e-12, atol=1e-13, err_msg="%d %s" % (case, k))
        o, g = ho.loss_grad_oracle(net, c, anchors, labels, hparam=HPARAM, dtype=np.float64)
        for k in HPARAM:
            np.testing.assert_allclose(float(o[k]), float(obj[k]), rtol=1e-11, atol=1e-18, err_msg="%d %s" % (case, k))
        np.testing.assert_allclose(g, grad, rtol=1e-9, atol=1e-15, err_msg=str(case))
