"""transform_labels (utils/data/__init__.py:112-145).  The goldens in tests/golden/labels.npz were produced
by the reference's own function (tests/golden/make_labels_golden.py); CPU tests pin the oracle to them, -m gpu tests pin
the device encoder to them bit for bit and to the oracle on a large synthetic batch."""
import os

import numpy as np
import pytest

from oracle import head_oracle as ho

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "labels.npz")
KEYS = ("mask", "prob", "coords", "offset_xy_min", "offset_xy_max", "areas")
CASES = ("voc13", "coco19", "single", "empty", "nonsquare", "same_cell", "crowd", "borders")


def _case(g, name):
    classes, cw, ch = [int(v) for v in g[name + "_meta"]]
    return g[name + "_class"], g[name + "_coord"], classes, cw, ch, [g[name + "_" + k] for k in KEYS]


@pytest.mark.parametrize("name", CASES)
def test_label_oracle_matches_reference_golden(name):
    g = np.load(GOLD)
    cls, coord, classes, cw, ch, want = _case(g, name)
    got = ho.transform_labels_oracle(cls, coord, classes, cw, ch)
    for k, a, b in zip(KEYS, got, want):
        assert a.dtype == np.float32 and a.shape == b.shape, k
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, k)


def test_golden_covers_collisions():
    g = np.load(GOLD)
    # 5 objects in 3 cells: objects 0, 1, 2 (classes 1, 7, 1) share cell 97 -> class bits {1, 7}, box = object 2 (the last)
    assert g["same_cell_mask"].sum() == 3 and g["same_cell_prob"].sum() == 4
    assert np.array_equal(g["same_cell_coords"][97, 0, 2:], np.sqrt(g["same_cell_coord"][2, 2:] - g["same_cell_coord"][2, :2]))
    assert g["crowd_mask"].sum() < 169 and len(g["crowd_class"]) == 300


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_encoder_bit_exact_vs_reference_golden(cuda, name):
    from yolo_tf_b200.utils.data import transform_labels
    g = np.load(GOLD)
    cls, coord, classes, cw, ch, want = _case(g, name)
    got = transform_labels(cls, coord, classes, cw, ch)
    for k, a, b in zip(KEYS, got, want):
        assert a.shape == b.shape and a.dtype == np.float32, k
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, k)


@pytest.mark.gpu
def test_device_encoder_batch_vs_oracle_and_loss(cuda):
    """BASELINE config 3's label stream (B=64, 13x13, 20 classes) encoded in one launch: bit-exact against the oracle image by
    image, and the loss kernel accepts the tensors as they are."""
    import torch
    from yolo_tf_b200.model.yolo2 import Model, Objectives
    from yolo_tf_b200.utils.data import transform_labels_batch
    rs = np.random.RandomState(31)
    B, C, cw, ch = 64, 20, 13, 13
    cls, xy = [], []
    for _ in range(B):
        n = rs.randint(0, 9)                                  # ragged, some images empty
        cx, cy, w, h = rs.uniform(0, 1, n), rs.uniform(0, 1, n), rs.uniform(0.05, 0.6, n), rs.uniform(0.05, 0.6, n)
        xy.append(np.stack([np.clip(cx - w / 2, 0, 1 - 1e-6), np.clip(cy - h / 2, 0, 1 - 1e-6), np.clip(cx + w / 2, 0, 1 - 1e-6),
                            np.clip(cy + h / 2, 0, 1 - 1e-6)], 1).astype(np.float32).reshape(-1, 4))
        cls.append(rs.randint(0, C, n))
    got = transform_labels_batch(cls, xy, C, cw, ch)
    for b in range(B):
        want = ho.transform_labels_oracle(cls[b], xy[b], C, cw, ch)
        for k, a, w_ in zip(KEYS, got, want):
            assert np.array_equal(a[b].cpu().numpy().view(np.uint32), w_.view(np.uint32)), (b, k)
    net = rs.normal(0, 1, size=(B, ch, cw, 5 * (5 + C))).astype(np.float32)
    model = Model(torch.from_numpy(net).to(cuda), C, ho.ANCHORS_VOC, training=True)
    obj = Objectives(model, *got, hparam=ho.HPARAM_DEFAULT)
    ref_obj, _ = ho.loss_grad_oracle(net, C, ho.ANCHORS_VOC, tuple(t.cpu().numpy() for t in got), dtype=np.float64)
    for k in ref_obj:
        assert abs(float(obj[k]) - ref_obj[k]) <= 1e-4 * abs(ref_obj[k]), k


@pytest.mark.gpu
def test_device_encoder_errors_like_the_reference(cuda):
    from yolo_tf_b200.utils.data import transform_labels
    with pytest.raises(IndexError):                           # centre on the bottom-right corner: index == cells
        transform_labels(np.array([0]), np.array([[1.0, 1.0, 1.0, 1.0]], np.float32), 20, 13, 13)
    with pytest.raises(IndexError):                           # class out of range
        transform_labels(np.array([20]), np.array([[0.1, 0.1, 0.2, 0.2]], np.float32), 20, 13, 13)
    with pytest.raises(AssertionError):                       # xmax < xmin (:142)
        transform_labels(np.array([0]), np.array([[0.5, 0.1, 0.2, 0.2]], np.float32), 20, 13, 13)


@pytest.mark.skipif(not os.path.exists("/root/reference/utils/data/__init__.py"), reason="the reference checkout only exists in the authoring container")
def test_label_oracle_against_the_live_reference_on_random_cases():
    """Beyond the committed goldens: 60 fresh random object lists per run through the reference's own transform_labels (its source
    compiled from the file, tests/golden/make_labels_golden.py) -- crowded cells, square and non-square grids, 1 .. 80 classes."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    try:
        from make_labels_golden import load_reference_transform_labels
    finally:
        sys.path.pop(0)
    from oracle.head_oracle import transform_labels_oracle
    tl = load_reference_transform_labels()
    rs = np.random.RandomState(99)
    for case in range(60):
        classes, cw, ch = int(rs.choice([1, 3, 20, 80])), int(rs.randint(1, 20)), int(rs.randint(1, 20))
        n = int(rs.randint(0, 40))
        cx, cy = rs.uniform(0, 1, n), rs.uniform(0, 1, n)
        w, h = rs.uniform(0.01, 0.7, n), rs.uniform(0.01, 0.7, n)
        coord = np.stack([np.clip(cx - w / 2, 0, 1 - 1e-6), np.clip(cy - h / 2, 0, 1 - 1e-6), np.clip(cx + w / 2, 0, 1 - 1e-6),
                          np.clip(cy + h / 2, 0, 1 - 1e-6)], 1).astype(np.float32).reshape(n, 4)
        cls = rs.randint(0, classes, n)
        want = tl(cls, coord, classes, cw, ch)
        got = transform_labels_oracle(cls, coord, classes, cw, ch)
        for a, b in zip(got, want):
            assert a.shape == b.shape and np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32)), case
