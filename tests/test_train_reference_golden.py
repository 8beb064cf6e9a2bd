"""CPU (-m "not gpu"): pins the TRAINING-STEP oracle (oracle/train_oracle.py: batch-statistics forward, 4-part loss, backward
through everything) to one training step of the REFERENCE'S OWN source -- `darknet(..., training=True)`, `Model`,
`Objectives`, the [yolo2_hparam] weighting -- run with torch float64 stand-ins for the slim / tf calls and torch autograd in
the place of tf.gradients (tests/golden/make_train_golden.py -> tests/golden/train_reference.npz).  Loss, objectives, network
output, d(total)/d(net) in full; the gradient of each of the 65 trainable variables by norm, leading entries and a random
projection."""
import os

import numpy as np

from oracle import head_oracle as ho
from oracle.darknet_oracle import init_params
from oracle.train_oracle import train_step_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_reference.npz")
HPARAM = {"prob": 1.0, "iou_best": 5.0, "iou_normal": 1.0, "coords": 1.0}


def _summary(g):
    flat = np.asarray(g, dtype=np.float64).reshape(-1)
    probe = np.random.RandomState(flat.size % (2 ** 31)).normal(size=flat.size)
    return np.concatenate([[np.sqrt((flat ** 2).sum())], flat[:8] if flat.size >= 8 else np.pad(flat, (0, 8 - flat.size)), [flat @ probe]])


def test_train_step_oracle_matches_one_step_of_the_reference_source():
    d = np.load(GOLD)
    classes, anchors_n, seed = (int(v) for v in d["meta"])
    x = d["x"]
    labels = ho.synthetic_labels(x.shape[0], classes, x.shape[2] // 32, x.shape[1] // 32, seed=seed)
    ref = train_step_oracle(x, init_params(classes, anchors_n, seed=1), classes, ho.ANCHORS_VOC, labels, HPARAM)
    np.testing.assert_allclose(ref["total"], float(d["total"]), rtol=1e-10)
    for k in HPARAM:
        np.testing.assert_allclose(ref["objectives"][k], float(d["obj_" + k]), rtol=1e-10, err_msg=k)
    np.testing.assert_allclose(ref["net"], d["net"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(ref["dnet"], d["dnet"], rtol=1e-8, atol=1e-15)
    names = [str(n) for n in d["grad_names"]]
    assert len(names) == 65 and set(n[len("yolo2_darknet/"):] for n in names) == set(ref["grads"])   # 21 x (w, gamma, beta) + (w, b)
    for n, want in zip(names, d["grad_summary"]):
        got = _summary(ref["grads"][n[len("yolo2_darknet/"):]])
        scale = max(want[0], 1e-30)                                 # the variable's gradient norm
        assert np.abs(got - want).max() <= 1e-7 * scale * max(1.0, np.sqrt(ref["grads"][n[len("yolo2_darknet/"):]].size) * 1e-2), n


def test_tiny_train_step_oracle_matches_one_step_of_the_reference_source():
    """The same pin for `tiny()` (model/yolo2/inference.py:25-50): batch-statistics BN through its 9 convs, the 2x2 stride-1
    SAME max-pool (:42) and its gradient (overlapping windows), loss and backward -- 26 trainable variables."""
    from oracle.darknet_oracle import tiny_layer_table
    d = np.load(GOLD)
    classes, anchors_n, seed = (int(v) for v in d["tiny_meta"])
    x = d["tiny_x"]
    labels = ho.synthetic_labels(x.shape[0], classes, x.shape[2] // 32, x.shape[1] // 32, seed=seed)
    table = tiny_layer_table(classes, anchors_n)
    ref = train_step_oracle(x, init_params(classes, anchors_n, seed=1, table=table), classes, ho.ANCHORS_VOC, labels, HPARAM, table=table)
    np.testing.assert_allclose(ref["total"], float(d["tiny_total"]), rtol=1e-10)
    for k in HPARAM:
        np.testing.assert_allclose(ref["objectives"][k], float(d["tiny_obj_" + k]), rtol=1e-10, err_msg=k)
    np.testing.assert_allclose(ref["net"], d["tiny_net"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(ref["dnet"], d["tiny_dnet"], rtol=1e-8, atol=1e-15)
    names = [str(n) for n in d["tiny_grad_names"]]
    assert len(names) == 26 and set(n[len("yolo2_tiny/"):] for n in names) == set(ref["grads"])      # 8 x (w, gamma, beta) + (w, b)
    for n, want in zip(names, d["tiny_grad_summary"]):
        g = ref["grads"][n[len("yolo2_tiny/"):]]
        scale = max(want[0], 1e-30)
        assert np.abs(_summary(g) - want).max() <= 1e-7 * scale * max(1.0, np.sqrt(g.size) * 1e-2), n
