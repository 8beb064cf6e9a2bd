"""Generates tests/golden/nms_*.npz by running the REFERENCE'S OWN NMS
(/root/reference/utils/postprocess.py, loaded by file path because the package
__init__ imports TensorFlow) in the authoring container.  The outputs pin
oracle/nms_oracle.{py,c}.  Run once, here:  python tests/golden/make_nms_golden.py

Environment recorded in each file: numpy version (NEP-50 float32 semantics matter).
/root/reference does not exist on the GPU box; only the committed .npz travel.
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/utils/postprocess.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_postprocess", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def boxes_from_grid(rs, hc, wc, anchors, spread=0.5):
    a = len(anchors)
    gy, gx = np.meshgrid(np.arange(hc), np.arange(wc), indexing="ij")
    centre = np.stack([gx, gy], -1).reshape(hc * wc, 1, 2) + rs.uniform(0, 1, size=(hc * wc, a, 2))
    wh = np.asarray(anchors).reshape(1, a, 2) * np.exp(rs.normal(0, spread, size=(hc * wc, a, 2)))
    lo = (centre - wh / 2).astype(np.float32)
    hi = (centre + wh / 2).astype(np.float32)
    return lo, hi


def scores(rs, cells, a, c, k, quant=None):
    """k (box,class) entries ~U(0.3,1) (above threshold), rest ~U(0,0.29)."""
    s = rs.uniform(0, 0.29, size=(cells * a * c))
    pick = rs.choice(s.size, size=min(k, s.size), replace=False)
    s[pick] = rs.uniform(0.3, 1.0, size=len(pick))
    s = s.astype(np.float32)
    if quant:
        s = (np.round(s * quant) / quant).astype(np.float32)
    return s.reshape(cells, a, c)


ANCH = [[1.08, 1.19], [3.42, 4.41], [6.63, 11.38], [9.42, 5.11], [16.62, 10.52]]


def cases():
    rs = np.random.RandomState(1234)
    out = {}
    # 1. tiny, heavy ties (scores quantised to 1/8), 3 classes
    lo, hi = boxes_from_grid(rs, 3, 3, ANCH[:2], 0.3)
    out["tiny_ties"] = (scores(rs, 9, 2, 3, 20, quant=8), lo, hi, 0.3, 0.4)
    # 2. single class
    lo, hi = boxes_from_grid(rs, 4, 4, ANCH[:3], 0.3)
    out["one_class"] = (scores(rs, 16, 3, 1, 12), lo, hi, 0.3, 0.4)
    # 3. nothing above threshold
    lo, hi = boxes_from_grid(rs, 4, 4, ANCH[:2], 0.3)
    out["all_below"] = (rs.uniform(0, 0.29, size=(16, 2, 4)).astype(np.float32), lo, hi, 0.3, 0.4)
    # 4. all boxes identical, all scores equal (maximum ties, IoU == 1)
    lo = np.tile(np.array([1.0, 1.0], np.float32), (12, 1, 1)).reshape(6, 2, 2)
    hi = np.tile(np.array([3.0, 2.5], np.float32), (12, 1, 1)).reshape(6, 2, 2)
    out["identical"] = (np.full((6, 2, 3), 0.5, np.float32), lo, hi, 0.3, 0.4)
    # 5. values exactly on both thresholds: score == f32(0.3); boxes with IoU == 0.5 at thr_iou 0.5
    lo = np.array([[0, 0], [1, 0], [0, 0], [5, 5]], np.float32).reshape(4, 1, 2)
    hi = np.array([[2, 1], [3, 1], [2, 1], [6, 6]], np.float32).reshape(4, 1, 2)
    sc = np.array([[0.9, np.float32(0.3)], [0.8, 0.9], [np.float32(0.3), 0.31], [0.29, 0.95]], np.float32).reshape(4, 1, 2)
    out["on_threshold"] = (sc, lo, hi, 0.3, 1.0 / 3.0)
    # 6. realistic 13x13x5, 20 classes, ~150 candidates, quantised to 1/64 (some ties)
    lo, hi = boxes_from_grid(rs, 13, 13, ANCH, 0.5)
    out["grid13_c20"] = (scores(rs, 169, 5, 20, 150, quant=64), lo, hi, 0.3, 0.4)
    # 7. realistic 13x13x5, 80 classes, ~200 candidates, no quantisation
    lo, hi = boxes_from_grid(rs, 13, 13, ANCH, 0.5)
    out["grid13_c80"] = (scores(rs, 169, 5, 80, 200), lo, hi, 0.3, 0.4)
    # 8. dense single class (long suppression chains), 7x7x5
    lo, hi = boxes_from_grid(rs, 7, 7, ANCH, 0.2)
    out["dense_chain"] = (scores(rs, 49, 5, 2, 300), lo, hi, 0.3, 0.4)
    # 9. negative and zero scores present (zeroing a negative changes it)
    lo, hi = boxes_from_grid(rs, 4, 4, ANCH[:3], 0.3)
    sc = scores(rs, 16, 3, 3, 25)
    sc[::3] *= -1
    out["negatives"] = (sc, lo, hi, 0.3, 0.4)
    return out


def main():
    ref = load_reference()
    for name, (conf, lo, hi, thr, thr_iou) in cases().items():
        conf_in = conf.copy()
        work = conf.copy()
        boxes = ref.non_max_suppress(work, lo, hi, thr, thr_iou)
        base = work.__array_interface__["data"][0]
        c = work.shape[-1]
        order = np.array([(b[0].__array_interface__["data"][0] - base) // (4 * c) for b in boxes], dtype=np.int32)
        assert sorted(order.tolist()) == list(range(len(order)))
        path = os.path.join(HERE, "nms_%s.npz" % name)
        np.savez_compressed(path, conf_in=conf_in, xy_min=lo, xy_max=hi, threshold=np.float64(thr),
                            threshold_iou=np.float64(thr_iou), conf_out=work, order=order,
                            numpy_version=np.array(np.__version__))
        print(name, conf.shape, "zeroed:", int((conf_in != work).sum()), "->", os.path.basename(path))


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("reference not present; golden files are generated in the authoring container only")
    main()
