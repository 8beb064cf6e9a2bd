"""CPU oracle for greedy per-class NMS.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product path
(``yolo_tf_b200``) never does.

Restates the algorithm of the reference's ``utils/postprocess.py``:

* ``pair_iou``      follows ``utils/postprocess.py:21-36`` (``iou``): the exact
  float32 operation order ``a1=(x1max-x1min)*(y1max-y1min)``, ``a2`` likewise,
  ``iw=max(min(x1max,x2max)-max(x1min,x2min),0)``, ``ih`` likewise,
  ``inter=iw*ih``, ``den=max((a1+a2)-inter, f32(1e-10))``, ``iou=inter/den``.
* ``nms_oracle``    follows ``utils/postprocess.py:39-51``
  (``non_max_suppress``): one box list carried across classes, Python *stable*
  sort, descending, by the class column; every box whose score is ``> threshold``
  when it is reached zeroes the class score of every LATER box (candidate or
  not) with ``iou >= threshold_iou``; mutation happens in the caller's array.

Parity pin: ``tests/golden/make_nms_golden.py`` ran the reference's own file
(loaded by path from /root/reference) in the authoring container (NumPy 2.3,
NEP-50 float32 semantics) and committed inputs + outputs under
``tests/golden/nms_*.npz``; ``tests/test_oracle_golden.py`` checks this module
and the C twin (``nms_oracle.c``) against them bit for bit.

Implementation differs from the reference in form (index permutation instead of
a list of row views; the inner "every later box" loop is evaluated as float32
numpy vectors, which is elementwise-identical IEEE arithmetic), not in results.
"""
import numpy as np

F32 = np.float32


def pair_iou(min_a, max_a, mins_b, maxs_b):
    """IoU of one box against a vector of boxes, float32, reference op order
    (utils/postprocess.py:28-36)."""
    area_a = (max_a[0] - min_a[0]) * (max_a[1] - min_a[1])
    area_b = (maxs_b[:, 0] - mins_b[:, 0]) * (maxs_b[:, 1] - mins_b[:, 1])
    iw = np.maximum(np.minimum(max_a[0], maxs_b[:, 0]) - np.maximum(min_a[0], mins_b[:, 0]), F32(0))
    ih = np.maximum(np.minimum(max_a[1], maxs_b[:, 1]) - np.maximum(min_a[1], mins_b[:, 1]), F32(0))
    inter = iw * ih
    den = np.maximum((area_a + area_b) - inter, F32(1e-10))
    return inter / den


def nms_oracle(conf, xy_min, xy_max, threshold, threshold_iou):
    """In-place greedy NMS on one image.

    conf [cells, A, C] float32 (mutated), xy_min/xy_max [cells, A, 2] float32.
    Returns the final box permutation (int64 [N]) = the order of the list the
    reference returns (utils/postprocess.py:51).
    """
    assert conf.dtype == np.float32 and xy_min.dtype == np.float32 and xy_max.dtype == np.float32
    classes = conf.shape[-1]
    score = conf.reshape(-1, classes)           # view: writes land in caller's array
    assert np.shares_memory(score, conf)
    lo = xy_min.reshape(-1, 2)
    hi = xy_max.reshape(-1, 2)
    n = score.shape[0]
    thr = F32(threshold)                        # NEP-50: python float is weak -> f32 compare
    thr_iou = F32(threshold_iou)
    order = list(range(n))
    for c in range(classes):
        # stable, descending; ties keep the order left by the previous class
        order.sort(key=lambda b: score[b, c], reverse=True)
        perm = np.asarray(order, dtype=np.int64)
        lo_p = lo[perm]
        hi_p = hi[perm]
        for i in range(n - 1):
            b = perm[i]
            if score[b, c] <= thr:
                continue
            later = perm[i + 1:]
            # reference asserts (postprocess.py:22-27): no NaN, min <= max
            assert not np.isnan(lo_p[i:]).any() and not np.isnan(hi_p[i:]).any()
            assert np.all(lo_p[i:] <= hi_p[i:])
            v = pair_iou(lo_p[i], hi_p[i], lo_p[i + 1:], hi_p[i + 1:])
            hit = v >= thr_iou
            if hit.any():
                score[later[hit], c] = 0
    return np.asarray(order, dtype=np.int64)


def nms_oracle_batch(conf, xy_min, xy_max, threshold, threshold_iou):
    """Batch wrapper: conf [B, N, C] (mutated), boxes [B, N, 2]; returns [B, N] order."""
    out = np.empty(conf.shape[:2], dtype=np.int64)
    for b in range(conf.shape[0]):
        out[b] = nms_oracle(conf[b][:, None, :], xy_min[b][:, None, :], xy_max[b][:, None, :],
                            threshold, threshold_iou)
    return out


def detections(conf_row_order, conf, xy_min, xy_max, threshold):
    """What detect.py:78-80 keeps after NMS: boxes whose max class score is
    > threshold, with that argmax class.  Returns (box index, class, score)."""
    score = conf.reshape(-1, conf.shape[-1])
    keep = []
    for b in conf_row_order:
        k = int(np.argmax(score[b]))
        if score[b, k] > F32(threshold):
            keep.append((int(b), k, float(score[b, k])))
    return keep
