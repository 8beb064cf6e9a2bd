"""CPU (-m "not gpu"): the visiting order the select kernel builds (yolo_tf_b200/csrc/y2_nms.cu: 64-bit keys
(ford(value) << 32) | (0xffff - index) sorted descending, then runs of EQUAL values re-ranked with precedes(): earlier class
columns descending, index ascending) against what the reference does (utils/postprocess.py:42-44): ONE Python list of boxes,
stable-sorted by class 0 descending, then by class 1, ... -- ties keep the order the previous classes left.  The model is numpy /
pure Python; the kernels are checked against the oracle on the GPU, this checks that the ordering RULE they implement is the
reference's, with heavy ties, -0.0 / +0.0 and negative scores."""
import numpy as np
import pytest


def ford(v):
    u = (np.float32(v) + np.float32(0)).view(np.uint32)
    return int(~u & np.uint32(0xffffffff)) if u & np.uint32(0x80000000) else int(u | np.uint32(0x80000000))


def precedes(conf, c, i, j):
    """does box i come before box j in class c's order, given conf[i, c] == conf[j, c]?  (y2_nms.cu:precedes, value part done)"""
    for cc in range(c - 1, -1, -1):
        if conf[i, cc] > conf[j, cc]:
            return True
        if conf[i, cc] < conf[j, cc]:
            return False
    return i < j


def kernel_order(conf, c, thr):
    cand = [int(n) for n in np.flatnonzero(conf[:, c] > np.float32(thr))]
    keys = sorted(((ford(conf[n, c]) << 32) | (0xffff - n) for n in cand), reverse=True)
    order = [0xffff - (k & 0xffff) for k in keys]
    out, p = list(order), 0
    while p < len(order):                                              # re-rank every run of equal values
        q = p
        while q < len(order) and conf[order[q], c] == conf[order[p], c]:
            q += 1
        run = order[p:q]
        for n in run:
            out[p + sum(1 for m in run if m != n and precedes(conf, c, m, n))] = n
        p = q
    return out


@pytest.mark.parametrize("quant,seed", [(4, 0), (8, 1), (2, 2), (None, 3)])
def test_key_sort_plus_tie_rerank_is_the_references_stable_sort(quant, seed):
    rs = np.random.RandomState(seed)
    N, C, thr = 300, 6, 0.3
    conf = rs.uniform(-0.2, 1.0, size=(N, C))
    if quant:
        conf = np.round(conf * quant) / quant
    conf = conf.astype(np.float32)
    conf[rs.choice(N, 20, replace=False), rs.randint(0, C, 20)] = -0.0
    boxes = list(range(N))                                             # the reference's list, carried from class to class
    for c in range(C):
        boxes.sort(key=lambda n: conf[n, c], reverse=True)             # postprocess.py:43 (stable; reverse keeps ties in place)
        want = [n for n in boxes if conf[n, c] > np.float32(thr)]
        assert kernel_order(conf, c, thr) == want, c
