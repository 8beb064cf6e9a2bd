"""Host-side model of nms_apply_kernel's work decomposition (yolo_tf_b200/csrc/y2_nms.cu), checked against the oracle
without a GPU: boxes of an image in area order (stable counting sort on the top 11 bits of the order-preserving bit pattern of
the area), tiles of 128 of them, register slots of 32, the per-slot area window that drops (kept box, slot) pairs, the slot
masks attached to staged kept boxes, the per-class relevance gate for columns with at most 8 kept boxes, the suppressed-candidate
lookup.  If any of these could change a result, the model's output would differ from the reference algorithm's
(oracle/nms_oracle.c through oracle/nms_c.py) somewhere on these inputs."""
import numpy as np
import pytest

from oracle.nms_c import nms_c_batch
from test_nms_cull_bound import f32, iou32          # tests/ is on sys.path (conftest.py)

AP_BOXES, AP_PF = 128, 8


def ford(v):
    u = (v + f32(0)).astype(f32).view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000))


def apply_model(conf, lo, hi, thr, thr_iou, kept, supp):
    """conf [N][C] original scores; kept[c] = indices of class c's kept boxes (visiting order); supp[c] = set of suppressed candidates."""
    N, C = conf.shape
    out = conf.copy()
    with np.errstate(all="ignore"):
        area = ((hi[:, 0] - lo[:, 0]).astype(f32) * (hi[:, 1] - lo[:, 1]).astype(f32)).astype(f32)
    perm = np.argsort(ford(area) >> np.uint32(21), kind="stable")
    quick = thr_iou > 0
    tc = f32(f32(0.998) * f32(thr_iou))
    tests = culled = 0
    for n0 in range(0, N, AP_BOXES):
        rows = perm[n0:n0 + AP_BOXES]
        slots = [rows[h:h + 32] for h in range(0, len(rows), 32)]
        win = []
        for sl in slots:
            a = area[sl]
            with np.errstate(all="ignore"):
                mn, mx = np.fmin.reduce(a), np.fmax.reduce(a)
                win.append((f32(tc * mn), f32(mx / tc)) if quick else (f32(-np.inf), f32(np.inf)))
        lo_t, hi_t = np.fmin.reduce([w[0] for w in win]), np.fmax.reduce([w[1] for w in win])
        for c in range(C):
            k = kept[c]
            if len(k) == 0:
                continue
            for n in rows:                                            # phase 1: candidates look their fate up
                if conf[n, c] > thr and n in supp[c]:
                    out[n, c] = 0.0
            ka = area[k]
            with np.errstate(invalid="ignore"):
                if len(k) <= AP_PF and not np.any((ka >= lo_t) & (ka <= hi_t)):
                    culled += len(k) * len(rows)
                    continue                                          # relevance gate
                for sl, (wl, wh) in zip(slots, win):
                    live = k[(ka >= wl) & (ka <= wh)]                 # the slot's bit of the staged kept boxes' masks
                    culled += (len(k) - len(live)) * len(sl)
                    for n in sl:
                        v = conf[n, c]
                        if v > thr or v.view(np.uint32) == 0:
                            continue                                  # candidates; scores that are +0.0 already
                        tests += len(live)
                        if len(live):
                            i, _, _ = iou32(lo[live], hi[live], lo[n][None], hi[n][None])
                            if np.any(i >= f32(thr_iou)):
                                out[n, c] = 0.0
    return out, tests, culled


def inputs(rs, g, C, K, quant=None, degenerate=False):
    anchors = np.array([[0.74, 0.87], [2.42, 2.66], [4.31, 7.04], [10.25, 4.59], [12.69, 11.87]])
    cells = g * g
    gy, gx = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    centre = np.stack([gx, gy], -1).reshape(cells, 1, 2) + rs.uniform(0, 1, size=(cells, 5, 2))
    wh = anchors.reshape(1, 5, 2) * np.exp(rs.normal(0, 0.5, size=(cells, 5, 2)))
    N = cells * 5
    if degenerate:
        wh.reshape(N, 2)[5:25, 1] = 0.0                               # zero-area boxes
        wh.reshape(N, 2)[30:34] *= 1e25                               # areas overflow
        dup = rs.randint(0, N, size=60)
        centre.reshape(N, 2)[dup[:30]] = centre.reshape(N, 2)[dup[30:]]
        wh.reshape(N, 2)[dup[:30]] = wh.reshape(N, 2)[dup[30:]]       # identical boxes
    lo = (centre - wh / 2).astype(f32).reshape(N, 2)
    hi = (centre + wh / 2).astype(f32).reshape(N, 2)
    conf = rs.uniform(-0.05, 0.29, size=N * C)
    conf[rs.choice(N * C, size=K, replace=False)] = rs.uniform(0.3, 1.0, size=K)
    conf[rs.choice(N * C, size=N * C // 10, replace=False)] = 0.0
    conf[rs.choice(N * C, size=50, replace=False)] = -0.0
    conf = conf.astype(f32)
    if quant:
        conf = (np.round(conf * quant) / quant).astype(f32)
    return conf.reshape(N, C), lo, hi


@pytest.mark.parametrize("g,C,K,thr_iou,quant,degenerate", [(13, 6, 150, 0.4, None, False), (13, 4, 600, 0.4, 16, False),
                                                          (7, 5, 40, 0.6, None, True), (13, 3, 200, 0.0, None, False),
                                                          (9, 4, 120, 0.05, 8, True)])
def test_apply_decomposition_gives_the_oracles_bits(g, C, K, thr_iou, quant, degenerate):
    rs = np.random.RandomState(g * 100 + C)
    conf, lo, hi = inputs(rs, g, C, K, quant, degenerate)
    thr = 0.3
    ref = conf[None].copy()
    with np.errstate(all="ignore"):
        order = nms_c_batch(ref, lo[None], hi[None], thr, thr_iou)
    ref = ref[0]
    # what the select kernel hands to apply: kept lists and suppressed candidates per class (from the oracle's result)
    kept, supp = [], []
    for c in range(C):
        cand = conf[:, c] > f32(thr)
        kept.append(np.flatnonzero(cand & (ref[:, c] > f32(thr))))
        supp.append(set(np.flatnonzero(cand & ~(ref[:, c] > f32(thr))).tolist()))
    got, tests, culled = apply_model(conf, lo, hi, f32(thr), thr_iou, kept, supp)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    if thr_iou >= 0.4 and not degenerate:
        assert culled > tests                                         # and it is worth it: most (kept box, box) pairs never get tested


def test_overflowing_areas_follow_numpy_not_fmax():
    """Boxes whose areas overflow float32: areas1 + areas2 - inter = inf + inf - inf = nan, np.maximum(nan, 1e-10) = nan
    (utils/postprocess.py:36), nan >= threshold_iou is False -- the reference does NOT suppress such a pair, whereas a C-style
    max would turn the denominator into 1e-10 and the pair into a hit.  The model above found the C oracle (and with it the
    kernels, which shared its fmaxf) on the wrong side of this; all three now propagate the NaN.  Python oracle (the reference's
    own numpy calls) vs C oracle, bit for bit."""
    from oracle.nms_oracle import nms_oracle
    rs = np.random.RandomState(4)
    n, C = 24, 3
    c = rs.uniform(0, 5, size=(n, 2))
    wh = np.exp(rs.uniform(-1, 1.5, size=(n, 2)))
    wh[3:9] *= 1e25                                                   # six boxes with area = inf, nested in each other
    lo, hi = (c - wh / 2).astype(f32), (c + wh / 2).astype(f32)
    conf = rs.uniform(0, 0.29, size=(n, C)).astype(f32)
    conf[[4, 6, 10, 15], :] = rs.uniform(0.5, 1.0, size=(4, C)).astype(f32)
    a = conf.copy().reshape(n, 1, C)
    with np.errstate(all="ignore"):
        order_py = nms_oracle(a, lo.reshape(n, 1, 2), hi.reshape(n, 1, 2), 0.3, 0.4)
        b = conf[None].copy()
        order_c = nms_c_batch(b, lo[None], hi[None], 0.3, 0.4)
    assert np.array_equal(a.reshape(n, C).view(np.uint32), b[0].view(np.uint32))
    assert np.array_equal(np.asarray(order_py), order_c[0])
    assert (a.reshape(n, C)[[3, 5, 7, 8]] != 0).all()                 # the other inf-area boxes are not suppressed by 4 and 6
