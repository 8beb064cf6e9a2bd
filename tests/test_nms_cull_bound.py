"""The area-window culling of nms_apply_kernel (yolo_tf_b200/csrc/y2_nms.cu) drops a (kept box, register slot) pair when
    area_kept < f32(f32(0.998 thr) * a_min)   or   area_kept > f32(a_max / f32(0.998 thr)),
a_min / a_max = the smallest / largest area in the slot.  It may only ever drop pairs the reference's float32 iou()
(utils/postprocess.py:21-36) would answer "no hit" for.  This CPU test restates the bound in numpy float32 and fuzzes it against the
oracle's own iou arithmetic on the pairs where it is tight: nested boxes with area ratios within a few ulps of the threshold,
identical boxes, degenerate (zero-area), huge and denormal boxes."""
import numpy as np
import pytest

f32 = np.float32


def iou32(a_lo, a_hi, b_lo, b_hi):
    """float32, operation by operation as oracle/nms_oracle.py:iou (the reference's utils/postprocess.py:21-36)."""
    with np.errstate(all="ignore"):
        a1 = ((a_hi[..., 0] - a_lo[..., 0]).astype(f32) * (a_hi[..., 1] - a_lo[..., 1]).astype(f32)).astype(f32)
        a2 = ((b_hi[..., 0] - b_lo[..., 0]).astype(f32) * (b_hi[..., 1] - b_lo[..., 1]).astype(f32)).astype(f32)
        wh = np.maximum((np.minimum(a_hi, b_hi) - np.maximum(a_lo, b_lo)).astype(f32), f32(0))
        inter = (wh[..., 0] * wh[..., 1]).astype(f32)
        den = np.maximum(((a1 + a2).astype(f32) - inter).astype(f32), f32(1e-10))
        return (inter / den).astype(f32), a1, a2


def window(area, thr):
    with np.errstate(all="ignore"):
        t = f32(f32(0.998) * f32(thr))
        return (t * area).astype(f32), (area / t).astype(f32)


def pairs(rs, n, thr):
    """[n] pairs (kept box, other box) concentrated where the bound is tight."""
    c = rs.uniform(-50, 50, size=(n, 2))
    wh_k = np.exp(rs.uniform(-6, 6, size=(n, 2)))
    kind = rs.randint(0, 6, size=n)
    # 0: other nested in kept, area ratio ~ thr;  1: kept nested in other;  2: identical;  3: random overlap;  4: zero-area other;
    # 5: extreme scales
    ratio = thr * (1.0 + rs.uniform(-3e-6, 3e-6, size=n))
    split = np.exp(rs.uniform(-0.3, 0.3, size=n))
    s = np.stack([np.sqrt(ratio) * split, np.sqrt(ratio) / split], -1)
    wh_o = np.where((kind == 0)[:, None], wh_k * np.minimum(s, 1.0), wh_k)
    wh_o = np.where((kind == 1)[:, None], wh_k / np.minimum(s, 1.0), wh_o)
    wh_o = np.where((kind == 3)[:, None], wh_k * np.exp(rs.uniform(-1.5, 1.5, size=(n, 2))), wh_o)
    wh_o = np.where((kind == 4)[:, None], wh_k * np.array([1.0, 0.0]), wh_o)
    scale = np.where(kind == 5, np.exp(rs.uniform(-45, 45, size=n)), 1.0)[:, None]
    off = np.where((kind == 3)[:, None], rs.uniform(-0.5, 0.5, size=(n, 2)) * wh_k, 0.0)
    k_lo, k_hi = (c - wh_k / 2) * scale, (c + wh_k / 2) * scale
    o_lo, o_hi = (c + off - wh_o / 2) * scale, (c + off + wh_o / 2) * scale
    return [x.astype(f32) for x in (k_lo, k_hi, o_lo, o_hi)]


@pytest.mark.parametrize("thr", [0.4, 0.5, 0.05, 0.9, 1.0, 1e-3])
def test_area_window_never_drops_a_hit(thr):
    rs = np.random.RandomState(int(thr * 1000) + 1)
    k_lo, k_hi, o_lo, o_hi = pairs(rs, 400000, thr)
    iou, ak, ao = iou32(k_lo, k_hi, o_lo, o_hi)
    hit = iou >= f32(thr)
    assert hit.sum() > 20000 and (~hit).sum() > 20000                  # the fuzz sits on both sides of the threshold
    lo, hi = window(ao, thr)                                           # a slot holding just this box: a_min = a_max = its area
    with np.errstate(all="ignore"):
        passes = (ak >= lo) & (ak <= hi)
    assert not (hit & ~passes).any(), "the window would drop a pair the reference suppresses"
    # and it is worth having: pairs whose areas differ by more than 1/thr (+ margin) are dropped
    far = (ak > ao * f32(1.01 / thr)) | (ak * f32(1.01 / thr) < ao)
    assert not (far & passes & np.isfinite(ak) & np.isfinite(ao) & (ao > 0)).any()


def test_restated_iou_is_the_oracles():
    from oracle.nms_oracle import pair_iou
    rs = np.random.RandomState(8)
    k_lo, k_hi, o_lo, o_hi = pairs(rs, 3000, 0.4)
    mine, _, _ = iou32(k_lo, k_hi, o_lo, o_hi)
    with np.errstate(all="ignore"):
        theirs = np.array([pair_iou(k_lo[i], k_hi[i], o_lo[i:i + 1], o_hi[i:i + 1])[0] for i in range(len(k_lo))], f32)
    assert np.array_equal(mine.view(np.uint32), theirs.view(np.uint32))


def test_a_hit_needs_two_positive_areas():
    """The culling also relies on: iou >= thr > 0 implies both areas > 0 (so slots of non-positive areas need no window)."""
    rs = np.random.RandomState(3)
    k_lo, k_hi, o_lo, o_hi = pairs(rs, 200000, 0.4)
    flip = rs.rand(len(k_lo)) < 0.3                                    # inverted / negative-area boxes
    o_lo2 = np.where(flip[:, None], o_hi, o_lo)
    o_hi2 = np.where(flip[:, None], o_lo, o_hi)
    iou, ak, ao = iou32(k_lo, k_hi, o_lo2, o_hi2)
    hit = iou >= f32(0.4)
    assert not (hit & ~((ak > 0) & (ao > 0))).any()
