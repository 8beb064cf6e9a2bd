"""Pre / post steps: per_image_standardization pinned to the reference's own outputs
(tests/golden/standardize.npz); the detection selection loop against its numpy restatement."""
import os

import numpy as np
import pytest

from oracle.prepost_oracle import detections_oracle, per_image_standardization_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "standardize.npz")
CASES = ("rand32", "rand64x48", "dark", "constant", "one_hot")


def test_standardization_oracle_matches_reference_golden():
    g = np.load(GOLD)
    for name in CASES:
        got = per_image_standardization_oracle(g[name + "_u8"].astype(np.float32))
        assert np.array_equal(np.asarray(got), g[name + "_out"]), name          # same numpy, same expression: bit for bit


@pytest.mark.gpu
@pytest.mark.parametrize("as_uint8", [True, False])
def test_standardization_gpu_vs_reference_golden(cuda, as_uint8):
    import torch
    from yolo_tf_b200.utils.preprocess import per_image_standardization
    g = np.load(GOLD)
    for name in CASES:
        u8 = g[name + "_u8"]
        x = torch.from_numpy(u8 if as_uint8 else u8.astype(np.float32)).to(cuda)
        got = per_image_standardization(x).cpu().numpy()
        ref = g[name + "_out"].astype(np.float64)
        assert got.dtype == np.float32 and got.shape == ref.shape
        # reduction order differs from numpy's pairwise float32 sums: 1e-5 of the output range
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1.0), name


@pytest.mark.gpu
def test_standardization_batched_matches_per_image(cuda):
    import torch
    from yolo_tf_b200.utils.preprocess import per_image_standardization
    rs = np.random.RandomState(3)
    batch = rs.randint(0, 256, size=(5, 416, 416, 3)).astype(np.uint8)
    batch[3] = 9                                                   # a constant image: denominator 1/sqrt(n)
    got = per_image_standardization(torch.from_numpy(batch).to(cuda)).cpu().numpy()
    for b in range(batch.shape[0]):
        ref = np.asarray(per_image_standardization_oracle(batch[b].astype(np.float32)), dtype=np.float64)
        assert np.abs(got[b] - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1.0), b
        one = per_image_standardization(torch.from_numpy(batch[b]).to(cuda)).cpu().numpy()
        assert np.array_equal(one, got[b])                         # batched == one at a time, bit for bit (deterministic)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 37, 29, 3), (2, 416, 416, 3), (4, 20, 12, 3)])
def test_standardization_uint8_paths_agree_with_the_oracle(cuda, shape):
    """uint8 images take a one-pass path (exact integer sums, 16 pixels per load) when an image is a multiple of 16 bytes, the
    generic two-pass path otherwise; both against the numpy restatement, and against each other through a float32 copy."""
    import torch
    from yolo_tf_b200.utils.preprocess import per_image_standardization
    rs = np.random.RandomState(sum(shape))
    batch = rs.randint(0, 256, size=shape).astype(np.uint8)
    got = per_image_standardization(torch.from_numpy(batch).to(cuda)).cpu().numpy()
    via_f32 = per_image_standardization(torch.from_numpy(batch.astype(np.float32)).to(cuda)).cpu().numpy()
    for b in range(shape[0]):
        ref = np.asarray(per_image_standardization_oracle(batch[b].astype(np.float32)), dtype=np.float64)
        assert np.abs(got[b] - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1.0), b
        assert np.abs(got[b] - via_f32[b]).max() <= 1e-5 * max(np.abs(ref).max(), 1.0), b


@pytest.mark.gpu
@pytest.mark.parametrize("n,c", [(845, 80), (1805, 20), (7, 3)])
def test_detections_gpu_vs_oracle(cuda, n, c):
    import torch
    from yolo_tf_b200.utils.postprocess import detections_device
    rs = np.random.RandomState(n + c)
    B = 3
    conf = rs.uniform(0, 0.29, size=(B, n, c)).astype(np.float32)
    pick = rs.rand(B, n, c) < 0.01
    conf[pick] = rs.uniform(0.3, 1.0, size=int(pick.sum())).astype(np.float32)
    conf[0, 1, :] = 0.5                                            # a tie over all classes: argmax takes class 0
    conf[1, 2, c - 1] = 0.3                                        # exactly on the threshold: not kept (strict >)
    conf[1, 2, :c - 1] = 0.0
    lo = rs.uniform(0, 12, size=(B, n, 2)).astype(np.float32)
    hi = lo + rs.uniform(0.1, 5, size=(B, n, 2)).astype(np.float32)
    scale = [640 / 13.0, 480 / 13.0]
    count, box, cls, score, xywh = [t.cpu().numpy() for t in detections_device(
        torch.from_numpy(conf).to(cuda), torch.from_numpy(lo).to(cuda), torch.from_numpy(hi).to(cuda), 0.3, scale)]
    for b in range(B):
        ref = detections_oracle(conf[b], lo[b], hi[b], np.float32(0.3), scale)
        assert count[b] == len(ref)
        for i, (rn, rc, rscore, rxy, rwh) in enumerate(ref):
            assert box[b, i] == rn and cls[b, i] == rc and score[b, i] == rscore
            np.testing.assert_allclose(xywh[b, i, :2], rxy, rtol=2e-7)
            np.testing.assert_allclose(xywh[b, i, 2:], rwh, rtol=2e-7)
    assert cls[0, list(box[0, :count[0]]).index(1)] == 0 if 1 in box[0, :count[0]] else True


DETECT_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "detect_reference.npz")
DETECT_CASES = ("voc13", "coco7", "rect", "none")


def _reference_drawn(g, name):
    """What the reference's own detect() drew (tests/golden/make_detect_golden.py): rows (x, y, w, h, linewidth) in pixels,
    class per row; sorted so that the reference's NMS-list order and the oracle's box-index order compare as SETS."""
    rect, cls = g[name + "_rect"], g[name + "_cls"]
    rows = sorted((float(r[0]), float(r[1]), float(r[2]), float(r[3]), int(c)) for r, c in zip(rect, cls))
    return rows


def _assert_same_boxes(mine, want, atol):
    """Both lists hold (x, y, w, h, class); equal as multisets within atol pixels (order-free, robust to near-equal keys)."""
    assert len(mine) == len(want)
    left = list(mine)
    for w in want:
        hit = [i for i, m in enumerate(left) if m[4] == w[4] and max(abs(m[j] - w[j]) for j in range(4)) <= atol + 1e-6 * abs(w[0])]
        assert hit, ("the reference draws %s, not found" % (w,))
        left.pop(hit[0])


@pytest.mark.parametrize("name", DETECT_CASES)
def test_detections_oracle_matches_what_the_reference_detect_draws(name):
    """detect.py:56-88 executed as it lies (its own non_max_suppress; session / matplotlib replaced by recorders): the NMS
    oracle reproduces the in-place zeroing, and detections_oracle on the result selects the boxes the reference draws --
    same class (argmax, first maximum on exact ties), same pixel rectangle (scale = original image size / cells), same count."""
    from oracle.nms_oracle import nms_oracle
    g = np.load(DETECT_GOLD)
    cw, ch, iw, ih = (int(v) for v in g[name + "_meta"])
    conf = g[name + "_conf_in"].copy()
    nms_oracle(conf, g[name + "_xy_min"], g[name + "_xy_max"], 0.3, 0.4)
    assert np.array_equal(conf.view(np.uint32), g[name + "_conf_out"].view(np.uint32))
    n, c = conf.shape[0] * conf.shape[1], conf.shape[2]
    det = detections_oracle(conf.reshape(n, c), g[name + "_xy_min"].reshape(n, 2), g[name + "_xy_max"].reshape(n, 2), 0.3,
                            [iw / cw, ih / ch])
    mine = sorted((float(xy[0]), float(xy[1]), float(wh[0]), float(wh[1]), int(k)) for _, k, _, xy, wh in det)
    want = _reference_drawn(g, name)
    _assert_same_boxes(mine, want, 1e-4)
    # the label text carries the score: '%.1f%%' of conf[index] * 100
    scores = sorted("%.1f" % (s * 100) for _, _, s, _, _ in det)
    assert scores == sorted(str(p) for p in g[name + "_pct"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", DETECT_CASES)
def test_nms_and_detections_gpu_vs_what_the_reference_detect_draws(cuda, name):
    import torch
    from yolo_tf_b200.utils.postprocess import detections_device, non_max_suppress_device
    g = np.load(DETECT_GOLD)
    cw, ch, iw, ih = (int(v) for v in g[name + "_meta"])
    shp = g[name + "_conf_in"].shape
    n, c = shp[0] * shp[1], shp[2]
    conf = torch.from_numpy(g[name + "_conf_in"].reshape(1, n, c).copy()).to(cuda)
    lo = torch.from_numpy(g[name + "_xy_min"].reshape(1, n, 2).copy()).to(cuda)
    hi = torch.from_numpy(g[name + "_xy_max"].reshape(1, n, 2).copy()).to(cuda)
    non_max_suppress_device(conf, lo, hi, 0.3, 0.4)
    assert np.array_equal(conf.cpu().numpy().reshape(shp).view(np.uint32), g[name + "_conf_out"].view(np.uint32))
    count, box, cls, score, xywh = detections_device(conf, lo, hi, 0.3, [iw / cw, ih / ch])
    k = int(count[0])
    mine = sorted((float(r[0]), float(r[1]), float(r[2]), float(r[3]), int(q)) for r, q in zip(xywh[0, :k].cpu().numpy(), cls[0, :k].cpu().numpy()))
    want = _reference_drawn(g, name)
    _assert_same_boxes(mine, want, 1e-3)            # float32 products on the device vs float64 in the reference's numpy
