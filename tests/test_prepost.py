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
