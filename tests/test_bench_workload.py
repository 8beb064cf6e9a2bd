"""CPU (-m "not gpu"): the synthetic workload bench.py times is what its JSON line says it is.  Round 1's checkpoint produced NO
score above the 0.3 detection threshold, so its "inference + NMS" headline timed an NMS that did nothing; the final
layer is now scaled so that a busy scene's worth of candidates reaches the NMS.  Checked with the CPU oracle on one image."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_default_workload_feeds_the_nms_about_900_candidates_per_image():
    import torch
    import bench
    from oracle.darknet_oracle import darknet_oracle
    from oracle.head_oracle import decode_oracle
    from oracle.nms_oracle import nms_oracle
    from oracle.prepost_oracle import per_image_standardization_oracle
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    params = bench.synthetic_checkpoint(80, 5)
    u8 = bench.synthetic_images_u8(np.random.RandomState(100), 1, 416)
    x = np.stack([per_image_standardization_oracle(u8[0].astype(np.float32))]).astype(np.float32)
    conf = decode_oracle(darknet_oracle(x, params, 80, 5), 80, bench.ANCHORS_COCO)
    scores = np.ascontiguousarray(conf["conf"][0])
    cands = int((scores > bench.THRESHOLD).sum())
    assert 600 <= cands <= 1300, cands                                   # measured on the B200: 926 per image (batch mean)
    per_class = (scores.reshape(-1, 80) > bench.THRESHOLD).sum(0)
    assert per_class.max() >= 64                                         # skewed like real score matrices: a few heavy classes
    nms_oracle(scores, np.ascontiguousarray(conf["xy_min"][0]), np.ascontiguousarray(conf["xy_max"][0]), bench.THRESHOLD, bench.THRESHOLD_IOU)
    kept = int((scores.reshape(-1, 80).max(1) > bench.THRESHOLD).sum())
    assert 100 <= kept < cands                                           # the NMS suppresses something and keeps something
    # the round-1 checkpoint (no scaling): not a single candidate
    plain = bench.synthetic_checkpoint(80, 5, dense_detections=False)
    c0 = decode_oracle(darknet_oracle(x, plain, 80, 5), 80, bench.ANCHORS_COCO)["conf"]
    assert int((c0 > bench.THRESHOLD).sum()) == 0


def test_bench_flop_accounting_matches_the_survey():
    import bench
    assert abs(bench.conv_flops(416, 416, 80, 5) / 1e9 - 35.002) < 0.01          # 35.002 GFLOP / image (416, C80)
    assert abs(bench.conv_flops(608, 608, 80, 5) / 1e9 - 74.768) < 0.01
    assert abs(bench.train_flops_per_image(416, 20) / 1e9 - 104.39) < 0.01       # fwd + dgrad + wgrad, no dgrad for conv0
    burst, which = bench.pick_peak({"bf16_tflops": 1633.1, "bf16_tflops_sustained": 1382.5}, 0.7)
    assert burst == 1633.1 and "burst" in which
    assert bench.pick_peak({"bf16_tflops": 1633.1, "bf16_tflops_sustained": 1382.5}, 5.0)[0] == 1382.5
