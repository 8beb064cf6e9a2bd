"""CPU restatement of the steps either side of the detection path (TEST INFRASTRUCTURE ONLY).

  per_image_standardization_oracle <- utils/preprocess.py:23-25.  PINNED: tests/golden/standardize.npz holds outputs of the
      reference's own function (tests/golden/make_preprocess_golden.py).
  detections_oracle                <- detect.py:72-87 (the loop after non_max_suppress): index = np.argmax(_conf) (first
      maximum), kept iff _conf[index] > threshold; _xy_min * scale, (_xy_max - _xy_min) * scale with
      scale = [image_width / cell_width, image_height / cell_height].  PINNED: tests/golden/detect_reference.npz holds what
      the reference's own detect() draws (the function compiled from detect.py as it lies, its own non_max_suppress, the TF
      session and matplotlib replaced by recorders: tests/golden/make_detect_golden.py); checked in tests/test_prepost.py.
"""
import numpy as np


def per_image_standardization_oracle(image):
    image = np.asarray(image)
    stddev = np.std(image)
    return (image - np.mean(image)) / max(stddev, 1.0 / np.sqrt(np.multiply.reduce(image.shape)))


def detections_oracle(conf, xy_min, xy_max, threshold, scale):
    """conf [N, C], xy_min / xy_max [N, 2] (one image, after NMS) -> list of (box index, class, score, xy_min_px, wh_px)
    in box-index order."""
    out = []
    scale = np.asarray(scale, dtype=np.float64)
    for n in range(conf.shape[0]):
        index = int(np.argmax(conf[n]))
        if conf[n][index] > threshold:
            wh = xy_max[n] - xy_min[n]
            out.append((n, index, conf[n][index], xy_min[n] * scale, wh * scale))
    return out
