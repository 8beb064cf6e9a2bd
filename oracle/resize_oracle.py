"""CPU restatement of the image resize on the detection path (TEST INFRASTRUCTURE ONLY).

detect.py:65 feeds the network `np.uint8(_image.resize((width, height)))`: PIL's `Image.resize` with its DEFAULT filter.
The algorithm therefore lives in a third-party dependency, Pillow (no version pinned by the reference; 2017's 4.x defaulted
to NEAREST, Pillow >= 7 -- 12.2.0 in this container -- defaults to BICUBIC).  As with NumPy for the NMS, the
oracle of record is the reference's call executed in THIS container, i.e. Pillow 12.2's bicubic; NEAREST is restated too.

Pillow's published algorithm (src/libImaging/Resample.c), 8 bits per channel:
  * per output coordinate xx: center = (xx + 0.5) * scale, support = 2 * max(scale, 1) (bicubic, a = -0.5, antialiased when
    shrinking), taps xmin = int(center - support + 0.5) clamped to 0, xmax = int(center + support + 0.5) clamped to the size;
    weights filter((x + xmin - center + 0.5) / max(scale, 1)) in double, normalised to sum 1;
  * weights to fixed point with 22 fractional bits: int(+-0.5 + w * 2^22), truncated towards zero;
  * each pass: acc = 2^21 + sum(pixel * weight) in int32, output = clip(acc >> 22, 0, 255); horizontal pass first into an
    8-bit intermediate, then vertical; a pass whose size does not change is skipped.
  * NEAREST (src/libImaging/Geometry.c, ImagingScaleAffine): source index = int(xo) with xo = scale / 2 for the first output
    coordinate and xo += scale (a running double sum, not a product) for each next one.
PINNED: tests/golden/resize.npz holds Pillow's own outputs for the sizes the tests use; here (where Pillow is importable) the
CPU tests also compare live over many size pairs.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
NEAREST, BICUBIC = 0, 3                      # PIL.Image.Resampling codes


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """-> (bounds int32 [out, 2] = (xmin, count), coeffs int32 [out, ksize]) exactly as precompute_coeffs + normalize_coeffs_8bpc"""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """img uint8 [H, W, C]; resample along `axis` (1 = horizontal, 0 = vertical)."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((len(bounds),) + src.shape[1:], np.int64)
    for i, (lo, n) in enumerate(bounds):
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[i, :n].astype(np.int64), src[lo:lo + n], axes=(0, 0))
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_oracle(img, out_w, out_h, resample=BICUBIC):
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit for bit what `Image.fromarray(img).resize((out_w, out_h), resample)` gives."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    if resample == NEAREST:
        def index(n_in, n_out):
            s = n_in / n_out
            idx, xo = np.empty(n_out, np.int64), s * 0.5
            for i in range(n_out):               # the running sum IS the algorithm: xo accumulates rounding, (i + 0.5) * s does not
                idx[i] = int(xo)
                xo += s
            assert idx.max() < n_in
            return idx
        return img[index(h, out_h)][:, index(w, out_w)]
    if resample != BICUBIC:
        raise ValueError("resize_oracle: NEAREST (0) or BICUBIC (3)")
    if w != out_w:
        img = _pass(img, *precompute_coeffs(w, out_w), axis=1)
    if h != out_h:
        img = _pass(img, *precompute_coeffs(h, out_h), axis=0)
    return img
