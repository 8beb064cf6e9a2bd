"""Darknet `.weights` import -- SURVEY section 8(f) row 2 (the reference's parse_darknet_yolo2.py).

File format as the reference reads it (parse_darknet_yolo2.py:79-99): 16-byte header `4i` (major, minor, revision,
seen), then per layer conv0..conv20 and the final `conv`, float32 little-endian:
    BN layers : beta, gamma, moving_mean, moving_variance, weights        final layer: biases, weights
with weights stored [Cout, Cin, kh, kw] and transposed to TensorFlow's HWIO.  The final layer is then re-ordered per
anchor from Darknet's (x, y, w, h, iou, classes...) to this model's (iou, x, y, w, h, classes...) by
`transpose_weights` / `transpose_biases` (parse_darknet_yolo2.py:34-48, :101).

Where the reference assigns into TF variables and saves a checkpoint, this loader assigns into the variable store
(`yolo_tf_b200.variables`) under the same names (`yolo2_darknet/conv{i}/...`); the engine re-packs them for the tensor
cores on its next use.
"""
import os

import numpy as np

from . import variables as V
from .model.yolo2 import inference


def transpose_weights(weights, num_anchors):
    """parse_darknet_yolo2.py:34-40."""
    ksize1, ksize2, channels_in, _ = weights.shape
    weights = weights.reshape([ksize1, ksize2, channels_in, num_anchors, -1])
    return np.concatenate([weights[..., 4:5], weights[..., 0:4], weights[..., 5:]], -1).reshape([ksize1, ksize2, channels_in, -1])


def transpose_biases(biases, num_anchors):
    """parse_darknet_yolo2.py:43-48."""
    biases = biases.reshape([num_anchors, -1])
    return np.concatenate([biases[:, 4:5], biases[:, 0:4], biases[:, 5:]], -1).reshape([-1])


def read(path, classes, num_anchors, scope='yolo2_darknet', center=True):
    """-> (header dict, {variable name: float32 ndarray}) with TF variable names and layouts.
    center=False (the `_darknet` / `_tiny` graphs, inference.py:62-66,125-126): those graphs have no BatchNorm/beta but a
    `<conv>/biases` variable, and the reference's walk (`for suffix in ['biases', 'beta', 'gamma', ...]`,
    parse_darknet_yolo2.py:85-90) assigns the file's first per-layer block to it."""
    path = os.path.expanduser(os.path.expandvars(path))
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 16:
        raise ValueError("%s: not a Darknet weights file (%d bytes)" % (path, raw.size))
    major, minor, revision, seen = (int(v) for v in raw[:16].view('<i4'))
    data = raw[16:raw.size - ((raw.size - 16) % 4)].view('<f4')
    geo = inference.layer_geometry(classes, num_anchors)
    pos, values = 0, {}

    def take(cnt, what):
        nonlocal pos
        if pos + cnt > data.size:
            raise ValueError("%s: truncated at %s (%d floats needed, %d left)" % (path, what, cnt, data.size - pos))
        out = data[pos:pos + cnt]
        pos += cnt
        return out

    for name, k, cin, cout, has_bn, _ in geo:
        prefix = '%s/%s/' % (scope, name)
        if has_bn:
            for suffix in ('beta', 'gamma', 'moving_mean', 'moving_variance'):        # parse_darknet_yolo2.py:85 order
                name = prefix + ('biases' if (suffix == 'beta' and not center) else 'BatchNorm/' + suffix)
                values[name] = np.array(take(cout, name), dtype=np.float32)
        else:
            values[prefix + 'biases'] = np.array(take(cout, prefix + 'biases'), dtype=np.float32)
        w = take(cout * cin * k * k, prefix + 'weights').reshape([cout, cin, k, k])   # Darknet format
        values[prefix + 'weights'] = np.ascontiguousarray(np.transpose(w, [2, 3, 1, 0]), dtype=np.float32)   # HWIO
    last = '%s/%s/' % (scope, geo[-1][0])
    values[last + 'weights'] = np.ascontiguousarray(transpose_weights(values[last + 'weights'], num_anchors))
    values[last + 'biases'] = np.ascontiguousarray(transpose_biases(values[last + 'biases'], num_anchors))
    header = {'major': major, 'minor': minor, 'revision': revision, 'seen': seen, 'remaining': int((data.size - pos) * 4)}
    return header, values


def load(path, classes, num_anchors, scope='yolo2_darknet', store=None, center=True):
    """Read `path` and assign every variable into the store (the sess.run(v.assign(p)) loop of the reference).
    Pass center=False when the config selects `_darknet` / `_tiny` (see read())."""
    header, values = read(path, classes, num_anchors, scope, center)
    (store if store is not None else V.default_store()).assign(values)
    return header
