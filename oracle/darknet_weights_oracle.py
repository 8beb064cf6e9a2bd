"""CPU restatement of the reference's Darknet `.weights` import (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/parse_darknet_yolo2.py:
  :34-48   transpose_weights / transpose_biases: final layer, per anchor (x, y, w, h, iou, classes...) -> (iou, x, y, w, h, classes...)
  :79-80   header = struct '4i' (major, minor, revision, seen)
  :81-99   per layer conv0..conv20, then the final `conv`: for suffix in [biases, beta, gamma, moving_mean, moving_variance,
           weights] (those the layer has): read prod(shape) float32; weights are stored [Cout, Cin, kh, kw] (Darknet) and
           transposed [2, 3, 1, 0] to HWIO
  :101     transpose() on the LAST layer only
Pinned: transpose_* against tests/golden/darknet_transpose.npz (made by the reference's own functions,
tests/golden/make_darknet_golden.py).  The file walk itself is restated (it needs a TF session in the reference).
Deliberately written as a sequential struct / array.fromfile walk, unlike the product reader (one numpy view of the
whole file), so the two are independent implementations."""
import array
import struct

import numpy as np

from oracle.darknet_oracle import layer_table


def transpose_weights_oracle(weights, num_anchors):
    k1, k2, cin, _ = weights.shape
    w = weights.reshape([k1, k2, cin, num_anchors, -1])
    return np.concatenate([w[..., 4:5], w[..., 0:4], w[..., 5:]], -1).reshape([k1, k2, cin, -1])


def transpose_biases_oracle(biases, num_anchors):
    b = biases.reshape([num_anchors, -1])
    return np.concatenate([b[:, 4:5], b[:, 0:4], b[:, 5:]], -1).reshape([-1])


def read_darknet_oracle(path, classes, num_anchors):
    """-> (header tuple, {variable suffix name: float32 array})  e.g. 'conv0/BatchNorm/gamma', 'conv/weights'."""
    out = {}
    table = layer_table(classes, num_anchors)
    with open(path, "rb") as f:
        header = struct.unpack("4i", f.read(16))
        for li, (name, k, cin, cout, then) in enumerate(table):
            last = li == len(table) - 1
            fields = [("biases", (cout,))] if last else [("BatchNorm/beta", (cout,)), ("BatchNorm/gamma", (cout,)),
                                                         ("BatchNorm/moving_mean", (cout,)), ("BatchNorm/moving_variance", (cout,))]
            fields.append(("weights", (k, k, cin, cout)))
            for suffix, shape in fields:
                cnt = int(np.prod(shape))
                buf = array.array("f")                       # native float32 (little-endian hosts), like struct '%df'
                buf.fromfile(f, cnt)
                p = np.array(buf, dtype=np.float32)
                if suffix == "weights":
                    p = np.transpose(p.reshape([cout, cin, k, k]), [2, 3, 1, 0])
                out[name + "/" + suffix] = np.ascontiguousarray(p.reshape(shape))
        remaining = len(f.read())
    out["conv/weights"] = transpose_weights_oracle(out["conv/weights"], num_anchors)
    out["conv/biases"] = transpose_biases_oracle(out["conv/biases"], num_anchors)
    return header, out, remaining


def write_darknet_oracle(path, params, classes, num_anchors, header=(0, 1, 0, 32013312)):
    """Inverse of read_darknet_oracle (for synthetic fixtures): params keyed like darknet_oracle.init_params."""
    table = layer_table(classes, num_anchors)
    with open(path, "wb") as f:
        f.write(struct.pack("4i", *header))
        for li, (name, k, cin, cout, then) in enumerate(table):
            last = li == len(table) - 1
            w = np.asarray(params[name + "/weights"], dtype=np.float32)
            if last:
                b = np.asarray(params[name + "/biases"], dtype=np.float32)
                # undo the (iou, coords, classes) order: back to Darknet's (coords, iou, classes)
                wr = w.reshape([k, k, cin, num_anchors, -1])
                w = np.concatenate([wr[..., 1:5], wr[..., 0:1], wr[..., 5:]], -1).reshape([k, k, cin, -1])
                br = b.reshape([num_anchors, -1])
                b = np.concatenate([br[:, 1:5], br[:, 0:1], br[:, 5:]], -1).reshape([-1])
                f.write(b.astype("<f4").tobytes())
            else:
                for suffix in ("beta", "gamma", "moving_mean", "moving_variance"):
                    f.write(np.asarray(params[name + "/BatchNorm/" + suffix], dtype="<f4").tobytes())
            f.write(np.ascontiguousarray(np.transpose(w, [3, 2, 0, 1])).astype("<f4").tobytes())
