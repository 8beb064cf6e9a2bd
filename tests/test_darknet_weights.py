"""Darknet `.weights` import: the final-layer permutation is pinned bit-exact to the reference's own
functions (tests/golden/darknet_transpose.npz, made by tests/golden/make_darknet_golden.py); the file walk is checked
between two independent implementations (product: numpy views; oracle: struct loops) on synthetic files."""
import os

import numpy as np
import pytest

from oracle.darknet_oracle import init_params
from oracle.darknet_weights_oracle import (read_darknet_oracle, transpose_biases_oracle, transpose_weights_oracle,
                                           write_darknet_oracle)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "darknet_transpose.npz")


def test_transpose_matches_reference_golden():
    from yolo_tf_b200.parse_darknet_yolo2 import transpose_biases, transpose_weights
    g = np.load(GOLD)
    for tag in ("voc", "coco", "small"):
        a = int(g[tag + "_anchors"])
        for fw, fb in ((transpose_weights, transpose_biases), (transpose_weights_oracle, transpose_biases_oracle)):
            assert np.array_equal(fw(g[tag + "_w_in"], a), g[tag + "_w_out"]), tag
            assert np.array_equal(fb(g[tag + "_b_in"], a), g[tag + "_b_out"]), tag


@pytest.fixture(scope="module")
def weights_file(tmp_path_factory):
    """One synthetic 270 MB file (the network's size is fixed) shared by the CPU tests."""
    params = init_params(20, 5, seed=21)
    path = str(tmp_path_factory.mktemp("darknet") / "yolo.weights")
    write_darknet_oracle(path, params, 20, 5, header=(0, 1, 0, 32013312))
    return path, params


def test_read_synthetic_file_roundtrip(weights_file):
    from yolo_tf_b200.parse_darknet_yolo2 import read
    path, params = weights_file
    classes = 20
    header, values = read(path, classes, 5)
    oh, ovalues, oremaining = read_darknet_oracle(path, classes, 5)
    assert (header["major"], header["minor"], header["revision"], header["seen"]) == oh == (0, 1, 0, 32013312)
    assert header["remaining"] == oremaining == 0
    assert set(values) == {"yolo2_darknet/" + k for k in params}
    for k, v in params.items():
        got = values["yolo2_darknet/" + k]
        assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
        assert np.array_equal(got, v), k                      # the file round-trips bit for bit
        assert np.array_equal(ovalues[k], v), k               # and both readers agree


def test_read_center_false_assigns_the_first_block_to_biases(weights_file):
    """`_darknet` / `_tiny` graphs (inference.py:62-66): no BatchNorm/beta, a `<conv>/biases` variable instead, and the reference's
    walk order `['biases', 'beta', 'gamma', ...]` (parse_darknet_yolo2.py:85) hands it the file's first per-layer block."""
    from yolo_tf_b200.parse_darknet_yolo2 import read
    path, params = weights_file
    _, values = read(path, 20, 5, center=False)
    assert not any(k.endswith("BatchNorm/beta") for k in values)
    assert set(values) == {"yolo2_darknet/" + k.replace("BatchNorm/beta", "biases") for k in params}
    for k, v in params.items():
        assert np.array_equal(values["yolo2_darknet/" + k.replace("BatchNorm/beta", "biases")], v), k


def test_read_reports_trailing_bytes_and_truncation(weights_file):
    from yolo_tf_b200.parse_darknet_yolo2 import read
    path, _ = weights_file                                   # runs after the round-trip test (file order): may modify the file
    size = os.path.getsize(path)
    with open(path, "ab") as f:
        f.write(b"\0" * 12)
    assert read(path, 20, 5)[0]["remaining"] == 12
    with open(path, "r+b") as f:
        f.truncate(size - 4000)
    with pytest.raises(ValueError, match="truncated"):
        read(path, 20, 5)
    with open(path, "r+b") as f:
        f.truncate(8)
    with pytest.raises(ValueError, match="not a Darknet weights file"):
        read(path, 20, 5)


@pytest.mark.gpu
def test_load_into_store_and_run(cuda, tmp_path):
    """load() assigns the TF-named variables; the forward equals the one with the same parameters assigned directly."""
    import torch
    from yolo_tf_b200 import variables
    from yolo_tf_b200.model.yolo2 import inference
    from yolo_tf_b200.parse_darknet_yolo2 import load
    classes = 20
    params = init_params(classes, 5, seed=4)
    path = str(tmp_path / "w.weights")
    write_darknet_oracle(path, params, classes, 5)
    x = torch.from_numpy(np.random.RandomState(1).normal(0, 1, size=(2, 64, 64, 3)).astype(np.float32)).to(cuda)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    _, ref = inference.darknet(x, classes, 5)
    ref = ref.clone()
    store = variables.reset_default_store()
    header = load(path, classes, 5)
    assert header["remaining"] == 0
    _, out = inference.darknet(x, classes, 5)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


def test_file_walk_matches_the_reference_importer_run_end_to_end(tmp_path):
    """tests/golden/darknet_walk_reference.npz = what the reference's own `main()` (parse_darknet_yolo2.py:58-116) assigned to
    every variable when run on a synthetic `.weights` stream (graph built by its own darknet(), TF session / variables replaced by
    a holder of numpy values: tests/golden/make_darknet_walk_golden.py).  The stream is regenerated from its seed; the product
    reader and the oracle reader must give the same 107 variables (shape, sum, leading entries, random projection), the same
    header and the same count of left-over bytes."""
    import struct
    from yolo_tf_b200.parse_darknet_yolo2 import read
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "darknet_walk_reference.npz"))
    classes, anchors, seed, nfloats, extra = (int(v) for v in g["meta"])
    path = str(tmp_path / "synthetic.weights")
    rs = np.random.RandomState(seed)
    with open(path, "wb") as f:
        f.write(struct.pack("4i", 0, 1, 0, 32013312))
        left = nfloats + extra
        while left > 0:
            n = min(left, 1 << 22)
            f.write(rs.standard_normal(n).astype("<f4").tobytes())
            left -= n

    def summary(v):
        flat = np.asarray(v, dtype=np.float64).reshape(-1)
        probe = np.random.RandomState(flat.size % (2 ** 31)).normal(size=flat.size)
        return np.concatenate([[flat.sum()], flat[:8] if flat.size >= 8 else np.pad(flat, (0, 8 - flat.size)), [flat @ probe]])

    header, values = read(path, classes, anchors)
    oheader, ovalues, oremaining = read_darknet_oracle(path, classes, anchors)
    ovalues = {"yolo2_darknet/" + k: v for k, v in ovalues.items()}                      # the oracle names variables without the scope
    assert (header["major"], header["minor"], header["revision"], header["seen"]) == (0, 1, 0, 32013312)
    assert header["remaining"] == oremaining == 4 * extra
    assert "%d bytes remaining" % (4 * extra) in str(g["log_remaining"][0])
    names = [str(n) for n in g["names"]]
    assert set(names) == set(values) == set(ovalues) and len(names) == 107
    for n, shp, want in zip(names, g["shapes"], g["summary"]):
        shape = tuple(int(s) for s in str(shp).split("x"))
        for got in (values[n], ovalues[n]):
            assert tuple(got.shape) == shape and got.dtype == np.float32, n
            np.testing.assert_allclose(summary(got), want, rtol=1e-12, atol=1e-9, err_msg=n)
    assert np.array_equal(values["yolo2_darknet/conv/biases"], g["final_biases"])           # incl. the per-anchor re-ordering
    assert np.array_equal(values["yolo2_darknet/conv0/weights"], g["conv0_weights"])        # Darknet [O,I,kh,kw] -> HWIO
