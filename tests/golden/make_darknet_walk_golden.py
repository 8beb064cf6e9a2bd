"""Generates tests/golden/darknet_walk_reference.npz by executing the REFERENCE'S OWN `.weights` importer end to end:
`main()` of /root/reference/parse_darknet_yolo2.py:58-116 (with its `transpose*` helpers), compiled from the file as it lies.
It builds the graph with the reference's own `darknet()` (through the slim stand-in of make_backbone_golden.py, construction
only), collects `tf.global_variables()`, sorts the layers the way the reference does, walks the file with `struct.unpack` in
the reference's suffix order, transposes Darknet's [Cout, Cin, kh, kw] to HWIO, assigns, and finally re-orders the last layer.
TensorFlow's session / variables / saver are replaced by a 40-line stand-in that just holds numpy values.

The `.weights` file is a 16-byte header + a float32 stream, so a synthetic one is regenerated from a seed by the generator and
by the test (270 MB for Darknet-19 with 20 classes: not a fixture); the file holds, per variable the reference assigned, its
shape, sum, first 8 entries and a random projection, plus the header and the count of bytes left over.
Run once, here:   python tests/golden/make_darknet_walk_golden.py   (about a minute: the reference unpacks 67 M floats in Python)"""
import ast
import configparser
import contextlib
import itertools
import operator
import os
import re
import shutil
import struct
import sys
import tempfile
import types

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_backbone_golden as mb  # noqa: E402
import make_head_golden as mh  # noqa: E402

REF = "/root/reference/parse_darknet_yolo2.py"
CLASSES, SEED, EXTRA_FLOATS = 20, 77, 5          # 5 floats more than the graph consumes: the reference reports them as remaining
ANCHORS = [[1.08, 1.19], [3.42, 4.41], [6.63, 11.38], [9.42, 5.11], [16.62, 10.52]]


def synthetic_weights_file(path, nfloats, seed=SEED):
    """header (major 0, minor 1, revision 0, seen 32013312 -- what yolo-voc.weights carries) + nfloats float32 N(0, 1)"""
    rs = np.random.RandomState(seed)
    with open(path, "wb") as f:
        f.write(struct.pack("4i", 0, 1, 0, 32013312))
        left = nfloats
        while left > 0:
            n = min(left, 1 << 22)
            f.write(rs.standard_normal(n).astype("<f4").tobytes())
            left -= n


def summary(v):
    flat = np.asarray(v, dtype=np.float64).reshape(-1)
    probe = np.random.RandomState(flat.size % (2 ** 31)).normal(size=flat.size)
    return np.concatenate([[flat.sum()], flat[:8] if flat.size >= 8 else np.pad(flat, (0, 8 - flat.size)), [flat @ probe]])


class Variable(object):
    def __init__(self, name, shape):
        self.op = types.SimpleNamespace(name=name)
        self.shape = tuple(int(s) for s in shape)
        self.value = np.zeros(self.shape, np.float32)
        self.assigned = 0

    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self.shape))

    def assign(self, p):
        def op():
            p_ = np.asarray(p, dtype=np.float32)
            assert p_.size == self.value.size, (self.op.name, p_.shape, self.shape)
            self.value = p_.reshape(self.shape)
            self.assigned += 1
        return op


class CreateGraph(mb.Graph):
    """Graph construction only: every variable the reference's builder asks for is created (zeros), in creation order."""
    skip_compute = True

    def __init__(self):
        mb.Graph.__init__(self, {}, False)
        self.created = {}

    def var(self, name, shape):
        if name not in self.created:
            self.created[name] = Variable(name, shape)
        return torch.zeros(tuple(int(s) for s in shape), dtype=torch.float64)


def main():
    tmp = tempfile.mkdtemp(prefix="y2_walk_")
    try:
        g = CreateGraph()
        tf = mh.make_tf()
        slim = mb.make_slim(tf, g)
        log = []

        class Session(object):
            def __enter__(self): return self
            def __exit__(self, *a): return False
            def run(self, x):
                if isinstance(x, Variable):
                    return x.value
                return x()                          # an assign op
        tf.Session = Session
        tf.float32 = "float32"
        tf.placeholder = lambda dtype, shape, name=None: mh.T(torch.zeros(shape, dtype=torch.float64))
        tf.contrib = types.SimpleNamespace(framework=types.SimpleNamespace(get_or_create_global_step=lambda: None))
        tf.global_variables_initializer = lambda: types.SimpleNamespace(run=lambda: None)
        tf.global_variables = lambda: list(g.created.values())
        tf.logging = types.SimpleNamespace(info=lambda m: log.append(m), warn=lambda m: log.append("WARN " + m))
        tf.train = types.SimpleNamespace(Saver=lambda: types.SimpleNamespace(save=lambda sess, path: None))
        # the reference's graph builder, compiled from its file (as in make_backbone_golden.py)
        import inspect
        ns_l, ns_r = {"tf": tf}, {"tf": tf, "np": np}
        exec(compile(ast.Module(body=[n for n in ast.parse(open(mb.REF_FN1).read()).body if isinstance(n, ast.FunctionDef) and n.name == "leaky_relu"],
                                type_ignores=[]), mb.REF_FN1, "exec"), ns_l)
        exec(compile(ast.Module(body=[n for n in ast.parse(open(mb.REF_FN2).read()).body if isinstance(n, ast.FunctionDef) and n.name == "reorg"],
                                type_ignores=[]), mb.REF_FN2, "exec"), ns_r)
        ns_i = {"tf": tf, "slim": slim, "inspect": inspect, "leaky_relu": ns_l["leaky_relu"], "reorg": ns_r["reorg"], "__name__": "model.yolo2.inference"}
        exec(compile(ast.Module(body=[n for n in ast.parse(open(mb.REF_INF).read()).body if isinstance(n, ast.FunctionDef) and n.name == "darknet"],
                                type_ignores=[]), mb.REF_INF, "exec"), ns_i)
        # files the reference's main() reads
        with open(os.path.join(tmp, "names"), "w") as f:
            f.write("\n".join("c%d" % i for i in range(CLASSES)) + "\n")
        pd.DataFrame(ANCHORS, columns=["w", "h"]).to_csv(os.path.join(tmp, "anchors.tsv"), sep="\t", index=False)
        config = configparser.ConfigParser()
        config.read_dict({"config": {"model": "yolo2"}, "yolo2": {"inference": "darknet", "anchors": os.path.join(tmp, "anchors.tsv")}})
        nfloats = 0
        from oracle.darknet_oracle import layer_table
        for name, k, cin, cout, then in layer_table(CLASSES, len(ANCHORS)):
            nfloats += k * k * cin * cout + (cout if then == "linear" else 4 * cout)
        path = os.path.join(tmp, "synthetic.weights")
        synthetic_weights_file(path, nfloats + EXTRA_FLOATS)
        utils = types.SimpleNamespace(get_cachedir=lambda c: tmp, get_downsampling=lambda c: (32, 32), get_logdir=lambda c: os.path.join(tmp, "log"))
        ns = {"os": os, "re": re, "struct": struct, "itertools": itertools, "operator": operator, "np": np, "pd": pd, "shutil": shutil, "tf": tf,
              "utils": utils, "inference": types.SimpleNamespace(darknet=ns_i["darknet"]), "config": config,
              "args": types.SimpleNamespace(file=path, delete=False, summary=False, logname="x")}
        tree = ast.parse(open(REF).read())
        fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("transpose_weights", "transpose_biases", "transpose", "main")]
        assert len(fns) == 4
        exec(compile(ast.Module(body=fns, type_ignores=[]), REF, "exec"), ns)
        ns["main"]()
        names = list(g.created)
        assert all(v.assigned >= 1 for v in g.created.values()), [n for n, v in g.created.items() if not v.assigned]
        arrays = {"meta": np.array([CLASSES, len(ANCHORS), SEED, nfloats, EXTRA_FLOATS]),
                  "names": np.array(names), "shapes": np.array(["x".join(str(s) for s in g.created[n].shape) for n in names]),
                  "summary": np.stack([summary(g.created[n].value) for n in names]),
                  "log_remaining": np.array([m for m in log if "remaining" in m]),
                  "final_biases": g.created["yolo2_darknet/conv/biases"].value, "conv0_weights": g.created["yolo2_darknet/conv0/weights"].value}
        np.savez_compressed(os.path.join(HERE, "darknet_walk_reference.npz"), **arrays)
        print(len(names), "variables,", nfloats, "floats consumed;", [m for m in log if "remaining" in m or "major" in m])
        print("wrote darknet_walk_reference.npz %.0f KiB" % (os.path.getsize(os.path.join(HERE, "darknet_walk_reference.npz")) / 1024))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
