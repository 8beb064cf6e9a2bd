"""Generates tests/golden/backbone_reference.npz by executing the REFERENCE'S OWN graph builders `darknet()`, `_darknet()`,
`tiny()` (/root/reference/model/yolo2/inference.py:25-126), its `reorg()` (model/yolo2/function.py:22-29) and its
`leaky_relu()` (model/yolo/function.py:21-24), compiled from the reference files as they lie (ast; nothing copied).

TensorFlow 1.0 / slim cannot be installed here, so the functions run against a stand-in for the slim / tf calls they make
(`slim.arg_scope`, `slim.layers.conv2d`, `slim.batch_norm`, `slim.layers.max_pool2d`, `slim.variable`, `tf.concat`,
`tf.reshape`, `tf.transpose`, `tf.nn.bias_add`, ...), each mapped to the torch-CPU float64 op of the same definition:
conv2d = stride-1 SAME cross-correlation with HWIO weights and no bias when a normalizer is given; batch_norm =
tf.nn.batch_normalization with moving statistics (inference) or batch mean / population variance (training), eps as passed;
max_pool2d SAME with the stride the reference passes.  Variables are looked up BY THE NAME slim would give them
(`<scope>/weights`, `<scope>/BatchNorm/gamma`, ..., `<scope>/biases`) in a synthetic checkpoint, and every lookup is recorded.

What this pins is everything the reference's SOURCE decides: the layer sequence, kernel sizes and channel counts (including
the Python-3 `channels / 2` floats), where the pools sit, the passthrough tap, reorg's element order, the concat order, the
variable names and shapes a checkpoint must carry, the `center=False` variant's separate `biases`.  Not TF's own arithmetic.
The checkpoint is regenerated from its seed by the tests (67 M parameters do not belong in a fixture); the file holds the
outputs, a few intermediate taps, and the recorded variable table.
Run once, here:   python tests/golden/make_backbone_golden.py"""
import ast
import contextlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_head_golden import T, _t, make_tf  # noqa: E402

REF_INF = "/root/reference/model/yolo2/inference.py"
REF_FN2 = "/root/reference/model/yolo2/function.py"
REF_FN1 = "/root/reference/model/yolo/function.py"


class Graph(object):
    """Variable lookup + the bits of graph state slim keeps (current variable scope, arg_scope defaults)."""

    def __init__(self, params, training):
        self.params, self.training = params, training
        self.scope = None
        self.defaults = []                # stack of (function names, kwargs)
        self.variables = []               # (name, shape) in creation order
        self.taps = {}

    def var(self, name, shape):
        shape = tuple(int(s) for s in shape)
        assert name in self.params, "the reference graph asks for variable %s which the checkpoint lacks" % name
        v = self.params[name]
        assert tuple(v.shape) == shape, "variable %s: graph wants %s, checkpoint has %s" % (name, shape, tuple(v.shape))
        self.variables.append((name, shape))
        return torch.as_tensor(v, dtype=torch.float64)


def make_slim(tf, g):
    slim = types.SimpleNamespace()
    slim.layers = types.SimpleNamespace()

    @contextlib.contextmanager
    def arg_scope(funcs, **kw):
        g.defaults.append((tuple(f.__name__ for f in funcs), kw))
        try:
            yield
        finally:
            g.defaults.pop()
    slim.arg_scope = arg_scope

    def with_defaults(name, kw):
        out = {}
        for names, d in g.defaults:
            if name in names:
                out.update(d)
        out.update(kw)
        return out

    def relu(x, name=None):
        return T(torch.relu(_t(x)))

    def conv2d(inputs, num_outputs, **kw):
        kw = with_defaults("conv2d", kw)
        k = kw.get("kernel_size", None)
        kh, kw_ = (k, k) if isinstance(k, int) else k
        scope = kw["scope"]
        x = _t(inputs)
        cin = x.shape[-1]
        assert float(num_outputs) == int(num_outputs)             # `channels / 2` is a float under Python 3 (inference.py:80)
        w = g.var(scope + "/weights", (kh, kw_, cin, int(num_outputs)))
        if getattr(g, "skip_compute", False):                     # graph construction only (make_darknet_walk_golden.py)
            y = torch.zeros(tuple(x.shape[:3]) + (int(num_outputs),), dtype=torch.float64)
        else:
            y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=(kh // 2, kw_ // 2)).permute(0, 2, 3, 1)
        norm = kw.get("normalizer_fn", None)
        prev, g.scope = g.scope, scope
        if norm is not None:
            y = _t(norm(T(y)))                                     # slim: no biases when a normalizer_fn is given
        else:
            y = y + g.var(scope + "/biases", (int(num_outputs),))
        g.scope = prev
        act = kw.get("activation_fn", relu)                       # slim's default activation is relu; None = linear
        out = act(T(y)) if act is not None else T(y)
        g.taps[scope] = _t(out).detach().numpy()
        return out
    slim.layers.conv2d = conv2d

    def batch_norm(inputs, center=True, scale=False, epsilon=0.001, is_training=True, **_):
        x = _t(inputs)
        c = x.shape[-1]
        s = g.scope + "/BatchNorm"
        beta = g.var(s + "/beta", (c,)) if center else torch.zeros(c, dtype=torch.float64)
        gamma = g.var(s + "/gamma", (c,)) if scale else torch.ones(c, dtype=torch.float64)
        mm, mv = g.var(s + "/moving_mean", (c,)), g.var(s + "/moving_variance", (c,))
        if is_training:
            mm, mv = x.mean(dim=(0, 1, 2)), x.var(dim=(0, 1, 2), unbiased=False)
        inv = torch.rsqrt(mv + epsilon) * gamma
        return T(x * inv + (beta - mm * inv))
    slim.batch_norm = batch_norm

    def max_pool2d(inputs, **kw):
        kw = with_defaults("max_pool2d", kw)
        assert list(kw["kernel_size"]) == [2, 2] and kw["padding"] == "SAME"
        stride = kw.get("stride", 2)                              # slim's default
        x = _t(inputs).permute(0, 3, 1, 2)
        if stride == 1:                                           # SAME, k = 2, s = 1: one pad row / column AFTER, ignored by max
            y = F.max_pool2d(F.pad(x, (0, 1, 0, 1), value=float("-inf")), 2, 1)
        else:
            assert stride == 2 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
            y = F.max_pool2d(x, 2, 2)
        return T(y.permute(0, 2, 3, 1))
    slim.layers.max_pool2d = max_pool2d
    slim.variable = lambda name, shape, initializer=None: T(g.var(g.scope + "/" + name, shape))

    tf.shape = lambda x: list(_t(x).shape)
    tf.nn.bias_add = lambda x, b, name=None: T(_t(x) + _t(b))
    tf.zeros_initializer = lambda: None
    tf.truncated_normal_initializer = lambda stddev=1.0: None
    tf.transpose = lambda x, perm, name=None: T(_t(x).permute(*perm))
    return slim


def load_reference(params, training):
    g = Graph(params, training)
    tf = make_tf()
    slim = make_slim(tf, g)
    ns_l = {"tf": tf}
    tree = ast.parse(open(REF_FN1).read())
    exec(compile(ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "leaky_relu"], type_ignores=[]),
                 REF_FN1, "exec"), ns_l)
    ns_r = {"tf": tf, "np": np}
    tree = ast.parse(open(REF_FN2).read())
    exec(compile(ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "reorg"], type_ignores=[]),
                 REF_FN2, "exec"), ns_r)
    import inspect
    ns = {"tf": tf, "slim": slim, "inspect": inspect, "leaky_relu": ns_l["leaky_relu"], "reorg": ns_r["reorg"],
          "__name__": "model.yolo2.inference"}
    tree = ast.parse(open(REF_INF).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("darknet", "_darknet", "tiny", "_tiny")]
    assert len(fns) == 4
    exec(compile(ast.Module(body=fns, type_ignores=[]), REF_INF, "exec"), ns)
    return g, ns


def checkpoint(func, classes, anchors, center):
    """The oracle's synthetic checkpoint (regenerated from its seed by the tests), under the names the reference graph uses."""
    from oracle.darknet_oracle import init_params, tiny_layer_table
    table = tiny_layer_table(classes, anchors) if "tiny" in func else None
    p = init_params(classes, anchors, seed=1, table=table)
    scope = "yolo2_" + func.lstrip("_")
    out = {}
    for k, v in p.items():
        if not center and k.endswith("/BatchNorm/beta"):
            k = k[:-len("/BatchNorm/beta")] + "/biases"          # center=False: BN has no beta, a separate `biases` follows it
        out[scope + "/" + k] = v
    return out


def run(func, x, classes, anchors, training):
    center = not func.startswith("_")
    g, ns = load_reference(checkpoint(func, classes, anchors, center), training)
    scope, net = ns[func](T(torch.as_tensor(x, dtype=torch.float64)), classes, anchors, training)
    return scope, _t(net).detach().numpy(), g


def main():
    rs = np.random.RandomState(17)
    arrays = {"torch_version": np.array(torch.__version__)}
    x64 = rs.normal(0, 1, size=(3, 64, 64, 3)).astype(np.float32)       # the geometry of the GPU layer-by-layer test (20, 64, 3)
    x96 = rs.normal(0, 1, size=(2, 96, 64, 3)).astype(np.float32)          # batch 2, non-square (3 x 2 cells)
    arrays["x64"], arrays["x96"] = x64, x96
    for tag, func, x, classes, training in (("darknet", "darknet", x64, 20, False), ("darknet_rect", "darknet", x96, 20, False),
                                            ("darknet_train", "darknet", x96, 20, True), ("darknet_nocenter", "_darknet", x64, 20, False),
                                            ("tiny", "tiny", x96, 20, False), ("tiny_nocenter", "_tiny", x64, 20, False)):
        scope, out, g = run(func, x, classes, 5, training)
        arrays[tag + "_out"] = out
        arrays[tag + "_scope"] = np.array(scope)
        arrays[tag + "_vars"] = np.array(["%s %s" % (n, "x".join(str(s) for s in shp)) for n, shp in g.variables])
        for name in ("conv0", "conv12", "conv19", "conv20") if "darknet" in tag else ("conv0", "conv5", "conv7"):
            t = g.taps["%s/%s" % (scope, name)]
            arrays["%s_tap_%s" % (tag, name)] = t[:, :4, :4, :16].copy()    # a corner of each tap is enough to pin it
        print(tag, scope, out.shape, "%d variables" % len(g.variables), "|out| max %.3f" % np.abs(out).max())
    np.savez_compressed(os.path.join(HERE, "backbone_reference.npz"), **arrays)
    print("wrote backbone_reference.npz %.0f KiB" % (os.path.getsize(os.path.join(HERE, "backbone_reference.npz")) / 1024))


if __name__ == "__main__":
    main()
