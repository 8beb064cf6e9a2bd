"""Generates tests/golden/head_reference.npz by executing the REFERENCE'S OWN `Model` and `Objectives` classes
(/root/reference/model/yolo2/__init__.py:27-94) and `calc_cell_xy` (/root/reference/model/yolo/__init__.py:29-34).

TensorFlow 1.0 cannot be installed here, so the class bodies are compiled from the reference files as they lie (ast; nothing
is copied into this repo) and run against a stand-in for the 20 TF ops they call (`tf.reshape`, `tf.nn.sigmoid`, `tf.exp`,
`tf.nn.softmax`, `tf.reduce_prod/max/sum`, `tf.sqrt`, `tf.concat`, `tf.expand_dims`, `tf.maximum/minimum`, `tf.truediv`,
`tf.equal`, `tf.to_float`, `tf.square`, `tf.identity`, `tf.name_scope`, `Tensor.get_shape().as_list()`), each mapped to the
torch-CPU op of the same definition.  What this pins is everything the reference's SOURCE decides: the channel order of the
head, which slices feed which op, the IoU / best-box / mask composition, the normalisation `cnt`, the four objectives; and,
through torch autograd over that same source, d(total_loss)/d(net) -- the gradient the reference leaves to tf.gradients.
What it cannot pin is TF's own kernel arithmetic (summation order, exp / sigmoid ulps): the float64 run removes that
question from the comparison, the float32 run shows its size.

Cases (small grids: cells only interact through cnt): VOC (C = 20, 7 x 7, batch 2) and COCO (C = 80, 5 x 5) heads, labels from the reference's own transform_labels
(tests/golden/make_labels_golden.py), an image without objects, tied best boxes (identical anchors' logits), a non-square grid.
Run once, here:   python tests/golden/make_head_golden.py"""
import ast
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF2 = "/root/reference/model/yolo2/__init__.py"
REF1 = "/root/reference/model/yolo/__init__.py"
HPARAM = {"prob": 1.0, "iou_best": 5.0, "iou_normal": 1.0, "coords": 1.0}       # config.ini:98-102
ANCHORS_VOC = np.array([[1.3221, 1.73145], [3.19275, 4.00944], [5.05587, 8.09892], [9.47112, 4.84053], [11.2364, 10.0071]])
ANCHORS_COCO = np.array([[0.738768, 0.874946], [2.42204, 2.65704], [4.30971, 7.04493], [10.246, 4.59428], [12.6868, 11.8741]])


class T(object):
    """A 'tf.Tensor' holding a torch tensor.  Python / numpy operands are converted to the tensor's dtype, as TF does."""
    __array_ufunc__ = None              # numpy defers `ndarray <op> T` to T.__r<op>__
    __array_priority__ = 1000

    def __init__(self, v):
        self.v = v

    def _c(self, o):
        return o.v if isinstance(o, T) else torch.as_tensor(np.asarray(o), dtype=self.v.dtype)

    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self.v.shape))

    def __getitem__(self, k):
        return T(self.v[k])

    def __add__(self, o): return T(self.v + self._c(o))
    def __radd__(self, o): return T(self._c(o) + self.v)
    def __sub__(self, o): return T(self.v - self._c(o))
    def __rsub__(self, o): return T(self._c(o) - self.v)
    def __mul__(self, o): return T(self.v * self._c(o))
    def __rmul__(self, o): return T(self._c(o) * self.v)
    def __truediv__(self, o): return T(self.v / self._c(o))


def _t(x, like=None):
    if isinstance(x, T):
        return x.v
    return torch.as_tensor(np.asarray(x), dtype=like.dtype if like is not None else None)


def make_tf():
    tf = types.SimpleNamespace()
    tf.nn = types.SimpleNamespace()

    @contextlib.contextmanager
    def name_scope(name):
        yield name
    tf.name_scope = name_scope
    tf.reshape = lambda x, shape, name=None: T(_t(x).reshape(shape))
    tf.identity = lambda x, name=None: x if isinstance(x, T) else T(_t(x))
    tf.nn.sigmoid = lambda x, name=None: T(torch.sigmoid(_t(x)))
    tf.nn.softmax = lambda x, name=None: T(torch.softmax(_t(x), dim=-1))
    tf.exp = lambda x, name=None: T(torch.exp(_t(x)))
    tf.sqrt = lambda x, name=None: T(torch.sqrt(_t(x)))
    tf.square = lambda x, name=None: T(_t(x) * _t(x))
    tf.reduce_prod = lambda x, axis=None, name=None: T(torch.prod(_t(x), dim=axis))
    tf.reduce_max = lambda x, axis=None, keep_dims=False, name=None: T(torch.amax(_t(x), dim=axis, keepdim=keep_dims))
    tf.reduce_sum = lambda x, name=None: T(torch.sum(_t(x)))
    tf.concat = lambda xs, axis, name=None: T(torch.cat([_t(x) for x in xs], dim=axis))
    tf.expand_dims = lambda x, axis, name=None: T(_t(x).unsqueeze(axis))

    def _bin(fn):
        def f(a, b, name=None):
            ref = a.v if isinstance(a, T) else b.v
            return T(fn(_t(a, ref), _t(b, ref)))
        return f
    tf.maximum = _bin(torch.maximum)
    tf.minimum = _bin(torch.minimum)
    tf.truediv = _bin(torch.true_divide)
    tf.equal = lambda a, b, name=None: T(torch.eq(_t(a), _t(b)))
    tf.to_float = lambda x, name=None: T(_t(x).to(DTYPE[0]))          # "float" = the graph's float type in this run
    return tf


DTYPE = [torch.float32]


def load_reference_classes():
    tf = make_tf()
    tree1 = ast.parse(open(REF1).read())
    fn = [n for n in tree1.body if isinstance(n, ast.FunctionDef) and n.name == "calc_cell_xy"][0]
    ns1 = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF1, "exec"), ns1)
    tree2 = ast.parse(open(REF2).read())
    cls = [n for n in tree2.body if isinstance(n, ast.ClassDef) and n.name in ("Model", "Objectives")]
    assert len(cls) == 2
    ns2 = {"np": np, "tf": tf, "yolo": types.SimpleNamespace(calc_cell_xy=ns1["calc_cell_xy"])}
    exec(compile(ast.Module(body=cls, type_ignores=[]), REF2, "exec"), ns2)
    return ns2["Model"], ns2["Objectives"]


MODEL_ATTRS = ("iou", "offset_xy", "wh", "prob", "areas", "offset_xy_min", "offset_xy_max", "wh01", "wh01_sqrt", "coords",
               "xy", "xy_min", "xy_max", "conf")


def run_reference(net, classes, anchors, labels, dtype):
    """-> (dict of Model attributes (training=False), dict of the 4 objectives, d(sum_k hparam_k obj_k)/d(net))"""
    DTYPE[0] = dtype
    Model, Objectives = load_reference_classes()
    x = torch.tensor(net, dtype=dtype, requires_grad=True)
    model = Model(T(x), classes, anchors, training=False)
    attrs = {k: getattr(model, k).v.detach().numpy() for k in MODEL_ATTRS}
    obj = Objectives(model, *[T(torch.tensor(np.asarray(l), dtype=dtype)) for l in labels])
    total = sum(obj[k].v * HPARAM[k] for k in HPARAM)
    total.backward()
    return attrs, {k: obj[k].v.detach().numpy() for k in HPARAM}, x.grad.numpy()


def cases():
    from make_labels_golden import load_reference_transform_labels
    tl = load_reference_transform_labels()
    rs = np.random.RandomState(31)

    def boxes(n):
        cx, cy = rs.uniform(0, 1, n), rs.uniform(0, 1, n)
        w, h = rs.uniform(0.05, 0.6, n), rs.uniform(0.05, 0.6, n)
        return np.stack([np.clip(cx - w / 2, 0, 1 - 1e-6), np.clip(cy - h / 2, 0, 1 - 1e-6), np.clip(cx + w / 2, 0, 1 - 1e-6),
                         np.clip(cy + h / 2, 0, 1 - 1e-6)], 1).astype(np.float32)

    def labels(batch, classes, cw, ch, counts):
        per = [tl(rs.randint(0, classes, n), boxes(n), classes, cw, ch) for n in counts]
        assert len(per) == batch
        return tuple(np.stack([p[i] for p in per], 0) for i in range(6))

    out = {}
    # cells are independent of each other except through cnt = B * cells * A, so small grids pin the same arithmetic
    out["voc"] = (rs.normal(0, 1, size=(2, 7, 7, 5 * 25)).astype(np.float32), 20, ANCHORS_VOC, labels(2, 20, 7, 7, [4, 0]))
    out["coco"] = (rs.normal(0, 1.5, size=(1, 5, 5, 5 * 85)).astype(np.float32), 80, ANCHORS_COCO, labels(1, 80, 5, 5, [6]))
    out["nonsquare"] = (rs.normal(0, 1, size=(1, 3, 5, 5 * 25)).astype(np.float32), 20, ANCHORS_VOC, labels(1, 20, 5, 3, [3]))
    # ties: two anchors with identical logits AND identical anchor sizes give identical IoUs -> both are "best"
    anchors_tie = ANCHORS_VOC.copy()
    anchors_tie[1] = anchors_tie[0]
    net = rs.normal(0, 1, size=(1, 5, 5, 5, 25)).astype(np.float32)
    net[:, :, :, 1] = net[:, :, :, 0]
    out["ties"] = (net.reshape(1, 5, 5, 125), 20, anchors_tie, labels(1, 20, 5, 5, [8]))
    return out


def main():
    arrays = {"torch_version": np.array(torch.__version__), "numpy_version": np.array(np.__version__)}
    for name, (net, classes, anchors, lab) in cases().items():
        arrays[name + "_net"] = net
        arrays[name + "_anchors"] = anchors
        arrays[name + "_meta"] = np.array([classes])
        for n, l in zip(("mask", "prob", "coords", "offset_xy_min", "offset_xy_max", "areas"), lab):
            arrays["%s_label_%s" % (name, n)] = l
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            attrs, obj, grad = run_reference(net, classes, anchors, lab, dt)
            if tag == "f64":                 # the float32 run keeps objectives + gradient only (size of the arithmetic noise)
                for k, v in attrs.items():
                    arrays["%s_%s_model_%s" % (name, tag, k)] = v
            for k, v in obj.items():
                arrays["%s_%s_obj_%s" % (name, tag, k)] = v
            arrays["%s_%s_grad" % (name, tag)] = grad
        print(name, {k: float(v) for k, v in obj.items()}, "|grad| max %.3e" % np.abs(grad).max())
    np.savez_compressed(os.path.join(HERE, "head_reference.npz"), **arrays)
    print("wrote", os.path.join(HERE, "head_reference.npz"), "%.0f KiB" % (os.path.getsize(os.path.join(HERE, "head_reference.npz")) / 1024))


if __name__ == "__main__":
    main()
