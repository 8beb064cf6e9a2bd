"""Generates tests/golden/labels.npz with the REFERENCE'S OWN transform_labels
(/root/reference/utils/data/__init__.py:112-145).  The module imports tensorflow at the top and the function uses the
removed alias `np.int` (:129), so the function's source lines are compiled from the reference file as they lie (ast,
nothing copied into this repo) with `np.int` mapped to the builtin it aliased.  Inputs are float32 coordinates and Python
ints for the grid, which is what train.py feeds (tf.float32 tensors through tf.py_func), so every intermediate is float32
under both the reference-era and the current NumPy promotion rules.  Run once, here:
    python tests/golden/make_labels_golden.py"""
import ast
import os

import numpy as np

REF = "/root/reference/utils/data/__init__.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_transform_labels():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "transform_labels"][0]
    mod = ast.Module(body=[fn], type_ignores=[])

    class NP(object):                        # numpy with the alias the reference still uses
        int = int

        def __getattr__(self, k):
            return getattr(np, k)
    ns = {"np": NP()}
    exec(compile(mod, REF, "exec"), ns)
    return ns["transform_labels"]


def cases():
    rs = np.random.RandomState(23)
    out = {}

    def boxes(n, lo=0.05, hi=0.6):
        cx, cy = rs.uniform(0, 1, n), rs.uniform(0, 1, n)
        w, h = rs.uniform(lo, hi, n), rs.uniform(lo, hi, n)
        c = np.stack([np.clip(cx - w / 2, 0, 1 - 1e-6), np.clip(cy - h / 2, 0, 1 - 1e-6), np.clip(cx + w / 2, 0, 1 - 1e-6),
                      np.clip(cy + h / 2, 0, 1 - 1e-6)], 1)
        return c.astype(np.float32)
    out["voc13"] = (rs.randint(0, 20, 6), boxes(6), 20, 13, 13)
    out["coco19"] = (rs.randint(0, 80, 12), boxes(12), 80, 19, 19)
    out["single"] = (np.array([3]), boxes(1), 20, 13, 13)
    out["empty"] = (np.zeros([0], np.int64), np.zeros([0, 4], np.float32), 20, 13, 13)
    out["nonsquare"] = (rs.randint(0, 20, 5), boxes(5), 20, 7, 11)
    # several objects in one cell: the last one wins the box, class bits accumulate (incl. a repeated class)
    c = np.array([[0.50, 0.50, 0.56, 0.58], [0.40, 0.42, 0.66, 0.66], [0.52, 0.51, 0.55, 0.57], [0.10, 0.10, 0.30, 0.20],
                  [0.49, 0.47, 0.59, 0.61]], np.float32)
    out["same_cell"] = (np.array([1, 7, 1, 4, 19]), c, 20, 13, 13)
    out["crowd"] = (rs.randint(0, 20, 300), boxes(300, 0.01, 0.2), 20, 13, 13)          # > one block of threads, many collisions
    # centres exactly on cell borders (x*13 integral) and degenerate zero-size boxes
    k = np.array([2, 5, 9], np.float32) / np.float32(13)
    c = np.stack([k - np.float32(0.03125), k - np.float32(0.0625), k + np.float32(0.03125), k + np.float32(0.0625)], 1).astype(np.float32)
    c = np.concatenate([c, np.array([[0.25, 0.75, 0.25, 0.75]], np.float32)], 0)
    out["borders"] = (np.array([0, 1, 2, 3]), c, 20, 13, 13)
    return out


def main():
    ref = load_reference_transform_labels()
    out = {"numpy_version": np.__version__}
    for name, (cls, coord, classes, cw, ch) in cases().items():
        res = ref(cls, coord, classes, cw, ch)
        out[name + "_class"] = np.asarray(cls, np.int32)
        out[name + "_coord"] = coord
        out[name + "_meta"] = np.array([classes, cw, ch], np.int32)
        for k, v in zip(("mask", "prob", "coords", "offset_xy_min", "offset_xy_max", "areas"), res):
            assert v.dtype == np.float32, (name, k, v.dtype)
            out[name + "_" + k] = v
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **out)
    print(sorted(k for k in out if k.endswith("_mask")), "mask sums", {k: float(v.sum()) for k, v in out.items() if k.endswith("_mask")})


if __name__ == "__main__":
    main()
