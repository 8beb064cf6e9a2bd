"""Generates tests/golden/train_reference.npz: ONE TRAINING STEP's loss and gradients obtained by running the REFERENCE'S OWN
source end to end -- `darknet(net, classes, num_anchors, training=True)` (model/yolo2/inference.py:61-120), `Model(...,
training=True)` and `Objectives(...)` (model/yolo2/__init__.py:27-94), the `[yolo2_hparam]` weighting of
`Builder.create_objectives` (:114-119, config.ini:98-102) -- with the torch float64 stand-ins of make_backbone_golden.py /
make_head_golden.py for the slim / tf calls, and torch autograd in the place of tf.gradients (train.py:127-129).
The gradients of all 107 variables (67 M numbers) do not belong in a fixture: per variable the file keeps the L2 norm, the
first 8 entries and the dot product with a fixed pseudo-random vector (seeded from the variable's size), plus the loss,
the four objectives, d(total)/d(net) and the network output in full.  The checkpoint and labels are regenerated from their
seeds by the test.  Run once, here:   python tests/golden/make_train_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_backbone_golden as mb  # noqa: E402
import make_head_golden as mh  # noqa: E402

CLASSES, ANCHORS_N, SEED_LABELS = 20, 5, 3


def summary(name, g):
    flat = np.asarray(g, dtype=np.float64).reshape(-1)
    probe = np.random.RandomState(flat.size % (2 ** 31)).normal(size=flat.size)
    return np.concatenate([[np.sqrt((flat ** 2).sum())], flat[:8] if flat.size >= 8 else np.pad(flat, (0, 8 - flat.size)), [flat @ probe]])


def one_step(func, x):
    """One training step of the reference's own `func` ('darknet' | 'tiny') graph + Model + Objectives -> dict of arrays."""
    from oracle import head_oracle as ho
    labels = ho.synthetic_labels(x.shape[0], CLASSES, x.shape[2] // 32, x.shape[1] // 32, seed=SEED_LABELS)
    params = mb.checkpoint(func, CLASSES, ANCHORS_N, True)
    g = mb.Graph(params, True)
    leaves = {}
    base_var = g.var

    def var(name, shape):                     # every variable the reference graph creates becomes an autograd leaf
        t = base_var(name, shape)
        if name not in leaves:
            leaves[name] = t.clone().requires_grad_("moving" not in name)
        return leaves[name]
    g.var = var
    tf = mh.make_tf()
    mh.DTYPE[0] = torch.float64
    slim = mb.make_slim(tf, g)
    import ast
    import inspect
    ns_l, ns_r = {"tf": tf}, {"tf": tf, "np": np}
    exec(compile(ast.Module(body=[n for n in ast.parse(open(mb.REF_FN1).read()).body if isinstance(n, ast.FunctionDef) and n.name == "leaky_relu"],
                            type_ignores=[]), mb.REF_FN1, "exec"), ns_l)
    exec(compile(ast.Module(body=[n for n in ast.parse(open(mb.REF_FN2).read()).body if isinstance(n, ast.FunctionDef) and n.name == "reorg"],
                            type_ignores=[]), mb.REF_FN2, "exec"), ns_r)
    ns = {"tf": tf, "slim": slim, "inspect": inspect, "leaky_relu": ns_l["leaky_relu"], "reorg": ns_r["reorg"], "__name__": "model.yolo2.inference"}
    exec(compile(ast.Module(body=[n for n in ast.parse(open(mb.REF_INF).read()).body if isinstance(n, ast.FunctionDef) and n.name == func],
                            type_ignores=[]), mb.REF_INF, "exec"), ns)
    # Model / Objectives of the reference, sharing the same tf stand-in
    ns2 = {"np": np, "tf": tf, "yolo": None}
    tree2 = ast.parse(open(mh.REF2).read())
    exec(compile(ast.Module(body=[n for n in tree2.body if isinstance(n, ast.ClassDef) and n.name in ("Model", "Objectives")], type_ignores=[]),
                 mh.REF2, "exec"), ns2)
    scope, net = ns[func](mh.T(torch.as_tensor(x, dtype=torch.float64)), CLASSES, ANCHORS_N, True)
    net.v.retain_grad()
    model = ns2["Model"](net, CLASSES, ho.ANCHORS_VOC, training=True)
    obj = ns2["Objectives"](model, *[mh.T(torch.tensor(np.asarray(l), dtype=torch.float64)) for l in labels])
    total = sum(obj[k].v * mh.HPARAM[k] for k in mh.HPARAM)
    total.backward()
    arrays = {"x": x, "meta": np.array([CLASSES, ANCHORS_N, SEED_LABELS]), "net": net.v.detach().numpy(), "dnet": net.v.grad.numpy(),
              "total": np.array(float(total.detach()))}
    for k in mh.HPARAM:
        arrays["obj_" + k] = np.array(float(obj[k].v.detach()))
    names = [n for n, t in leaves.items() if t.requires_grad]
    assert all(leaves[n].grad is not None for n in names), [n for n in names if leaves[n].grad is None]
    arrays["grad_names"] = np.array(names)
    arrays["grad_summary"] = np.stack([summary(n, leaves[n].grad.numpy()) for n in names])
    return arrays


def main():
    rs = np.random.RandomState(29)
    x = rs.normal(0, 1, size=(2, 64, 96, 3)).astype(np.float32)                  # 2 x 3 cells, batch 2
    arrays = one_step("darknet", x)
    # tiny() (inference.py:25-50) through the same machinery, incl. its stride-1 SAME max-pool; keys prefixed "tiny_"
    xt = np.random.RandomState(31).normal(0, 1, size=(3, 96, 64, 3)).astype(np.float32)
    arrays.update({"tiny_" + k: v for k, v in one_step("tiny", xt).items()})
    np.savez_compressed(os.path.join(HERE, "train_reference.npz"), **arrays)
    print("darknet total", float(arrays["total"]), "tiny total", float(arrays["tiny_total"]), len(arrays["grad_names"]), "+", len(arrays["tiny_grad_names"]), "gradients")
    print("wrote train_reference.npz %.0f KiB" % (os.path.getsize(os.path.join(HERE, "train_reference.npz")) / 1024))


if __name__ == "__main__":
    main()
