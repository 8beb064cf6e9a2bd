"""Golden vectors that pin oracle/optimizer_oracle.py (and, through it, the device Adam step) to something EXTERNAL to this
repo -- TensorFlow 1.0 itself cannot be installed here.  Two independent evaluations of the documented update must agree
to 1e-12 before anything is written:

  (a) the documented TF-1.0 formulas (tf.train.AdamOptimizer docstring / training_ops ApplyAdam; clip_ops.clip_by_norm as
      slim.learning.create_train_op applies it per tensor, train.py:127-129) written out as scalar Python-float loops;
  (b) torch.optim.Adam (an external implementation) in float64, with its epsilon re-mapped per step: torch adds eps AFTER the
      bias correction of sqrt(v), TF before -- eps_torch(t) = eps_tf / sqrt(1 - beta2^t) makes the two updates identical.

Cases: default hyper-parameters (config.ini:35-38: beta1 0.9, beta2 0.999, epsilon 1e-8), extreme epsilon (1.0, 1e-3),
clip active / inactive / off, an all-zero gradient tensor under clipping (rsqrt(0) = inf; min(inf, 1/clip) = 1/clip),
3 steps each with fresh gradients.  Run:  python tests/golden/make_adam_golden.py  ->  tests/golden/adam_reference.npz
"""
import math
import os

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adam_reference.npz")
CASES = {
    # name: (lr, beta1, beta2, eps, clip_norm)
    "default": (1e-3, 0.9, 0.999, 1e-8, 0.0),
    "train_py_lr": (1e-6, 0.9, 0.999, 1e-8, 0.0),
    "eps_one": (1e-2, 0.9, 0.999, 1.0, 0.0),
    "eps_1e-3_clip_active": (1e-2, 0.8, 0.99, 1e-3, 0.5),
    "clip_inactive": (1e-3, 0.9, 0.999, 1e-8, 1e3),
    "clip_tiny": (1e-3, 0.9, 0.999, 1e-8, 1e-4),
}
SHAPES = [(3, 3, 2, 4), (4,), (4,), (1, 1, 4, 5), (5,)]


def clip_by_norm(g, clip):
    """clip_ops.clip_by_norm: t * clip_norm * minimum(rsqrt(sum(t*t)), 1/clip_norm)."""
    ss = sum(x * x for x in g)
    l2inv = float("inf") if ss == 0.0 else 1.0 / math.sqrt(ss)
    return [x * clip * min(l2inv, 1.0 / clip) for x in g]


def tf_formula(p, g_steps, lr, b1, b2, eps, clip):
    p, m, v = list(p), [0.0] * len(p), [0.0] * len(p)
    for t, g in enumerate(g_steps, 1):
        g = clip_by_norm(g, clip) if clip > 0 else list(g)
        alpha = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        for i in range(len(p)):
            m[i] = m[i] + (g[i] - m[i]) * (1.0 - b1)
            v[i] = v[i] + (g[i] * g[i] - v[i]) * (1.0 - b2)
            p[i] = p[i] - (m[i] * alpha) / (math.sqrt(v[i]) + eps)
    return p, m, v


def torch_adam(ps, g_steps, lr, b1, b2, eps, clip):
    params = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in ps]
    opt = torch.optim.Adam(params, lr=lr, betas=(b1, b2), eps=eps)
    for t, gs in enumerate(g_steps, 1):
        for group in opt.param_groups:
            group["eps"] = eps / math.sqrt(1.0 - b2 ** t)
        for p, g in zip(params, gs):
            g = torch.tensor(g, dtype=torch.float64)
            if clip > 0:
                n = g.norm()
                g = g * clip / max(float(n), clip) if float(n) > 0 else g
            p.grad = g
        opt.step()
    st = [opt.state[p] for p in params]
    return [p.detach().numpy() for p in params], [s["exp_avg"].numpy() for s in st], [s["exp_avg_sq"].numpy() for s in st]


def main():
    rs = np.random.RandomState(2017)
    out = {"case_names": np.array(sorted(CASES)), "shapes": np.array([str(s) for s in SHAPES])}
    for name, (lr, b1, b2, eps, clip) in sorted(CASES.items()):
        ps = [rs.normal(0, 1, size=s).astype(np.float32).astype(np.float64) for s in SHAPES]
        g_steps = []
        for t in range(3):
            gs = [(rs.normal(0, 10.0 ** rs.randint(-6, 2), size=s)).astype(np.float32).astype(np.float64) for s in SHAPES]
            gs[2] = np.zeros(SHAPES[2])                       # an all-zero gradient tensor (a BN beta nobody touches)
            g_steps.append(gs)
        a_p, a_m, a_v = [], [], []
        for k in range(len(SHAPES)):
            p, m, v = tf_formula(ps[k].ravel().tolist(), [g[k].ravel().tolist() for g in g_steps], lr, b1, b2, eps, clip)
            a_p.append(np.array(p).reshape(SHAPES[k])); a_m.append(np.array(m).reshape(SHAPES[k])); a_v.append(np.array(v).reshape(SHAPES[k]))
        b_p, b_m, b_v = torch_adam(ps, g_steps, lr, b1, b2, eps, clip)
        for x, y in zip(a_p + a_m + a_v, b_p + b_m + b_v):
            np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-300)
        out[name + "_hyper"] = np.array([lr, b1, b2, eps, clip])
        for k in range(len(SHAPES)):
            out["%s_p0_%d" % (name, k)] = ps[k].astype(np.float32)
            for t in range(3):
                out["%s_g%d_%d" % (name, t, k)] = g_steps[t][k].astype(np.float32)
            out["%s_p3_%d" % (name, k)] = a_p[k]
            out["%s_m3_%d" % (name, k)] = a_m[k]
            out["%s_v3_%d" % (name, k)] = a_v[k]
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
