"""CPU restatement of the reference's optimizer step (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Reference: train.py:70-72 (`get_optimizer` -> tf.train.AdamOptimizer(lr, beta1, beta2, epsilon)), train.py:118-124
(`tf.train.exponential_decay(args.learning_rate, global_step, decay_steps, decay_rate, staircase)`), train.py:127-129
(`slim.learning.create_train_op(total_loss, optimizer, global_step, clip_gradient_norm=args.gradient_clip)`),
config.ini [optimizer_adam] / [exponential_decay].

The arithmetic lives in TensorFlow 1.0 (not installable here).  PINNED to tests/golden/adam_reference.npz: the documented
TF-1.0 update evaluated two independent ways that agree to 1e-12 -- scalar float64 loops of the published formulas and
torch.optim.Adam (external implementation) with its epsilon re-mapped to TF's placement -- over 3 steps with extreme epsilon,
active / inactive / tiny clip norms and an all-zero gradient tensor (tests/golden/make_adam_golden.py,
tests/test_optimizer_oracle.py).  TF's own float32 kernel rounding stays unobservable.  Restated from TF-1.0's published kernels:
  * training_ops ApplyAdam functor:  alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
        m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  var -= (m * alpha) / (sqrt(v) + epsilon)
  * clip_ops.clip_by_norm (per tensor, as slim.learning.clip_gradient_norms applies it):
        t * clip_norm * minimum(rsqrt(reduce_sum(t*t)), 1 / clip_norm)
  * learning_rate_decay.exponential_decay:  lr * decay_rate ** (global_step / decay_steps), floor() if staircase
All float32, in that operation order.
"""
import numpy as np

f32 = np.float32


def exponential_decay_oracle(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
    p = f32(global_step) / f32(decay_steps)
    if staircase:
        p = np.floor(p)
    return f32(f32(learning_rate) * np.power(f32(decay_rate), p, dtype=np.float32))


def clip_by_norm_oracle(g, clip_norm):
    g = np.asarray(g, dtype=np.float32)
    sumsq = f32(np.sum(g.astype(np.float64) ** 2))          # the reduction order of a GPU differs; compared with tolerance
    l2inv = f32(1.0) / np.sqrt(sumsq, dtype=np.float32)
    return (g * f32(clip_norm)) * np.minimum(l2inv, f32(1.0) / f32(clip_norm))


def adam_oracle(params, grads, m, v, learning_rate, beta1, beta2, epsilon, t, clip_norm=0.0):
    """One step over lists of float32 arrays (updated copies returned): (params, m, v)."""
    b1, b2, eps = f32(beta1), f32(beta2), f32(epsilon)
    alpha = f32(float(learning_rate) * np.sqrt(1.0 - float(beta2) ** t) / (1.0 - float(beta1) ** t))
    out_p, out_m, out_v = [], [], []
    for p, g, mi, vi in zip(params, grads, m, v):
        g = np.asarray(g, dtype=np.float32)
        if clip_norm > 0:
            g = clip_by_norm_oracle(g, clip_norm)
        mi = (mi + (g - mi) * (f32(1) - b1)).astype(np.float32)
        vi = (vi + (g * g - vi) * (f32(1) - b2)).astype(np.float32)
        p = (p - (mi * alpha) / (np.sqrt(vi) + eps)).astype(np.float32)
        out_p.append(p); out_m.append(mi); out_v.append(vi)
    return out_p, out_m, out_v


def adam_golden_cases(path):
    """tests/golden/adam_reference.npz -> iterator of (name, (lr, beta1, beta2, eps, clip), p0[], g_steps[3][], p3[], m3[], v3[])."""
    g = np.load(path)
    n = len(g["shapes"])
    for name in g["case_names"]:
        name = str(name)
        lr, b1, b2, eps, clip = (float(x) for x in g[name + "_hyper"])
        yield (name, (lr, b1, b2, eps, clip), [g["%s_p0_%d" % (name, k)] for k in range(n)],
               [[g["%s_g%d_%d" % (name, t, k)] for k in range(n)] for t in range(3)],
               [g["%s_p3_%d" % (name, k)] for k in range(n)], [g["%s_m3_%d" % (name, k)] for k in range(n)],
               [g["%s_v3_%d" % (name, k)] for k in range(n)])
