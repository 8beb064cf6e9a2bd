"""Shared by the -m gpu training tests (test_gpu_train.py, test_gpu_tiny.py): the end-to-end gradient criterion."""
import numpy as np

GRAD_TOL = 2e-4


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def check_gradients_like_float32(grads, ref, f32, prefix="yolo2_darknet/", factor=4.0):
    """End-to-end variable gradients against the float64 oracle.  The step is discontinuous (leaky sign at 0, pool argmax,
    best-anchor equality), so a handful of decisions flip under ANY rounding and a flipped decision moves a gradient tensor by
    per cent: the float32 evaluation of the same oracle is 1e-3 .. 3e-1 off float64 on these tensors, at places that differ
    from ours.  The bar is therefore float32's own accuracy, per tensor and in aggregate:
      * every tensor: max-norm error <= max(GRAD_TOL, factor x the WORST tensor error of float32), relative L2 error likewise;
      * the median over the tensors <= factor x float32's median (max-norm and L2).
    ONE flipped decision deep in the network perturbs every upstream tensor at once, so with the handful of flips of a tiny
    batch the ratio ours / float32 is itself a coin toss (measured 0.4 .. 10^4 at B <= 4, where float32 sometimes has no flip at
    all): those sizes only get check_gradients_sane.  This bound is applied where there are hundreds of values per channel:
    B = 16 at 224 x 224 (factor 4) and BASELINE configs[2], B = 64 at 416 x 416 (factor 2; measured 1.0 .. 1.4).  The kernels
    are bit-reproducible, and so is this test.
    (The well-conditioned, strict per-layer check is test_gpu_train.py::test_backward_per_layer_teacher_forced.)"""
    rows = []
    for name, g_ref in ref["grads"].items():
        got = grads[prefix + name].cpu().numpy().astype(np.float64)
        g64, g32 = np.asarray(g_ref, np.float64), np.asarray(f32["grads"][name], np.float64)
        nrm = max(np.linalg.norm(g64), 1e-300)
        rows.append((name, _rel(got, g64), _rel(g32, g64), np.linalg.norm(got - g64) / nrm, np.linalg.norm(g32 - g64) / nrm))
    e_max, f_max, e_l2, f_l2 = (np.array([r[k] for r in rows]) for k in (1, 2, 3, 4))
    print("gradients vs fp64, max-norm: ours worst %.1e median %.1e | fp32 oracle worst %.1e median %.1e" % (e_max.max(), np.median(e_max), f_max.max(), np.median(f_max)))
    print("gradients vs fp64, rel. L2 : ours worst %.1e median %.1e | fp32 oracle worst %.1e median %.1e" % (e_l2.max(), np.median(e_l2), f_l2.max(), np.median(f_l2)))
    for name, a, _, c, _ in rows:
        assert a <= max(GRAD_TOL, factor * f_max.max()), (name, a, f_max.max())
        assert c <= max(GRAD_TOL, factor * f_l2.max()), (name, c, f_l2.max())
    assert np.median(e_max) <= max(GRAD_TOL, factor * np.median(f_max)) and np.median(e_l2) <= max(GRAD_TOL, factor * np.median(f_l2))


def check_gradients_sane(grads, ref, prefix="yolo2_darknet/", min_cos=0.98):
    """Tiny batches (a few dozen values per channel in the deep layers): ONE flipped decision moves a gradient tensor by tens
    of per cent, in float32 as much as here, and whether a flip happens is a coin toss -- no accuracy bound is meaningful.
    Sanity of the plumbing only: every tensor points the way the float64 truth does."""
    cos = {}
    for name, g_ref in ref["grads"].items():
        a = grads[prefix + name].cpu().numpy().astype(np.float64).ravel()
        b = np.asarray(g_ref, dtype=np.float64).ravel()
        cos[name] = float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))
    print("lowest cosine(ours, fp64 truth):", sorted(cos.items(), key=lambda kv: kv[1])[:4])
    assert min(cos.values()) >= min_cos, sorted(cos.items(), key=lambda kv: kv[1])[:4]
