"""Shared by the -m gpu training tests (test_gpu_train.py, test_gpu_tiny.py): the end-to-end gradient criterion."""
import numpy as np

GRAD_TOL = 2e-4


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def check_gradients_like_float32(grads, ref, f32, prefix="yolo2_darknet/"):
    """End-to-end variable gradients against the float64 oracle.  The step is discontinuous (leaky sign at 0, pool argmax,
    best-anchor equality), so a handful of decisions flip under ANY rounding and a flipped decision moves a gradient tensor by
    per cent: the float32 evaluation of the same oracle is 1e-3 .. 3e-1 off float64 on these tensors, at places that differ
    from ours.  The bar is therefore float32's own accuracy, per tensor and in aggregate:
      * every tensor: max-norm error <= max(GRAD_TOL, 2 x the WORST tensor error of float32), relative L2 error likewise;
      * the median over the tensors <= 2 x float32's median (max-norm and L2).
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
        assert a <= max(GRAD_TOL, 2 * f_max.max()), (name, a, f_max.max())
        assert c <= max(GRAD_TOL, 2 * f_l2.max()), (name, c, f_l2.max())
    assert np.median(e_max) <= max(GRAD_TOL, 2 * np.median(f_max)) and np.median(e_l2) <= max(GRAD_TOL, 2 * np.median(f_l2))
