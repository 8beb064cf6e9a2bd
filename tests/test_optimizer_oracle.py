"""CPU tests of the optimizer oracle (oracle/optimizer_oracle.py) and of the host-side schedule in
yolo_tf_b200/optimizer.py.  The reference delegates this arithmetic to TensorFlow 1.0 (absent); the oracle is pinned to
tests/golden/adam_reference.npz -- the documented TF-1.0 update evaluated two independent ways that agree to 1e-12 (scalar
float64 loops of the published formulas, and torch.optim.Adam with its epsilon re-mapped; tests/golden/make_adam_golden.py) --
and checked against hand-derived values."""
import math
import os

import numpy as np
import pytest

ADAM_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adam_reference.npz")

from oracle.optimizer_oracle import adam_golden_cases, adam_oracle, clip_by_norm_oracle, exponential_decay_oracle


def test_exponential_decay_matches_formula_and_host_schedule():
    from yolo_tf_b200.optimizer import exponential_decay
    # config.ini [exponential_decay]: decay_steps 100000, decay_rate 0.96, staircase 1; train.py default lr 1e-6
    for step in (0, 1, 99999, 100000, 250000):
        want = 1e-6 * 0.96 ** math.floor(step / 100000)
        assert abs(exponential_decay(1e-6, step, 100000, 0.96, True) - want) <= 1e-18
        assert abs(float(exponential_decay_oracle(1e-6, step, 100000, 0.96, True)) - want) <= 1e-12 * 1e-6 + 1e-13
    assert abs(exponential_decay(0.1, 50, 100, 0.5, False) - 0.1 * 0.5 ** 0.5) <= 1e-15


def test_clip_by_norm_oracle():
    g = np.array([3.0, 4.0], dtype=np.float32)                       # norm 5
    np.testing.assert_allclose(clip_by_norm_oracle(g, 10.0), g, rtol=1e-6)          # below the clip: unchanged
    np.testing.assert_allclose(clip_by_norm_oracle(g, 1.0), g / 5.0, rtol=1e-6)     # above: rescaled to norm 1


def test_adam_oracle_first_steps_by_hand():
    p0, g = np.array([1.0], np.float32), np.array([0.5], np.float32)
    lr, b1, b2, eps = 0.1, 0.9, 0.999, 1e-8
    (p1,), (m1,), (v1,) = adam_oracle([p0], [g], [np.zeros(1, np.float32)], [np.zeros(1, np.float32)], lr, b1, b2, eps, 1)
    assert abs(m1[0] - 0.05) < 1e-7 and abs(v1[0] - 0.00025) < 1e-8      # float32(1 - 0.999) = 0.00099998713
    alpha = lr * math.sqrt(1 - b2) / (1 - b1)
    assert abs(p1[0] - (1.0 - alpha * 0.05 / (math.sqrt(0.00025) + eps))) < 1e-6     # = 1 - lr for a constant gradient
    (p2,), (m2,), (v2,) = adam_oracle([p1], [g], [m1], [v1], lr, b1, b2, eps, 2)
    assert abs(m2[0] - (0.05 + (0.5 - 0.05) * 0.1)) < 1e-7
    assert abs(p2[0] - (1.0 - 2 * lr)) < 1e-5


def test_get_optimizer_surface():
    import configparser
    from yolo_tf_b200.optimizer import AdamOptimizer, get_optimizer
    cfg = configparser.ConfigParser()
    cfg.read_dict({"optimizer_adam": {"beta1": "0.9", "beta2": "0.999", "epsilon": "1e-8"}})
    opt = get_optimizer(cfg, "adam")(lambda step: 1e-3 * 0.5 ** step)
    assert isinstance(opt, AdamOptimizer) and opt.beta1 == 0.9 and opt.epsilon == 1e-8
    assert opt.rate(0) == 1e-3 and opt.rate(2) == 2.5e-4
    try:
        get_optimizer(cfg, "momentum")
        raise AssertionError("momentum must not silently fall back")
    except NotImplementedError:
        pass


@pytest.mark.parametrize("case", list(adam_golden_cases(ADAM_GOLD)), ids=lambda c: c[0])
def test_adam_oracle_matches_the_external_golden(case):
    """3 steps incl. extreme epsilon, active / inactive / tiny clip and an all-zero gradient tensor under clipping."""
    name, (lr, b1, b2, eps, clip), p, g_steps, p3, m3, v3 = case
    m = [np.zeros_like(x) for x in p]
    v = [np.zeros_like(x) for x in p]
    p0 = [x.copy() for x in p]
    for t in (1, 2, 3):
        p, m, v = adam_oracle(p, g_steps[t - 1], m, v, lr, b1, b2, eps, t, clip)
    for k in range(len(p)):
        assert p[k].dtype == np.float32
        # the float32 oracle against the float64 golden: the UPDATE (p3 - p0) to 2e-6 relative + an ulp of the parameter
        # float32(1 - 0.999) = 0.00099998713: TF's float32 kernel (and the oracle) carry v 1.3e-5 below the ideal-arithmetic
        # golden, the update 0.65e-5 above it -- the bars are set just over that
        np.testing.assert_allclose(p[k] - p0[k], p3[k] - p0[k], rtol=1.5e-5, atol=2.5e-7 * max(1.0, float(np.abs(p0[k]).max())), err_msg="%s p[%d]" % (name, k))
        # moments: per tensor max-norm (an m that nearly cancels over the 3 steps has no relative accuracy of its own)
        assert np.abs(m[k] - m3[k]).max() <= 2e-6 * max(np.abs(m3[k]).max(), 1e-300), "%s m[%d]" % (name, k)
        assert np.abs(v[k] - v3[k]).max() <= 3e-5 * max(np.abs(v3[k]).max(), 1e-300), "%s v[%d]" % (name, k)
