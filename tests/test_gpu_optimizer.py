"""-m gpu parity of the device optimizer step (y2_adam_step) against oracle/optimizer_oracle.py on the real 65-tensor
bucket of the 20-class model, and of the reference-shaped train op (yolo_tf_b200.optimizer.create_train_op)."""
import ctypes

import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import init_params
from oracle.optimizer_oracle import adam_oracle

pytestmark = pytest.mark.gpu


def _engine(classes):
    import torch
    from yolo_tf_b200.model.yolo2 import inference
    return inference._Engine.get(torch.device("cuda:0"), classes, 5)


def _tensor_sizes(eng):
    sizes = []
    for (k, cin, cout, bn) in eng.layers:
        sizes.append(k * k * cin * cout)
        sizes += [cout, cout] if bn else [cout]
    return sizes


@pytest.mark.parametrize("clip", [0.0, 0.5])
def test_adam_step_vs_oracle(cuda, clip):
    import torch
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    eng = _engine(20)
    sizes = _tensor_sizes(eng)
    n = sum(sizes)
    assert n == L.y2_param_count(eng.h) and len(sizes) == L.y2_num_param_tensors(eng.h)
    rs = np.random.RandomState(3)
    params = [rs.normal(0, 0.05, size=s).astype(np.float32) for s in sizes]
    m = [np.zeros(s, np.float32) for s in sizes]
    v = [np.zeros(s, np.float32) for s in sizes]
    dp = [torch.from_numpy(p.copy()).to(cuda) for p in params]
    dm, dv = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    ptrs = (ctypes.c_void_p * len(dp))(*[t.data_ptr() for t in dp])
    need = L.y2_adam_workspace_bytes(eng.h)
    ws = torch.empty(need + 256, dtype=torch.uint8, device=cuda)
    off = (-ws.data_ptr()) % 256
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    for t in (1, 2, 3):
        grads = [(rs.normal(0, 1, size=s) * (10.0 ** rs.uniform(-3, 0))).astype(np.float32) for s in sizes]
        flat = torch.from_numpy(np.concatenate(grads)).to(cuda)
        _lib.check(L.y2_adam_step(eng.h, _lib.ptr(flat), _lib.ptr(dm), _lib.ptr(dv), ptrs, len(dp), lr, b1, b2, eps, t, clip,
                                  ctypes.c_void_p(ws.data_ptr() + off), need, None))
        torch.cuda.synchronize()
        old = params
        params, m, v = adam_oracle(params, grads, m, v, lr, b1, b2, eps, t, clip)
        gm, gv = dm.cpu().numpy(), dv.cpu().numpy()
        o = 0
        for i, s in enumerate(sizes):
            got = dp[i].cpu().numpy()
            step_ref = params[i] - old[i]
            # the update itself (not the parameter it is added to) carries the arithmetic: compare it to 1e-4 relative
            # + one float32 ulp of the parameter the update is added to
            assert np.abs((got - old[i]) - step_ref).max() <= 1e-4 * max(np.abs(step_ref).max(), 1e-12) + 1.5e-7 * max(np.abs(old[i]).max(), 1e-2), (t, i)
            assert np.abs(gm[o:o + s] - m[i]).max() <= 1e-5 * max(np.abs(m[i]).max(), 1e-30), (t, i)
            assert np.abs(gv[o:o + s] - v[i]).max() <= 1e-5 * max(np.abs(v[i]).max(), 1e-30), (t, i)
            params[i] = got.copy()                 # teacher-force: the next step starts from the device state
            o += s
        m = [gm[sum(sizes[:i]):sum(sizes[:i + 1])].copy() for i in range(len(sizes))]
        v = [gv[sum(sizes[:i]):sum(sizes[:i + 1])].copy() for i in range(len(sizes))]


def test_adam_step_vs_the_external_golden(cuda):
    """tests/golden/adam_reference.npz (the documented TF-1.0 update evaluated by scalar float64 loops AND by torch.optim.Adam,
    tests/golden/make_adam_golden.py) embedded in the real 65-tensor bucket: golden tensor k occupies the head of bucket tensor
    k, the rest of that tensor is zero (zero gradient, zero moments: it neither moves nor changes the tensor's clip norm)."""
    import torch
    import os
    from oracle.optimizer_oracle import adam_golden_cases
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    eng = _engine(20)
    sizes = _tensor_sizes(eng)
    n = sum(sizes)
    need = L.y2_adam_workspace_bytes(eng.h)
    ws = torch.empty(need + 256, dtype=torch.uint8, device=cuda)
    off = (-ws.data_ptr()) % 256
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for name, (lr, b1, b2, eps, clip), p0, g_steps, p3, m3, v3 in adam_golden_cases(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adam_reference.npz")):
        slots = [0, 1, 2, 3, 4]                                   # conv0 weights / gamma / beta, conv1 weights / gamma
        assert all(p0[k].size <= sizes[slots[k]] for k in range(len(p0)))
        dp = [torch.zeros(s, device=cuda) for s in sizes]
        for k, sl in enumerate(slots):
            dp[sl][:p0[k].size] = torch.from_numpy(p0[k].ravel()).to(cuda)
        dm, dv = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
        ptrs = (ctypes.c_void_p * len(dp))(*[t.data_ptr() for t in dp])
        for t in (1, 2, 3):
            flat = torch.zeros(n, device=cuda)
            for k, sl in enumerate(slots):
                flat[offs[sl]:offs[sl] + p0[k].size] = torch.from_numpy(g_steps[t - 1][k].ravel()).to(cuda)
            _lib.check(L.y2_adam_step(eng.h, _lib.ptr(flat), _lib.ptr(dm), _lib.ptr(dv), ptrs, len(dp), lr, b1, b2, eps, t, clip,
                                      ctypes.c_void_p(ws.data_ptr() + off), need, None))
        torch.cuda.synchronize()
        for k, sl in enumerate(slots):
            sz = p0[k].size
            got = dp[sl][:sz].cpu().numpy().astype(np.float64)
            upd, ref = got - p0[k].ravel(), (p3[k] - p0[k]).ravel()
            assert np.abs(upd - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-30) + 2.5e-7 * max(1.0, np.abs(p0[k]).max()), (name, k)
            gm = dm[offs[sl]:offs[sl] + sz].cpu().numpy()
            assert np.abs(gm - m3[k].ravel()).max() <= 1e-5 * max(np.abs(m3[k]).max(), 1e-300), (name, "m", k)
            # float32(1 - beta2) is 1.3e-5 off the ideal 0.001 (TF computes it in float32 too): v sits that far from the float64 golden
            gv = dv[offs[sl]:offs[sl] + sz].cpu().numpy()
            assert np.abs(gv - v3[k].ravel()).max() <= 3e-5 * max(np.abs(v3[k]).max(), 1e-300), (name, "v", k)
            assert float(dp[sl][sz:].abs().max()) == 0.0 if sz < sizes[sl] else True


def test_train_op_applies_adam_to_the_store(cuda):
    """create_train_op(...)(data, labels): the variables of the store move by exactly the Adam step of the gradients the
    same call produced (tensor order, clip, schedule, global_step), and the next forward uses the new weights."""
    import torch
    from yolo_tf_b200 import variables
    from yolo_tf_b200.model.yolo2 import Builder
    from yolo_tf_b200.optimizer import AdamOptimizer, create_train_op, exponential_decay
    classes, size, batch = 20, 64, 2
    params = init_params(classes, 5, seed=7)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    hp = dict(ho.HPARAM_DEFAULT)
    builder = Builder.from_values([str(i) for i in range(classes)], size, size, ho.ANCHORS_VOC, hparam=hp)
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)).to(cuda)
    labels = [torch.from_numpy(t).to(cuda) for t in ho.synthetic_labels(batch, classes, size // 32, size // 32, 9)]
    opt = AdamOptimizer(lambda step: exponential_decay(1e-3, step, 1, 0.5, True))
    train_op = create_train_op(builder, opt, global_step=0, clip_gradient_norm=0.25)
    state_m = state_v = None
    for step in range(2):
        before = {k: v.clone() for k, v in store.global_variables().items()}
        loss = train_op(x, labels)
        torch.cuda.synchronize()
        assert np.isfinite(float(loss))
        flat, views = builder.backward(allreduce=False)       # same forward state -> the same gradients the op consumed
        names = list(views.keys())
        g = [views[n].reshape(-1).cpu().numpy() for n in names]
        p0 = [before[n].reshape(-1).cpu().numpy() for n in names]
        if state_m is None:
            state_m = [np.zeros_like(a) for a in p0]
            state_v = [np.zeros_like(a) for a in p0]
        lr = exponential_decay(1e-3, step, 1, 0.5, True)
        p1, state_m, state_v = adam_oracle(p0, g, state_m, state_v, lr, 0.9, 0.999, 1e-8, step + 1, 0.25)
        after = store.global_variables()
        for n, a, b0 in zip(names, p1, p0):
            got = after[n].reshape(-1).cpu().numpy()
            ref_step = a - b0
            assert np.abs((got - b0) - ref_step).max() <= 2e-4 * max(np.abs(ref_step).max(), 1e-12) + 1.5e-7 * max(np.abs(b0).max(), 1e-2), (step, n)
        moved = max(float((after[n] - before[n]).abs().max()) for n in names)
        assert moved > 0
    assert train_op.global_step == 2
