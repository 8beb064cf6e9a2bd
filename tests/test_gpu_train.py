"""-m gpu parity tests for the training step (BASELINE config 3): the tcgen05 weight-gradient GEMM in isolation
(vs torch fp64), and the whole step -- forward with batch-statistics BN, loss, backward -- through the
reference-shaped API (Builder(training=True) -> create_objectives -> backward) vs the CPU autograd oracle.

Tolerances: forward / objectives / d(total)/d(net) 1e-4 relative (north_star), incl. BASELINE configs[2] at its full size
(B = 64, 416 x 416, 20 classes); end-to-end variable gradients per tensor against the float64 oracle, bounded by a multiple of
what the float32 evaluation of the same oracle achieves (the step is discontinuous: leaky sign, pool argmax, best-anchor mask).
"""
import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import init_params
from oracle.train_oracle import train_step_oracle
from tests_gpu_train_helpers import GRAD_TOL, check_gradients_like_float32, check_gradients_sane

pytestmark = pytest.mark.gpu


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("b,hw,cin,k,cout,max_ctas", [
    (2, 13, 512, 3, 256, 0), (1, 32, 32, 3, 64, 0), (2, 26, 64, 3, 128, 0), (3, 13, 1024, 1, 425, 0),
    (1, 13, 3072, 3, 1024, 0), (2, 26, 128, 1, 64, 0), (2, 13, 512, 3, 256, -37), (4, 52, 128, 3, 256, 0),
    # K-aligned split + cooperative hand-off: 1 tile x 148 k-ranges, 5 tiles x 29, 3 tiles x 49 (tap-packed Cin = 32)
    (8, 104, 128, 1, 64, 0), (4, 104, 64, 3, 128, 0), (2, 208, 32, 3, 64, 0)])
def test_wgrad_kernel_vs_fp64(cuda, b, hw, cin, k, cout, max_ctas):
    import torch
    import torch.nn.functional as F
    from yolo_tf_b200 import _lib
    rs = np.random.RandomState(cin + cout + hw)
    x = torch.as_tensor(rs.normal(0, 1, size=(b, hw, hw, cin)).astype(np.float32)).to(cuda)
    dy = torch.as_tensor(rs.normal(0, 1, size=(b, hw, hw, cout)).astype(np.float32)).to(cuda)
    dw = torch.full((k, k, cin, cout), float("nan"), device=cuda)
    _lib.check(_lib.lib().y2_conv2d_wgrad(_lib.ptr(x), b, hw, hw, cin, _lib.ptr(dy), k, cout, _lib.ptr(dw), max_ctas, None))
    torch.cuda.synchronize()
    xd = x.double().permute(0, 3, 1, 2).requires_grad_(False)
    w0 = torch.zeros(cout, cin, k, k, dtype=torch.float64, device=cuda, requires_grad=True)
    y = F.conv2d(xd, w0, padding=k // 2)
    y.backward(dy.double().permute(0, 3, 1, 2))
    ref = w0.grad.permute(2, 3, 1, 0).cpu().numpy()           # OIHW -> HWIO
    got = dw.cpu().numpy()
    assert not np.isnan(got).any()
    assert _rel(got, ref) <= 1e-4


def _run_train_step(cuda, classes, size, batch, anchors, seed, oracle=True):
    import torch
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder
    params = init_params(classes, 5, seed=seed)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    rs = np.random.RandomState(seed + 10)
    x = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    cw = size // 32
    labels = ho.synthetic_labels(batch, classes, cw, cw, seed=seed)
    builder = Builder.from_values([str(i) for i in range(classes)], size, size, anchors)
    builder(torch.from_numpy(x).to(cuda), training=True)
    builder.create_objectives(labels)
    flat, grads = builder.backward(allreduce=False)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    ref = train_step_oracle(x, params, classes, anchors, labels, ho.HPARAM_DEFAULT) if oracle else None
    return builder, flat, grads, ref, store


@pytest.mark.parametrize("classes,size,batch,anchors,seed,fwd_tol", [
    # batch-statistics BN amplifies every layer's error ~1.16x per following layer (it removes the noise-free mean and
    # renormalises): round 1's bf16 planes landed at 1.4-2.4e-4 here.  The training forward now runs on fp16 planes with
    # 8-k-block accumulation chains: measured 1.7-2.3e-5, torch-float32 itself 1.3-1.8e-5 -> the 1e-4 bar.
    (20, 160, 4, ho.ANCHORS_VOC, 1, 1e-4), (20, 64, 4, ho.ANCHORS_VOC, 1, 1e-4),
    (20, 96, 3, ho.ANCHORS_VOC, 2, 1e-4), (80, 64, 2, ho.ANCHORS_COCO, 3, 1e-4)])
def test_train_step_vs_autograd_oracle(cuda, classes, size, batch, anchors, seed, fwd_tol):
    """Truth = the float64 autograd oracle.  The training step is discontinuous in its inputs (max-pool argmax,
    leaky sign at 0, best-anchor equality mask), so even torch-float32 differs from float64 by several per cent on
    a few gradient tensors at these tiny sizes: the gradient bar is float32's own accuracy (_check_gradients_like_float32)."""
    import torch
    builder, flat, grads, ref, store = _run_train_step(cuda, classes, size, batch, anchors, seed)
    params = init_params(classes, 5, seed=seed)
    rs = np.random.RandomState(seed + 10)
    x = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    labels = ho.synthetic_labels(batch, classes, size // 32, size // 32, seed=seed)
    f32 = train_step_oracle(x, params, classes, anchors, labels, ho.HPARAM_DEFAULT, dtype=torch.float32)
    report = {"net": (_rel(builder.output.cpu().numpy(), ref["net"]), _rel(f32["net"], ref["net"])),
              "dnet": (_rel(builder.objectives.grad_inputs.cpu().numpy(), ref["dnet"]), _rel(f32["dnet"], ref["dnet"]))}
    errs, floor = {}, {}
    for name, g_ref in ref["grads"].items():
        errs[name] = _rel(grads["yolo2_darknet/" + name].cpu().numpy(), g_ref)
        floor[name] = _rel(f32["grads"][name], g_ref)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    print("forward/dnet (ours, fp32 oracle) vs fp64:", {k: ("%.1e" % a, "%.1e" % b) for k, (a, b) in report.items()})
    print("worst gradient errors (ours, fp32-oracle floor):", [(k, "%.1e" % v, "%.1e" % floor[k]) for k, v in worst])
    assert report["net"][0] <= fwd_tol
    for k, v in ref["objectives"].items():
        assert abs(float(builder.objectives[k]) - v) <= fwd_tol * max(abs(v), 1e-9), k
    assert report["dnet"][0] <= fwd_tol
    check_gradients_sane(grads, ref)              # accuracy bounds on the gradients: the two tests below
    assert flat.numel() == sum(v.size for v in ref["grads"].values())
    # slim UPDATE_OPS: moving averages follow the batch statistics (decay 0.999)
    for name, v in ref["new_moving"].items():
        got = store.global_variables()["yolo2_darknet/" + name].cpu().numpy()
        assert np.abs(got - v).max() <= 1e-5 * max(1.0, np.abs(v).max()), name


@pytest.mark.parametrize("classes,size,batch,seed,factor", [(20, 224, 16, 5, 4.0), (20, 416, 64, 1, 2.0)])
def test_train_step_at_baseline_config3_size_vs_fp64_oracle(cuda, classes, size, batch, seed, factor):
    """BASELINE configs[2] ITSELF -- B = 64, 416 x 416, 20 classes -- (and a mid-size case) against the float64 autograd oracle (the same restatement,
    oracle/train_oracle.py, evaluated by torch on the device: ~7 TFLOP of float64, minutes on the host cores).  Round 1 only
    tested this size for repeatability; the inference bug found at B = 32 showed that small-batch parity does not carry over.
    forward `net`, the 4 objectives and d(total)/d(net): 1e-4 (north_star).  Variable gradients: float32's own accuracy, per
    tensor and in aggregate (_check_gradients_like_float32; at this size float32 itself is 1e-4 .. 4e-2 off float64)."""
    import torch
    builder, flat, grads, _, store = _run_train_step(cuda, classes, size, batch, ho.ANCHORS_VOC, seed, oracle=False)
    params = init_params(classes, 5, seed=seed)
    x = np.random.RandomState(seed + 10).normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    labels = ho.synthetic_labels(batch, classes, size // 32, size // 32, seed=seed)
    ref = train_step_oracle(x, params, classes, ho.ANCHORS_VOC, labels, ho.HPARAM_DEFAULT, device="cuda")
    f32 = train_step_oracle(x, params, classes, ho.ANCHORS_VOC, labels, ho.HPARAM_DEFAULT, dtype=torch.float32, device="cuda")
    torch.cuda.empty_cache()
    net_err, dnet_err = _rel(builder.output.cpu().numpy(), ref["net"]), _rel(builder.objectives.grad_inputs.cpu().numpy(), ref["dnet"])
    print("B=%d %dx%d forward %.2e (fp32 oracle %.2e), dnet %.2e (fp32 %.2e)" % (batch, size, size, net_err, _rel(f32["net"], ref["net"]), dnet_err, _rel(f32["dnet"], ref["dnet"])))
    assert net_err <= 1e-4 and dnet_err <= 1e-4
    for k, v in ref["objectives"].items():
        assert abs(float(builder.objectives[k]) - v) <= 1e-4 * max(abs(v), 1e-9), k
    check_gradients_like_float32(grads, ref, f32, factor=factor)


def test_training_then_inference_uses_updated_moving_stats(cuda):
    import torch
    from oracle.darknet_oracle import darknet_oracle
    from yolo_tf_b200.model.yolo2 import inference
    builder, flat, grads, ref, store = _run_train_step(cuda, 20, 64, 2, ho.ANCHORS_VOC, 4)
    params = {k[len("yolo2_darknet/"):]: v.cpu().numpy() for k, v in store.global_variables().items()}
    rs = np.random.RandomState(0)
    x = rs.normal(0, 1, size=(1, 64, 64, 3)).astype(np.float32)
    _, out = inference.darknet(torch.from_numpy(x).to(cuda), 20, 5)
    expect = darknet_oracle(x, params, 20, 5)
    assert _rel(out.cpu().numpy(), expect) <= 1e-4


def test_backward_per_layer_teacher_forced(cuda):
    """Well-conditioned backward parity: every layer's backward is checked against torch float64 autograd GIVEN
    OUR OWN forward state (raw conv output z, layer input, incoming gradient dL/dy), so no discrete decision
    (leaky sign, pool argmax) can differ between the two sides.  Covers BN+leaky backward, wgrad, dgrad, the
    max-pool / reorg / concat gradient routing, for a representative set of layers."""
    import ctypes
    import torch
    import torch.nn.functional as F
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import inference
    classes, size, batch = 20, 96, 3
    builder, flat0, grads0, ref, store = _run_train_step(cuda, classes, size, batch, ho.ANCHORS_VOC, 2)
    L = _lib.lib()
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    geo = inference.layer_geometry(classes, 5)
    dnet = builder.objectives.grad_inputs
    rs = np.random.RandomState(12)
    x_img = torch.from_numpy(rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)).to(cuda)   # same seed/stream as _run_train_step
    V = store.global_variables()
    spatial, hh = [], size
    for (_, k, cin, cout, bn, pool) in geo:
        spatial.append(hh)
        if pool:
            hh //= 2

    def get(kind, layer, shape):
        t = torch.empty(shape, device=cuda)
        _lib.check(L.y2_train_get_tensor(eng.h, kind, layer, _lib.ptr(t), None))
        return t

    def backward_with_probe(layer):
        k, cin, cout = geo[layer][1:4]
        s = spatial[layer]
        gy = torch.zeros(batch, s, s, cout, device=cuda)
        gin = torch.zeros(batch, s, s, cin, device=cuda)
        _lib.check(L.y2_train_probe(eng.h, layer, _lib.ptr(gy), _lib.ptr(gin)))
        flat, views = eng.backward(dnet)
        torch.cuda.synchronize()
        _lib.check(L.y2_train_probe(eng.h, -1, None, None))
        return gy, gin, views

    report = {}
    probes = {}
    for layer in (20, 19, 17, 14, 13, 12, 8, 4, 1, 0):
        name, k, cin, cout, bn, pool = geo[layer]
        s = spatial[layer]
        gy, gin, views = backward_with_probe(layer)
        probes[layer] = (gy, gin)
        z = get(0, layer, (batch, s, s, cout)).double()
        if layer == 0:
            xin = x_img.double()
        elif layer == 20:
            xin = get(3, 0, (batch, s, s, 3072)).double()
        elif geo[layer - 1][5]:
            xin = get(2, layer - 1, (batch, s, s, cin)).double()
        else:
            xin = get(1, layer - 1, (batch, s, s, cin)).double()
        gamma = V["yolo2_darknet/%s/BatchNorm/gamma" % name].double()
        beta = V["yolo2_darknet/%s/BatchNorm/beta" % name].double()
        w = V["yolo2_darknet/%s/weights" % name].double()
        # BN(batch stats) + leaky backward on OUR z
        zl = z.clone().requires_grad_(True)
        g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        mean = zl.mean(dim=(0, 1, 2))
        var = ((zl - mean) ** 2).mean(dim=(0, 1, 2))
        inv = torch.rsqrt(var + 1e-5) * g_
        yb = zl * inv + (b_ - mean * inv)
        y = torch.maximum(yb, 0.1 * yb)
        (y * gy.double()).sum().backward()
        dz = zl.grad
        # conv backward with that dz
        xl = xin.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        wl = w.permute(3, 2, 0, 1).contiguous().requires_grad_(True)
        zc = F.conv2d(xl, wl, padding=k // 2)
        zc.backward(dz.permute(0, 3, 1, 2))
        r = {"z_vs_conv(xin)": _rel(z.cpu().numpy(), zc.detach().permute(0, 2, 3, 1).cpu().numpy()),
             "dgamma": _rel(views[name + "/BatchNorm/gamma"].cpu().numpy(), g_.grad.cpu().numpy()),
             "dbeta": _rel(views[name + "/BatchNorm/beta"].cpu().numpy(), b_.grad.cpu().numpy()),
             "dW": _rel(views[name + "/weights"].cpu().numpy(), wl.grad.permute(2, 3, 1, 0).cpu().numpy())}
        if layer > 0:
            r["dX"] = _rel(gin.cpu().numpy(), xl.grad.permute(0, 2, 3, 1).cpu().numpy())
        report[name] = r
    print("teacher-forced per-layer backward errors:", {k: {a: "%.1e" % b for a, b in v.items()} for k, v in report.items()})
    # gradient routing: conv13's input gradient -> max-pool backward (+ reorg of conv20's concat gradient) = conv12's dL/dy
    y12 = get(1, 12, (batch, spatial[12], spatial[12], 512)).double().permute(0, 3, 1, 2).requires_grad_(True)
    pooled = F.max_pool2d(y12, 2, 2)
    gcat = probes[20][1].double()                                     # [B, s13, s13, 3072]
    r12 = y12.permute(0, 2, 3, 1).reshape(batch, spatial[13], 2, spatial[13], 2, 512).permute(0, 1, 3, 2, 4, 5).reshape(batch, spatial[13], spatial[13], 2048)
    ((pooled * probes[13][1].double().permute(0, 3, 1, 2)).sum() + (r12 * gcat[..., :2048]).sum()).backward()
    route12 = _rel(probes[12][0].cpu().numpy(), y12.grad.permute(0, 2, 3, 1).cpu().numpy())
    route19 = _rel(probes[19][0].cpu().numpy(), gcat[..., 2048:].cpu().numpy())
    print("routing errors: conv12 (pool + reorg) %.1e, conv19 (concat slice) %.1e" % (route12, route19))
    worst = max(max(v.values()) for v in report.values())
    assert worst <= 2e-4, report
    assert route12 <= 1e-6 and route19 == 0.0


@pytest.mark.parametrize("classes,size,batch", [(20, 416, 64), (20, 96, 3)])
def test_full_size_backward_is_repeatable_and_linear(cuda, classes, size, batch):
    """BASELINE config 3 at its full size (B = 64, 416x416, 20 classes), through size-independent properties: given one
    forward state, the backward is a LINEAR map of the incoming gradient, so (a) calling it twice gives bit-identical
    buckets (fixed summation order everywhere: stream-K hand-offs, chain-cap running sums, two-stage reductions) and
    (b) doubling dL/dnet doubles every gradient bit for bit (a power-of-two scale commutes with the bf16 hi/lo split, with
    every rounding and with the tensor core's truncating accumulator)."""
    import torch
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder, inference
    params = init_params(classes, 5, seed=6)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(batch, size, size, 3, device="cuda", generator=g)
    cw = size // 32
    labels = [torch.from_numpy(t).to(cuda) for t in ho.synthetic_labels(batch, classes, cw, cw, seed=6)]
    builder = Builder.from_values([str(i) for i in range(classes)], size, size, ho.ANCHORS_VOC)
    builder(x, training=True)
    builder.create_objectives(labels)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    dnet = builder.objectives.grad_inputs
    f1, _ = eng.backward(dnet)
    f1 = f1.clone()
    f1b, _ = eng.backward(dnet)
    f1b = f1b.clone()
    f2, _ = eng.backward((dnet * 2.0).contiguous())
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    assert torch.isfinite(f1).all() and float(f1.abs().max()) > 0
    assert torch.equal(f1, f1b)
    assert torch.equal(f2, f1 * 2.0)
