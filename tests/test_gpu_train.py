"""-m gpu parity tests for the training step (BASELINE config 3): the tcgen05 weight-gradient GEMM in isolation
(vs torch fp64), and the whole step -- forward with batch-statistics BN, loss, backward -- through the
reference-shaped API (Builder(training=True) -> create_objectives -> backward) vs the CPU autograd oracle.

Tolerances: forward / loss 1e-4 relative (north_star); gradients max|a-b|/max|b| per tensor <= GRAD_TOL.
"""
import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import init_params
from oracle.train_oracle import train_step_oracle

pytestmark = pytest.mark.gpu
GRAD_TOL = 2e-4


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("b,hw,cin,k,cout,max_ctas", [
    (2, 13, 512, 3, 256, 0), (1, 32, 32, 3, 64, 0), (2, 26, 64, 3, 128, 0), (3, 13, 1024, 1, 425, 0),
    (1, 13, 3072, 3, 1024, 0), (2, 26, 128, 1, 64, 0), (2, 13, 512, 3, 256, -37), (4, 52, 128, 3, 256, 0)])
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


def _run_train_step(cuda, classes, size, batch, anchors, seed):
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
    ref = train_step_oracle(x, params, classes, anchors, labels, ho.HPARAM_DEFAULT)
    return builder, flat, grads, ref, store


@pytest.mark.parametrize("classes,size,batch,anchors,seed,fwd_tol", [
    (20, 160, 4, ho.ANCHORS_VOC, 1, 1e-4),      # >= 100 samples per channel in every layer: the 1e-4 target holds
    (20, 64, 4, ho.ANCHORS_VOC, 1, 5e-4),       # 16 samples per channel at the 2x2 layers: batch-stat BN amplifies rounding
    (20, 96, 3, ho.ANCHORS_VOC, 2, 5e-4), (80, 64, 2, ho.ANCHORS_COCO, 3, 5e-4)])
def test_train_step_vs_autograd_oracle(cuda, classes, size, batch, anchors, seed, fwd_tol):
    """Truth = the float64 autograd oracle.  The training step is discontinuous in its inputs (max-pool argmax,
    leaky sign at 0, best-anchor equality mask), so even torch-float32 differs from float64 by several per cent on
    a few gradient tensors at these tiny sizes.  Per-tensor criterion: our error <= max(GRAD_TOL, 4 x the error
    of the float32 oracle against the same float64 truth)."""
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
        assert abs(float(builder.objectives[k]) - v) <= 5 * fwd_tol * max(abs(v), 1e-9), k
    assert report["dnet"][0] <= max(5 * fwd_tol, 4 * report["dnet"][1])
    bad = {k: (v, floor[k]) for k, v in errs.items() if v > max(GRAD_TOL, 4 * floor[k])}
    assert not bad, bad
    assert flat.numel() == sum(v.size for v in ref["grads"].values())
    # slim UPDATE_OPS: moving averages follow the batch statistics (decay 0.999)
    for name, v in ref["new_moving"].items():
        got = store.global_variables()["yolo2_darknet/" + name].cpu().numpy()
        assert np.abs(got - v).max() <= 1e-5 * max(1.0, np.abs(v).max()), name


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
