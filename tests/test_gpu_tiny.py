"""-m gpu parity tests for the tiny YOLOv2 backbone (model/yolo2/inference.py:25-50): the same sm_100a
kernels behind a different layer table -- conv0 stored with 32 channels (16 logical + 16 exact zeros), the halo-tile conv
for the 32-input-channel 3x3 layer, the fused 2x2/2 max-pool epilogues, and the 2x2 stride-1 SAME max-pool (:42).

Tolerance (north_star): 1e-4 relative, max|a-b| / max|b| per tensor, against the CPU oracle (torch fp32, oneDNN).
"""
import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle.darknet_oracle import init_params, max_pool_s1_same_oracle, tiny_layer_table, tiny_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


def _setup_store(params, scope="yolo2_tiny"):
    from yolo_tf_b200 import variables
    store = variables.reset_default_store()
    store.assign({scope + "/" + k: v for k, v in params.items()})
    return store


@pytest.mark.parametrize("classes,size,batch", [(20, 416, 2), (80, 416, 1), (20, 64, 3), (20, 608, 1), (20, 64, 32)])
def test_tiny_forward_layer_by_layer_vs_oracle(cuda, classes, size, batch):
    import torch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import inference
    table = tiny_layer_table(classes, 5)
    params = init_params(classes, 5, seed=7, table=table)
    _setup_store(params)
    rs = np.random.RandomState(8)
    x = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    taps = {}
    ref = tiny_oracle(x, params, classes, 5, taps=taps)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5, inference.ARCH_TINY)
    assert [(k, cin, cout) for k, cin, cout, bn in eng.layers] == [(k, cin, cout) for _, k, cin, cout, _ in table]
    xd = torch.from_numpy(x).to(cuda)
    worst = {}
    for fuse in (0, 1):
        _lib.check(_lib.lib().y2_set_option(eng.h, b"fuse_pool", fuse))
        scope, out = inference.tiny(xd, classes, 5)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().y2_check_async_errors())
        assert scope == "yolo2_tiny" and tuple(out.shape) == (batch, size // 32, size // 32, 5 * (5 + classes))
        tag = "" if fuse == 0 else "[fused]"
        for i, (name, k, cin, cout, then) in enumerate(table[:-1]):
            if i == 0:
                got = eng.activation(0, True, taps["conv0/pool"].shape).cpu().numpy()
                worst["conv0/pool" + tag] = _rel(got, taps["conv0/pool"].astype(np.float64))
                continue
            if fuse == 0 or then == "pool_s1" or then is None:
                got = eng.activation(i, False, taps[name].shape).cpu().numpy()
                worst[name + tag] = _rel(got, taps[name].astype(np.float64))
            if then in ("pool", "pool_s1"):
                gp = eng.activation(i, True, taps[name + "/pool"].shape).cpu().numpy()
                worst[name + "/pool" + tag] = _rel(gp, taps[name + "/pool"].astype(np.float64))
        worst["output" + tag] = _rel(out.cpu().numpy(), ref.astype(np.float64))
    _lib.check(_lib.lib().y2_set_option(eng.h, b"fuse_pool", 1))
    print("tiny per-layer rel err:", {k: "%.1e" % v for k, v in worst.items()})
    assert max(worst.values()) <= TOL, worst


def test_tiny_stride1_pool_is_exact_selection(cuda):
    """conv5's activation and its stride-1 SAME pool come from the same forward: the pooled tensor must be EXACTLY the
    clipped-window maximum of the un-pooled one (pure selection on the stored hi+lo values, no arithmetic)."""
    import torch
    from yolo_tf_b200.model.yolo2 import inference
    classes, size, batch = 20, 96, 2
    table = tiny_layer_table(classes, 5)
    _setup_store(init_params(classes, 5, seed=9, table=table))
    x = np.random.RandomState(10).normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    inference.tiny(torch.from_numpy(x).to(cuda), classes, 5)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5, inference.ARCH_TINY)
    a = eng.activation(5, False, (batch, 3, 3, 512)).cpu()
    p = eng.activation(5, True, (batch, 3, 3, 512)).cpu()
    want = max_pool_s1_same_oracle(a.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert torch.equal(p, want)


def test_tiny_padded_channels_are_exact_zeros_and_variables_keep_logical_shapes(cuda):
    """conv0 has 16 output channels stored as 32: the reference-shaped variables keep their logical shapes
    ([3,3,3,16], [3,3,16,32]) and the result must not depend on what the caller's buffers held before."""
    import torch
    from yolo_tf_b200 import variables
    from yolo_tf_b200.model.yolo2 import inference
    classes = 20
    table = tiny_layer_table(classes, 5)
    params = init_params(classes, 5, seed=11, table=table)
    store = _setup_store(params)
    x = np.random.RandomState(12).normal(0, 1, size=(1, 64, 64, 3)).astype(np.float32)
    xd = torch.from_numpy(x).to(cuda)
    _, out1 = inference.tiny(xd, classes, 5)
    gv = store.global_variables()
    assert tuple(gv["yolo2_tiny/conv0/weights"].shape) == (3, 3, 3, 16)
    assert tuple(gv["yolo2_tiny/conv1/weights"].shape) == (3, 3, 16, 32)
    assert tuple(gv["yolo2_tiny/conv/weights"].shape) == (1, 1, 1024, 125)
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5, inference.ARCH_TINY)
    eng.ws.fill_(0xFF)                                   # poison the workspace (NaN patterns in every plane)
    _, out2 = inference.tiny(xd, classes, 5)
    assert torch.equal(out1, out2)
    # default-initialised variables: the reference's truncated_normal(stddev=0.1) for this function (:33)
    variables.reset_default_store()
    inference.tiny(xd, classes, 5)
    w = variables.default_store().global_variables()["yolo2_tiny/conv3/weights"].cpu().numpy()
    assert np.abs(w).max() <= 0.2 and 0.08 < w.std() < 0.1


def test_tiny_builder_detect_pipeline(cuda):
    """config/yolo2/tiny-20.ini selects `inference = tiny`: Builder -> Model -> device NMS, checked against the oracles."""
    import torch
    from oracle.nms_c import nms_c_batch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import Builder, inference
    from yolo_tf_b200.utils.postprocess import non_max_suppress_device
    classes = 20
    table = tiny_layer_table(classes, 5)
    params = init_params(classes, 5, seed=13, table=table)
    _setup_store(params)
    x = np.random.RandomState(14).normal(0, 1, size=(2, 128, 128, 3)).astype(np.float32)
    builder = Builder.from_values([str(i) for i in range(classes)], 128, 128, ho.ANCHORS_VOC, inference_name="tiny")
    builder(torch.from_numpy(x).to(cuda))
    ref = ho.decode_oracle(tiny_oracle(x, params, classes, 5), classes, ho.ANCHORS_VOC)
    for k in ("conf", "xy_min", "xy_max"):
        got = getattr(builder.model, k).cpu().numpy()
        assert np.abs(got - ref[k]).max() <= 1e-4 * max(1.0, np.abs(ref[k]).max()), k
    conf = builder.model.conf.reshape(2, -1, classes).contiguous()
    lo = builder.model.xy_min.reshape(2, -1, 2).contiguous()
    hi = builder.model.xy_max.reshape(2, -1, 2).contiguous()
    c_ref = conf.cpu().numpy().copy()
    non_max_suppress_device(conf, lo, hi, 0.02, 0.4)
    nms_c_batch(c_ref, lo.cpu().numpy(), hi.cpu().numpy(), 0.02, 0.4)
    assert np.array_equal(conf.cpu().numpy().view(np.uint32), c_ref.view(np.uint32))
    assert inference.TINY_DOWNSAMPLING == (32, 32) and inference._TINY_DOWNSAMPLING == (32, 32)


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("size,batch,seed", [((96, 64), 3, 31), ((128, 128), 8, 5), ((416, 416), 8, 7)])
def test_tiny_training_step_vs_autograd_oracle(cuda, size, batch, seed):
    """SURVEY 8(f) row 4, the training half: `tiny(training=True)` -> Model -> Objectives -> backward through the reference-shaped
    API against oracle/train_oracle.py on tiny's layer table (pinned to one step of the reference's own tiny() source,
    tests/golden/train_reference.npz): batch-statistics BN over 9 convs, the stride-1 SAME max-pool and its overlapping-window
    gradient, conv0 / conv1 stored with 32 channels but trained as the 16-channel variables they are.
    forward / objectives / d(total)/d(net): 1e-4; variable gradients: float32's own accuracy at 416 x 416 (as for Darknet-19),
    direction only at the tiny sizes (tests_gpu_train_helpers.py)."""
    import torch
    from oracle.train_oracle import train_step_oracle
    from tests_gpu_train_helpers import check_gradients_like_float32, check_gradients_sane
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import Builder
    classes = 20
    h, w = size
    table = tiny_layer_table(classes, 5)
    params = init_params(classes, 5, seed=seed, table=table)
    store = _setup_store(params)
    x = np.random.RandomState(seed + 10).normal(0, 1, size=(batch, h, w, 3)).astype(np.float32)
    labels = ho.synthetic_labels(batch, classes, w // 32, h // 32, seed=seed)
    builder = Builder.from_values([str(i) for i in range(classes)], w, h, ho.ANCHORS_VOC, inference_name="tiny")
    builder(torch.from_numpy(x).to(cuda), training=True)
    builder.create_objectives(labels)
    flat, grads = builder.backward(allreduce=False)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    dev = "cuda" if h >= 416 else "cpu"
    ref = train_step_oracle(x, params, classes, ho.ANCHORS_VOC, labels, ho.HPARAM_DEFAULT, table=table, device=dev)
    f32 = train_step_oracle(x, params, classes, ho.ANCHORS_VOC, labels, ho.HPARAM_DEFAULT, dtype=torch.float32, table=table, device=dev)
    net_err, dnet_err = _rel(builder.output.cpu().numpy(), ref["net"]), _rel(builder.objectives.grad_inputs.cpu().numpy(), ref["dnet"])
    print("tiny %s B=%d: forward %.2e (fp32 oracle %.2e), dnet %.2e (fp32 %.2e)" % (size, batch, net_err, _rel(f32["net"], ref["net"]), dnet_err,
                                                                                 _rel(f32["dnet"], ref["dnet"])))
    assert net_err <= 1e-4 and dnet_err <= 1e-4
    for k, v in ref["objectives"].items():
        assert abs(float(builder.objectives[k]) - v) <= 1e-4 * max(abs(v), 1e-9), k
    assert set(grads) == {"yolo2_tiny/" + k for k in ref["grads"]} and flat.numel() == sum(v.size for v in ref["grads"].values())
    assert tuple(grads["yolo2_tiny/conv0/weights"].shape) == (3, 3, 3, 16) and tuple(grads["yolo2_tiny/conv1/weights"].shape) == (3, 3, 16, 32)
    if h >= 416:
        check_gradients_like_float32(grads, ref, f32, "yolo2_tiny/", factor=4.0)
    else:
        check_gradients_sane(grads, ref, "yolo2_tiny/")
    for name, v in ref["new_moving"].items():          # slim UPDATE_OPS
        got = store.global_variables()["yolo2_tiny/" + name].cpu().numpy()
        assert np.abs(got - v).max() <= 1e-5 * max(1.0, np.abs(v).max()), name


def test_tiny_train_op_moves_the_variables(cuda):
    """create_train_op on the tiny graph: one Adam step changes every variable and the next forward uses the new weights."""
    import torch
    from yolo_tf_b200 import variables
    from yolo_tf_b200.model.yolo2 import Builder
    from yolo_tf_b200.optimizer import AdamOptimizer, create_train_op
    classes = 20
    params = init_params(classes, 5, seed=3, table=tiny_layer_table(classes, 5))
    store = _setup_store(params)
    x = torch.from_numpy(np.random.RandomState(4).normal(0, 1, size=(4, 128, 128, 3)).astype(np.float32)).to(cuda)
    labels = ho.synthetic_labels(4, classes, 4, 4, seed=4)
    builder = Builder.from_values([str(i) for i in range(classes)], 128, 128, ho.ANCHORS_VOC, inference_name="tiny")
    op = create_train_op(builder, AdamOptimizer(1e-3), clip_gradient_norm=1.0)
    before = {k: v.clone() for k, v in store.global_variables().items()}
    l0 = float(op(x, labels))
    l1 = float(op(x, labels))
    torch.cuda.synchronize()
    after = store.global_variables()
    assert np.isfinite(l0) and np.isfinite(l1) and l1 != l0
    moved = [k for k in before if "moving" not in k and not torch.equal(before[k], after[k])]
    assert len(moved) == 26, sorted(set(before) - set(moved))
    assert tuple(after["yolo2_tiny/conv0/weights"].shape) == (3, 3, 3, 16)
