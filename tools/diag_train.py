"""Training-step parity diagnostic at any size, incl. BASELINE config 3 itself (B = 64, 416 x 416, 20 classes): our step
vs the float64 autograd oracle evaluated by torch on the device (same restatement, oracle/train_oracle.py), with the
float32 evaluation of the same oracle as the floor.  Prints one JSON object.
    python tools/diag_train.py [batch] [size] [classes] [seed]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import head_oracle as ho  # noqa: E402
from oracle.darknet_oracle import init_params, layer_table  # noqa: E402
from oracle.train_oracle import train_step_oracle  # noqa: E402
from yolo_tf_b200 import _lib, variables  # noqa: E402
from yolo_tf_b200.model.yolo2 import Builder, inference  # noqa: E402


def rel(a, b):
    a = torch.as_tensor(a).double().cuda()
    b = torch.as_tensor(b).double().cuda()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cuda()
    b = torch.as_tensor(b).double().cuda()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
    classes = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if os.environ.get("Y2_KCAP"):                                  # accumulation-chain cap in k-blocks (default 32)
        import ctypes
        _lib.lib().y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
        _lib.lib().y2_debug_set(8, float(os.environ["Y2_KCAP"]))
    anchors = ho.ANCHORS_VOC if classes == 20 else ho.ANCHORS_COCO
    params = init_params(classes, 5, seed=seed)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    x = np.random.RandomState(seed + 10).normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    cw = size // 32
    labels = ho.synthetic_labels(batch, classes, cw, cw, seed=seed)
    builder = Builder.from_values([str(i) for i in range(classes)], size, size, anchors)
    eng0 = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    for key in ("train_f16", "train_kcap"):                       # training-forward numerics (defaults: fp16 planes, 8 k-blocks)
        if os.environ.get("Y2_" + key.upper()):
            _lib.check(_lib.lib().y2_set_option(eng0.h, key.encode(), int(os.environ["Y2_" + key.upper()])))
    builder(torch.from_numpy(x).cuda(), training=True)
    builder.create_objectives(labels)
    flat, grads = builder.backward(allreduce=False)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    L = _lib.lib()
    taps64, taps32 = {}, {}
    ref = train_step_oracle(x, params, classes, anchors, labels, ho.HPARAM_DEFAULT, taps=taps64, device="cuda", taps_numpy=False)
    f32 = train_step_oracle(x, params, classes, anchors, labels, ho.HPARAM_DEFAULT, dtype=torch.float32, taps=taps32, device="cuda", taps_numpy=False)
    out = {"batch": batch, "size": size, "classes": classes, "seed": seed,
           "train_f16": os.environ.get("Y2_TRAIN_F16", "default"), "train_kcap": os.environ.get("Y2_TRAIN_KCAP", "default"),
           "net": [rel(builder.output, ref["net"]), rel(f32["net"], ref["net"])],
           "dnet": [rel(builder.objectives.grad_inputs, ref["dnet"]), rel(f32["dnet"], ref["dnet"])],
           "objectives": {k: [abs(float(builder.objectives[k]) - v) / max(abs(v), 1e-30), abs(f32["objectives"][k] - v) / max(abs(v), 1e-30)]
                          for k, v in ref["objectives"].items()}}
    layers = {}
    for i, (name, k, cin, cout, then) in enumerate(layer_table(classes, 5)[:-1]):
        t64, t32 = taps64[name], taps32[name]
        kind = 1
        if then == "pool":                                   # only the pooled tensor is materialised in training
            kind = 2
            t64 = torch.nn.functional.max_pool2d(t64.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
            t32 = torch.nn.functional.max_pool2d(t32.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        got = torch.empty(tuple(t64.shape), device="cuda")
        _lib.check(L.y2_train_get_tensor(eng.h, kind, i, _lib.ptr(got), None))     # y = leaky(BN(z))
        layers[name] = [rel(got, t64.contiguous()), rel(t32.contiguous(), t64.contiguous())]
        del got
    out["layers_y"] = layers
    g = {}
    for name, g_ref in ref["grads"].items():
        g[name] = [rel(grads["yolo2_darknet/" + name], g_ref), rel(f32["grads"][name], g_ref),
                   rel_l2(grads["yolo2_darknet/" + name], g_ref), rel_l2(f32["grads"][name], g_ref)]
    out["grads"] = g
    out["worst_grad"] = max(v[0] for v in g.values())
    out["worst_grad_fp32_floor"] = max(v[1] for v in g.values())
    out["worst_grad_l2"] = max(v[2] for v in g.values())
    out["worst_grad_l2_fp32_floor"] = max(v[3] for v in g.values())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
