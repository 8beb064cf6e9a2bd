"""Per-layer parity of the forward plan at the bench configuration (B=32, 416x416, C=80) against the CPU oracle on the
first images, for the plan options (fuse_pool / halo / pair).  Diagnostic tool.
    python tools/diag_layers.py [batch] [size] [classes] [nref] [kcap]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.darknet_oracle import darknet_oracle, init_params, layer_table  # noqa: E402
from yolo_tf_b200 import _lib, variables  # noqa: E402
from yolo_tf_b200.model.yolo2 import inference  # noqa: E402


def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
    classes = int(sys.argv[3]) if len(sys.argv) > 3 else 80
    nref = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    params = init_params(classes, 5, seed=1)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    x = np.random.RandomState(2).normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
    taps = {}
    ref = darknet_oracle(x[:nref], params, classes, 5, taps=taps, dtype=torch.float64)
    inference._Engine.KEEP_ACTIVATIONS = True          # one workspace slot per layer: every tap stays readable after the forward
    eng = inference._Engine.get(torch.device("cuda:0"), classes, 5)
    xd = torch.from_numpy(x).cuda()
    L = _lib.lib()
    import ctypes
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    if len(sys.argv) > 5:
        L.y2_debug_set(8, float(sys.argv[5]))                 # accumulation-chain cap (k-blocks), 0 = unlimited
    for opts in ({"fuse_pool": 0, "halo": 0, "pair": 0}, {"fuse_pool": 0, "halo": 1, "pair": 0}, {"fuse_pool": 1, "halo": 1, "pair": 0},
                 {"fuse_pool": 1, "halo": 1, "pair": 1}):
        for k, v in opts.items():
            _lib.check(L.y2_set_option(eng.h, k.encode(), v))
        _, out = inference.darknet(xd, classes, 5)
        torch.cuda.synchronize()
        _lib.check(L.y2_check_async_errors())
        errs = []
        for i, (name, k, cin, cout, then) in enumerate(layer_table(classes, 5)[:-1]):
            has_pool = then in ("pool", "passthrough+pool")
            key = name + "/pool" if has_pool else name
            shape = (batch,) + taps[key].shape[1:]
            try:
                got = eng.activation(i, has_pool, shape)[:nref].cpu().numpy()
                errs.append("%s %.1e" % (key, rel(got, taps[key])))
            except _lib.Y2Error:
                errs.append("%s n/a" % key)
        print(opts, "output %.2e" % rel(out[:nref].cpu().numpy(), ref))
        print("   ", "  ".join(errs), flush=True)
    for k, v in {"fuse_pool": 1, "halo": 1, "pair": 1}.items():
        _lib.check(L.y2_set_option(eng.h, k.encode(), v))


if __name__ == "__main__":
    main()
