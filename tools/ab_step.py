"""Same-box A/B of kernel options on the bench step (Darknet-19 forward + decode + NMS, B=32, 416, C=80): configurations
are interleaved after a thermal warm-up so the power-cap clock drift hits all of them alike.  Diagnostic tool.
Usage: python tools/ab_step.py "name:key=val,key=val" ...     keys = y2_debug_set ids (5 = PDL, 6 = TMA-store epilogue),
       or halo=<v> / fuse_pool=<v> / pair=<v> (y2_set_option).  Writes gpurun_out/ab_step.json."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from yolo_tf_b200 import _lib, variables  # noqa: E402
from yolo_tf_b200.model.yolo2 import Builder, inference  # noqa: E402
from yolo_tf_b200.utils.postprocess import non_max_suppress_device  # noqa: E402


def main():
    cfgs = []
    for a in sys.argv[1:]:
        name, _, kv = a.partition(":")
        cfgs.append((name, [tuple(x.split("=")) for x in kv.split(",") if x]))
    if not cfgs:
        cfgs = [("base", [])]
    L = _lib.lib()
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    B, size, C = 32, 416, 80
    params = bench.synthetic_checkpoint(C, 5)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    builder = Builder.from_values([str(i) for i in range(C)], size, size, bench.ANCHORS_COCO)
    rs = np.random.RandomState(100)
    xs = [torch.from_numpy(rs.normal(0, 1, size=(B, size, size, 3)).astype(np.float32)).cuda() for _ in range(4)]
    N = (size // 32) ** 2 * 5

    def step(x):
        builder(x)
        m = builder.model
        non_max_suppress_device(m.conf.view(B, N, C), m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2), 0.3, 0.4, check=False)

    eng = inference._Engine.get(torch.device("cuda:0"), C, 5)

    def apply(kvs):
        for k, v in [("5", "1"), ("6", "1"), ("3", "0"), ("8", "32")]:
            L.y2_debug_set(int(k), float(v))
        opts = {"halo": 1, "fuse_pool": 1, "pair": 1, "conv0_tc": 2}
        for k, v in kvs:
            if k in opts:
                opts[k] = int(v)
            else:
                L.y2_debug_set(int(k), float(v))
        for k, v in opts.items():
            _lib.check(L.y2_set_option(eng.h, k.encode(), v))      # also invalidates the cached plan

    for i in range(600):                                          # ~2 s: reach the power-capped steady state
        step(xs[i % 4])
    torch.cuda.synchronize()
    res = {name: [] for name, _ in cfgs}
    for rep in range(4):
        for name, kvs in cfgs:
            apply(kvs)
            for i in range(10):
                step(xs[i % 4])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(100):
                step(xs[i % 4])
            e1.record()
            torch.cuda.synchronize()
            _lib.check(L.y2_check_async_errors())
            res[name].append(e0.elapsed_time(e1) / 100)
    out = {name: {"ms_per_step": v, "median": float(np.median(v))} for name, v in res.items()}
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/ab_step.json", "w"), indent=1)
    apply([])


if __name__ == "__main__":
    main()
