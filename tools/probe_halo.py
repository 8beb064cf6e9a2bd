"""Halo-tile experiment for the 32-channel 3x3 layer (conv1): does tcgen05.mma accept an A operand that is a shifted
window of a larger SWIZZLE_64B tile (descriptor start not aligned to the swizzle atom, 8-row groups 10 rows apart)?
Runs y2_conv2d in im2col mode (0), single-halo mode (1, with and without the descriptor base offset) and the aligned
three-copies mode (2), checks each against torch fp64 and times them; then splits conv1's time into load / MMA /
epilogue shares with the diagnostic skip flags.  Diagnostic tool, not part of the product path.
Writes gpurun_out/probe_halo.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402
from tools.probe_conv import ref_conv  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run(L, B, H, W, cout, halo, flags, check=True, precision=0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, H, W, 32, device="cuda", generator=g)
    w = torch.randn(3, 3, 32, cout, device="cuda", generator=g) / 17.0
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1
    y = torch.full((B, H, W, cout), float("nan"), device="cuda")
    L.y2_debug_set(4, float(halo))
    L.y2_debug_set(3, float(flags))
    rc = L.y2_conv2d(_lib.ptr(x), B, H, W, 32, _lib.ptr(w), 3, cout, _lib.ptr(scale), _lib.ptr(bias), 1, _lib.ptr(y),
                     precision, 0, 0, None)
    torch.cuda.synchronize()
    L.y2_debug_set(4, 0.0)
    L.y2_debug_set(3, 0.0)
    res = {"shape": [B, H, W, 32, 3, cout], "halo": halo, "flags": flags, "precision": precision, "rc": rc}
    if rc != 0:
        res["error"] = L.y2_last_error().decode()
        return res
    res["ms"] = float(L.y2_debug_last_conv_ms())
    if check:
        ref = ref_conv(x, w, scale, bias, 1)
        err = (y.double() - ref).abs()
        res["nan"] = int(torch.isnan(y).sum())
        res["rel_inf"] = float(torch.nan_to_num(err, nan=1e30).max() / ref.abs().max())
        # where are the errors: interior / border
        bad = torch.nan_to_num(err, nan=1e30).amax(dim=(0, 3)) > 1e-3 * float(ref.abs().max())
        res["bad_pixels"] = int(bad.sum())
        if 0 < res["bad_pixels"] <= 400000:
            ys, xs = torch.nonzero(bad, as_tuple=True)
            res["bad_x_mod8"] = sorted(set((xs % 8).tolist()))
            res["bad_y_mod16"] = sorted(set((ys % 16).tolist()))
    return res


def net_level(L):
    """Whole Darknet-19 forward (B=32, 416, C=80) with the halo option 0/1/2: output agreement and ms per forward."""
    import numpy as np
    import bench
    from yolo_tf_b200 import variables
    from yolo_tf_b200.model.yolo2 import inference
    C, B, size = 80, 32, 416
    params = bench.synthetic_checkpoint(C, 5)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    x = torch.from_numpy(np.random.RandomState(3).normal(0, 1, size=(B, size, size, 3)).astype(np.float32)).cuda()
    eng = inference._Engine.get(torch.device("cuda:0"), C, 5)
    res, base = [], None
    for halo in (0, 1, 0, 1):
        _lib.check(L.y2_set_option(eng.h, b"halo", halo))
        for _ in range(3):
            _, o = inference.darknet(x, C, 5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            _, o = inference.darknet(x, C, 5)
        e1.record()
        torch.cuda.synchronize()
        rc = L.y2_check_async_errors()
        r = {"what": "darknet forward", "halo": halo, "ms": e0.elapsed_time(e1) / 30, "async_rc": rc}
        if rc != 0:
            r["error"] = L.y2_last_error().decode()
        o = o.clone()
        if base is None:
            base = o
        r["rel_vs_halo0"] = float((o - base).abs().max() / base.abs().max())
        res.append(r)
        print(json.dumps(r), flush=True)
    _lib.check(L.y2_set_option(eng.h, b"halo", 0))
    return res


def main():
    L = _lib.lib()
    import ctypes
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    L.y2_debug_last_conv_ms.restype = ctypes.c_float
    out = []
    # correctness ladder
    for (B, H, W, cout) in [(1, 16, 8, 64), (3, 48, 40, 32), (4, 208, 208, 64)]:
        for halo, flags in [(0, 0), (1, 0)]:
            for prec in (0, 1):
                r = run(L, B, H, W, cout, halo, flags, precision=prec)
                out.append(r)
                print(json.dumps(r), flush=True)
                if L.y2_check_async_errors() != 0:
                    print("async error:", L.y2_last_error().decode(), flush=True)
    # timing at the bench shape (conv1: B=32, 208x208, 32 -> 64)
    for halo, flags in [(0, 0), (0, 1), (0, 3), (1, 0), (1, 1), (1, 3)]:
        r = run(L, 32, 208, 208, 64, halo, flags, check=False)
        r["what"] = "timing conv1 shape"
        out.append(r)
        print(json.dumps(r), flush=True)
        L.y2_check_async_errors()
    out.extend(net_level(L))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_halo.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
