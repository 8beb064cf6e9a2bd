"""GPU bring-up ladder for the tcgen05 conv kernel: runs y2_conv2d on progressively harder cases,
compares with torch fp32 conv2d (TF32 off) and writes gpurun_out/probe_conv.json.  Diagnostic tool,
not part of the product path."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tf_b200 import _lib  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def ref_conv(x, w, scale, bias, leaky):
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(3, 2, 0, 1).double(), padding=w.shape[0] // 2)
    y = y.permute(0, 2, 3, 1)
    if scale is not None:
        y = y * scale.double()
    if bias is not None:
        y = y + bias.double()
    if leaky:
        y = torch.maximum(y, 0.1 * y)
    return y


def run_case(name, B, H, W, cin, k, cout, precision=0, block_n=0, max_ctas=0, pattern="random", leaky=1, affine=True,
             dump=False):
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 1000)
    x = torch.randn(B, H, W, cin, device="cuda", generator=g)
    if pattern == "identity":
        w = torch.zeros(k, k, cin, cout, device="cuda")
        for n in range(cout):
            w[k // 2, k // 2, n % cin, n] = 1.0
    else:
        w = torch.randn(k, k, cin, cout, device="cuda", generator=g) / np.sqrt(k * k * cin)
    scale = (torch.rand(cout, device="cuda", generator=g) + 0.5) if affine else None
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1 if affine else None
    y = torch.full((B, H, W, cout), float("nan"), device="cuda")
    t0 = time.time()
    rc = L.y2_conv2d(_lib.ptr(x), B, H, W, cin, _lib.ptr(w), k, cout, _lib.ptr(scale), _lib.ptr(bias), leaky,
                     _lib.ptr(y), precision, block_n, max_ctas, None)
    torch.cuda.synchronize()
    res = {"name": name, "shape": [B, H, W, cin, k, cout], "precision": precision, "block_n": block_n,
           "max_ctas": max_ctas, "rc": rc, "secs": round(time.time() - t0, 3)}
    if rc != 0:
        res["error"] = L.y2_last_error().decode()
        return res
    ref = ref_conv(x, w, scale, bias, leaky)
    err = (y.double() - ref).abs()
    res["nan"] = int(torch.isnan(y).sum())
    res["max_abs_err"] = float(torch.nan_to_num(err, nan=1e30).max())
    res["ref_max"] = float(ref.abs().max())
    res["rel_inf"] = res["max_abs_err"] / max(res["ref_max"], 1e-30)
    res["rel_l2"] = float(torch.nan_to_num(err, nan=0.0).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt())
    if dump or res["rel_inf"] > 1e-2:
        yy = y.reshape(-1, cout)
        rr = ref.reshape(-1, cout)
        res["y_head"] = yy[:4, :8].tolist()
        res["ref_head"] = rr[:4, :8].float().tolist()
        bad_rows = (torch.nan_to_num(err, nan=1e30).reshape(-1, cout).max(1).values > 1e-2 * res["ref_max"]).nonzero().flatten()
        res["bad_rows"] = int(bad_rows.numel())
        res["bad_rows_first"] = bad_rows[:16].tolist()
        bad_cols = (torch.nan_to_num(err, nan=1e30).reshape(-1, cout).max(0).values > 1e-2 * res["ref_max"]).nonzero().flatten()
        res["bad_cols"] = int(bad_cols.numel())
        res["bad_cols_first"] = bad_cols[:16].tolist()
    return res


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    cases = [
        dict(name="a_1x1_identity_bf16", B=1, H=8, W=16, cin=64, k=1, cout=32, precision=1, pattern="identity", leaky=0, affine=False, dump=True),
        dict(name="b_1x1_random_bf16", B=1, H=8, W=16, cin=64, k=1, cout=32, precision=1, leaky=0, affine=False),
        dict(name="c_1x1_random_split3", B=1, H=8, W=16, cin=64, k=1, cout=32, precision=0, leaky=0, affine=False),
        dict(name="d_1x1_k128", B=1, H=8, W=16, cin=128, k=1, cout=64, precision=0),
        dict(name="e_1x1_k1024_n256", B=2, H=16, W=16, cin=1024, k=1, cout=256, precision=0),
        dict(name="f_3x3_identity", B=1, H=8, W=16, cin=64, k=3, cout=64, precision=1, pattern="identity", leaky=0, affine=False, dump=True),
        dict(name="g_3x3_c64", B=1, H=8, W=16, cin=64, k=3, cout=64, precision=0),
        dict(name="h_3x3_c64_multi_image", B=3, H=13, W=13, cin=64, k=3, cout=128, precision=0),
        dict(name="i_3x3_c32_sw64", B=2, H=16, W=16, cin=32, k=3, cout=64, precision=0),
        dict(name="j_1x1_n425_tail", B=3, H=13, W=13, cin=1024, k=1, cout=425, precision=0, leaky=0),
        dict(name="k_3x3_splitk4", B=2, H=13, W=13, cin=512, k=3, cout=256, precision=0, max_ctas=4),
        dict(name="l_3x3_n512_two_ntiles", B=2, H=26, W=26, cin=256, k=3, cout=512, precision=0),
        dict(name="m_3x3_bn128", B=2, H=26, W=26, cin=128, k=3, cout=256, precision=0, block_n=128),
        dict(name="n_conv1_shape", B=2, H=208, W=208, cin=32, k=3, cout=64, precision=0),
        dict(name="o_conv13_shape", B=8, H=13, W=13, cin=512, k=3, cout=1024, precision=0),
        dict(name="p_conv20_shape", B=4, H=13, W=13, cin=3072, k=3, cout=1024, precision=0),
        dict(name="q_conv13_bf16", B=8, H=13, W=13, cin=512, k=3, cout=1024, precision=1),
        dict(name="r_608_conv18", B=2, H=19, W=19, cin=1024, k=3, cout=1024, precision=0),
    ]
    out = []
    for c in cases:
        try:
            r = run_case(**c)
        except Exception as e:  # keep going: one bad case must not hide the others
            r = {"name": c["name"], "exception": repr(e)}
        out.append(r)
        brief = {k: r.get(k) for k in ("name", "rc", "rel_inf", "rel_l2", "nan", "error", "exception", "secs")}
        print(json.dumps(brief), flush=True)
        with open("gpurun_out/probe_conv.json", "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
