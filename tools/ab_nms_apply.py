"""Same-box A/B of nms_apply_kernel's work-item schemes (y2_debug_set key 11) at the BASELINE configs[4] sweep points and at the
detection batch (B = 32, skewed class distribution taken from the bench workload's decode output is not available here, so a
synthetic skew: 40 % of an image's candidates in one class).  Prints JSON lines."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hbm_kernels import nms_inputs  # noqa: E402
from yolo_tf_b200 import _lib  # noqa: E402


def main():
    L = _lib.lib()
    L.y2_debug_set.argtypes = [ctypes.c_int, ctypes.c_double]
    for (B, g, K, skew) in ((512, 13, 100, 0), (512, 13, 1000, 0), (512, 13, 10000, 0), (512, 19, 10000, 0), (32, 13, 900, 1)):
        rs = np.random.RandomState(5)
        conf, lo, hi = nms_inputs(rs, B, g, 80, K)
        if skew:                                   # move ~40 % of the candidates into class 7
            N = conf.shape[1]
            for b in range(B):
                idx = rs.choice(N, size=int(0.4 * K), replace=False)
                conf[b, idx, 7] = rs.uniform(0.3, 1.0, size=idx.size).astype(np.float32)
        d0, dlo, dhi = torch.from_numpy(conf).cuda(), torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
        work = torch.empty_like(d0)
        N = conf.shape[1]
        nbytes = L.y2_nms_workspace_bytes(B, N, 80)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        row = {"B": B, "N": N, "K": K, "skew": skew}
        ref = None
        for rep_round in range(2):                 # two interleaved rounds
            for mode in (0, 1, 2):
                L.y2_debug_set(11, float(mode))
                times = []
                for rep in range(8):
                    work.copy_(d0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.check(L.y2_nms(_lib.ptr(work), _lib.ptr(dlo), _lib.ptr(dhi), B, N, 80, 0.3, 0.4, None, None, _lib.ptr(ws), nbytes, None))
                    e1.record()
                    torch.cuda.synchronize()
                    if rep >= 3:
                        times.append(e0.elapsed_time(e1))
                if ref is None:
                    ref = work.clone()
                assert torch.equal(work, ref), "modes disagree"
                row.setdefault("mode%d_ms" % mode, []).append(round(float(np.median(times)), 4))
        L.y2_debug_set(11, 0.0)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
