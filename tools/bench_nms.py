"""BASELINE config 5: NMS throughput sweep on one B200 -- grids 13x13x5 (N=845) and 19x19x5 (N=1805),
K in {100..10k} (box,class) candidates above 0.3 per image, batch 512, C=80, thresholds 0.3/0.4.
Reports input boxes/s (B*N/t), candidates/s (B*K/t), algorithmic GB/s (B*(2*4*N*C + 16*N)/t) against the
measured HBM peak, beside the CPU oracle (C twin on 2 images, reference-shaped Python on 1 image, bounded).
Writes JSON lines to stdout and gpurun_out/nms_sweep_<tag>.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_tf_b200 import _lib  # noqa: E402

ANCH = np.array([[0.738768, 0.874946], [2.42204, 2.65704], [4.30971, 7.04493], [10.246, 4.59428], [12.6868, 11.8741]])


def make_inputs(rs, B, g, C, K):
    A, cells = 5, g * g
    gy, gx = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    centre = np.stack([gx, gy], -1).reshape(1, cells, 1, 2) + rs.uniform(0, 1, size=(B, cells, A, 2))
    wh = ANCH.reshape(1, 1, A, 2) * np.exp(rs.normal(0, 0.5, size=(B, cells, A, 2)))
    lo = (centre - wh / 2).astype(np.float32).reshape(B, cells * A, 2)
    hi = (centre + wh / 2).astype(np.float32).reshape(B, cells * A, 2)
    N = cells * A
    conf = rs.uniform(0, 0.29, size=(B, N * C)).astype(np.float32)
    for b in range(B):
        pick = rs.choice(N * C, size=K, replace=False)
        conf[b, pick] = rs.uniform(0.3, 1.0, size=K)
    return conf.reshape(B, N, C), lo, hi


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    L = _lib.lib()
    B, C = 512, 80
    out = []
    for g in (13, 19):
        for K in (100, 300, 1000, 3000, 10000):
            rs = np.random.RandomState(5)
            conf, lo, hi = make_inputs(rs, B, g, C, K)
            N = conf.shape[1]
            d0, dlo, dhi = torch.from_numpy(conf).cuda(), torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
            work = torch.empty_like(d0)
            nbytes = L.y2_nms_workspace_bytes(B, N, C)
            ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            times = []
            for rep in range(8):
                work.copy_(d0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(L.y2_nms(_lib.ptr(work), _lib.ptr(dlo), _lib.ptr(dhi), B, N, C, 0.3, 0.4, None, None, _lib.ptr(ws), nbytes, None))
                e1.record()
                torch.cuda.synchronize()
                if rep >= 3:
                    times.append(e0.elapsed_time(e1))
            ms = float(np.mean(times))
            alg_bytes = B * (2 * 4 * N * C + 16 * N)
            # CPU side, bounded
            from oracle.nms_c import nms_c_batch
            from oracle.nms_oracle import nms_oracle
            c2 = conf[:2].copy()
            t0 = time.perf_counter()
            nms_c_batch(c2, lo[:2], hi[:2], 0.3, 0.4)
            t_c = (time.perf_counter() - t0) / 2
            assert np.array_equal(work[:2].cpu().numpy().view(np.uint32), c2.view(np.uint32)), "GPU NMS differs from the oracle"
            t_py = None
            if K <= 1000:
                c1 = conf[0].copy().reshape(g * g, 5, C)
                t0 = time.perf_counter()
                nms_oracle(c1, lo[0].reshape(g * g, 5, 2), hi[0].reshape(g * g, 5, 2), 0.3, 0.4)
                t_py = time.perf_counter() - t0
            rec = {"grid": "%dx%dx5" % (g, g), "N": N, "K": K, "B": B, "C": C, "gpu_ms": ms,
                   "boxes_per_s": B * N / (ms / 1e3), "candidates_per_s": B * K / (ms / 1e3),
                   "algorithmic_GBs": alg_bytes / (ms / 1e3) / 1e9, "hbm_peak_GBs": peaks["hbm_gbs"],
                   "frac_of_hbm_peak": alg_bytes / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                   "cpu_c_oracle_s_per_image": t_c, "cpu_python_oracle_s_per_image": t_py,
                   "speedup_vs_c_oracle_1thread": t_c * B / (ms / 1e3), "bit_exact_vs_oracle": True}
            out.append(rec)
            print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/nms_sweep_%s.json" % tag, "w"), indent=1)


if __name__ == "__main__":
    main()
