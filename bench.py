#!/usr/bin/env python
"""bench.py -- images/sec of the YOLOv2-Darknet19 detection hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]                      (the B200-native arm)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]     (the reference's CPU path)
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads:
  --workload infer (default; BASELINE.json configs[1], and configs[3] with --size 608): YOLOv2-Darknet19, 80 classes,
      batch 32 per GPU.  One "step" = one batch through the detection pipeline detect.py:59-87 drives:
      uint8 image -> per_image_standardization -> darknet backbone -> head decode -> NMS -> detection list.
      The synthetic checkpoint's final layer is scaled so that ~900 (box, class) scores per image exceed the 0.3
      threshold: the NMS does real work (round 1's weights produced no candidate at all).
  --workload train (configs[2]): 20 classes, batch 64 per GPU, forward with batch statistics + 4-part loss + backward
      (+ ONE all-reduce of the flat gradient bucket when N > 1) + Adam.
Images are independent, so N GPUs = N shards (weak scaling; no data-path collective in inference).

Prints ONE JSON line (rank 0): value = device-resident throughput; e2e = through the public API from pinned host buffers with
H2D/D2H inside the timed region; roofline = the tcgen05 conv kernels' algorithmic FLOP/s vs the measured bf16 peak (burst or
sustained chosen by the length of the timed region, both reported); nms = BASELINE configs[4] sweep points (boxes/s,
candidates/s, algorithmic GB/s); cpu_baseline = the CPU oracle on a bounded sample.
"""
import argparse
import ctypes
import glob
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ANCHORS_COCO = [[0.738768, 0.874946], [2.42204, 2.65704], [4.30971, 7.04493], [10.246, 4.59428], [12.6868, 11.8741]]
ANCHORS_VOC = [[1.08, 1.19], [3.42, 4.41], [6.63, 11.38], [9.42, 5.11], [16.62, 10.52]]
HPARAM = {"prob": 1.0, "iou_best": 5.0, "iou_normal": 1.0, "coords": 1.0}      # config.ini:98-102
THRESHOLD, THRESHOLD_IOU = 0.3, 0.4          # detect.py:127-128
CLASS_LOGIT_GAIN, IOU_LOGIT_BIAS = 16.0, 2.0


def synthetic_checkpoint(classes, num_anchors, seed=1, dense_detections=True):
    """Conditioned random weights (He-style, non-trivial BN statistics) under the TF variable names.
    dense_detections: the final (linear) layer's class-logit columns are multiplied by CLASS_LOGIT_GAIN and its objectness
    bias raised by IOU_LOGIT_BIAS, so that the softmax is peaked and ~900 of the 67 600 (box, class) scores of a 416 x 416
    image exceed the 0.3 detection threshold (calibrated with the CPU oracle) -- a busy-scene NMS load instead of none."""
    from yolo_tf_b200.model.yolo2.inference import layer_geometry
    rs = np.random.RandomState(seed)
    p = {}
    for name, k, cin, cout, has_bn, _ in layer_geometry(classes, num_anchors):
        std = math.sqrt(2.0 / (1.01 * k * k * cin)) * (1.0 if has_bn else 0.25)
        p[name + "/weights"] = rs.normal(0.0, std, size=(k, k, cin, cout)).astype(np.float32)
        if has_bn:
            p[name + "/BatchNorm/gamma"] = rs.uniform(0.7, 1.2, size=cout).astype(np.float32)
            p[name + "/BatchNorm/beta"] = rs.normal(0, 0.1, size=cout).astype(np.float32)
            p[name + "/BatchNorm/moving_mean"] = rs.normal(0, 0.1, size=cout).astype(np.float32)
            p[name + "/BatchNorm/moving_variance"] = rs.uniform(0.8, 1.3, size=cout).astype(np.float32)
        else:
            p[name + "/biases"] = rs.normal(0, 0.1, size=cout).astype(np.float32)
            if dense_detections:
                w = p[name + "/weights"].reshape(k, k, cin, num_anchors, 5 + classes)
                w[..., 5:] *= CLASS_LOGIT_GAIN
                b = p[name + "/biases"].reshape(num_anchors, 5 + classes)
                b[:, 5:] *= CLASS_LOGIT_GAIN
                b[:, 0] += IOU_LOGIT_BIAS
    return p


def synthetic_images_u8(rs, batch, size):
    """What detect.py:59-65 hands to the preprocessing: a resized uint8 RGB image."""
    return rs.randint(0, 256, size=(batch, size, size, 3)).astype(np.uint8)


def synthetic_boxes(rs, batch, classes):
    """per image n ~ U{1..8} objects, class U{0..C-1}, centre U(0,1)^2, w, h ~ U(0.05, 0.6)
    clipped to the image.  Returns (list of class arrays, list of [n, 4] (xmin, ymin, xmax, ymax) arrays)."""
    cls, xy = [], []
    for _ in range(batch):
        n = rs.randint(1, 9)
        cls.append(rs.randint(0, classes, size=n))
        cx, cy = rs.uniform(0, 1, size=n), rs.uniform(0, 1, size=n)
        w, h = rs.uniform(0.05, 0.6, size=n), rs.uniform(0.05, 0.6, size=n)
        xy.append(np.stack([np.clip(cx - w / 2, 0, 1 - 1e-6), np.clip(cy - h / 2, 0, 1 - 1e-6),
                            np.clip(cx + w / 2, 0, 1 - 1e-6), np.clip(cy + h / 2, 0, 1 - 1e-6)], 1))
    return cls, xy


def conv_flops(h, w, classes, num_anchors, tensor_core_only=False):
    from yolo_tf_b200.model.yolo2.inference import layer_geometry
    total, hh, ww = 0, h, w
    for i, (name, k, cin, cout, has_bn, pool) in enumerate(layer_geometry(classes, num_anchors)):
        if not (tensor_core_only and i == 0):
            total += 2 * hh * ww * k * k * cin * cout
        if pool:
            hh //= 2
            ww //= 2
    return total


def train_flops_per_image(size, classes):
    """fwd + dgrad + wgrad, no dgrad for conv0 (104.39 GFLOP at 416, C = 20)."""
    return 3 * conv_flops(size, size, classes, 5) - 2 * size * size * 27 * 32


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def pick_peak(peaks, timed_seconds):
    """The burst figure for a timed region that ends before the clocks settle under the power cap (< 2 s), the sustained
    one for a long region.  Returns (peak, which)."""
    burst = float(peaks.get("bf16_tflops", 1590.0))
    sustained = float(peaks.get("bf16_tflops_sustained", burst))
    return (burst, "bf16_tflops (burst)") if timed_seconds < 2.0 else (sustained, "bf16_tflops_sustained")


class ClockSampler(object):
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()                      # exact PID, never a pattern
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, pw = [], [], set(), []
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []        # upper half = samples under load
        return {"sm_mhz": (sorted(busy)[len(busy) // 2] if busy else None), "sm_max_mhz": (max(mx) if mx else None),
                "power_w_max": (max(pw) if pw else None), "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------------
# CPU legs: the ONLY place bench.py touches oracle/
def _best_threads(fn):
    """oneDNN does not scale to every core count on these shapes: pick the fastest of {all, 64, 32, 16, 8} threads."""
    import torch
    cores = os.cpu_count() or 1
    best, threads = None, cores
    for t in sorted({cores, 64, 32, 16, 8}, reverse=True):
        if t > cores:
            continue
        torch.set_num_threads(t)
        fn()                                   # warm
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, threads = dt, t
    torch.set_num_threads(threads)
    return threads


def cpu_infer_images_per_sec(params, classes, size, n_images, steps, warmup, seed=7):
    """The reference's CPU detection path restated (oracle/): numpy standardisation + conv stack on the host threads + decode
    + Python NMS + detection selection.  Returns (images/s, threads, description, ms per step)."""
    from oracle.darknet_oracle import darknet_oracle
    from oracle.head_oracle import decode_oracle
    from oracle.nms_oracle import nms_oracle
    from oracle.prepost_oracle import detections_oracle, per_image_standardization_oracle
    rs = np.random.RandomState(seed)
    u8 = synthetic_images_u8(rs, n_images, size)
    x1 = np.stack([per_image_standardization_oracle(u8[0].astype(np.float32))])
    threads = _best_threads(lambda: darknet_oracle(x1, params, classes, len(ANCHORS_COCO)))
    cw = size // 32
    cands = []

    def one_pass():
        x = np.stack([per_image_standardization_oracle(im.astype(np.float32)) for im in u8]).astype(np.float32)
        net = darknet_oracle(x, params, classes, len(ANCHORS_COCO))
        m = decode_oracle(net, classes, ANCHORS_COCO)
        for b in range(n_images):
            conf = np.ascontiguousarray(m["conf"][b])
            cands.append(int((conf > THRESHOLD).sum()))
            lo, hi = np.ascontiguousarray(m["xy_min"][b]), np.ascontiguousarray(m["xy_max"][b])
            nms_oracle(conf, lo, hi, THRESHOLD, THRESHOLD_IOU)
            detections_oracle(conf.reshape(-1, classes), lo.reshape(-1, 2), hi.reshape(-1, 2), THRESHOLD, (size / cw, size / cw))

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    desc = ("%d step(s) x %d images of the same workload (~%d NMS candidates per image): numpy standardisation + torch-CPU fp32 conv "
            "stack (oneDNN, best of {all,64,32,16,8} = %d threads; TF1 itself is not installable, torch-CPU is the stand-in and is "
            "expected to be faster than TF-1.0 Eigen) + numpy decode + reference-shaped pure-Python NMS (1 thread) + detection selection"
            % (steps, n_images, int(np.mean(cands)) if cands else 0, threads))
    return steps * n_images / dt, threads, desc, dt / steps * 1000.0


def cpu_train_images_per_sec(params, classes, size, n_images, steps, warmup, seed=7):
    """The reference's CPU training step restated (oracle/train_oracle.py, float32): forward with batch statistics + loss +
    autograd backward on the host threads."""
    import torch
    from oracle.head_oracle import synthetic_labels
    from oracle.train_oracle import train_step_oracle
    rs = np.random.RandomState(seed)
    x = rs.normal(0, 1, size=(n_images, size, size, 3)).astype(np.float32)
    labels = synthetic_labels(n_images, classes, size // 32, size // 32, seed=3)
    fn = lambda: train_step_oracle(x, params, classes, ANCHORS_VOC, labels, HPARAM, dtype=torch.float32)
    threads = _best_threads(lambda: train_step_oracle(x[:1], params, classes, ANCHORS_VOC, [t[:1] for t in labels], HPARAM, dtype=torch.float32))
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    desc = ("%d step(s) x %d images: torch-CPU float32 autograd restatement of the reference's training step (forward with batch "
            "statistics + 4-part loss + backward; oneDNN, best of {all,64,32,16,8} = %d threads; TF1 is not installable)" % (steps, n_images, threads))
    return steps * n_images / dt, threads, desc, dt / steps * 1000.0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params = synthetic_checkpoint(args.classes, 5)
    n = args.cpu_images
    steps = min(args.steps, 5)
    if args.workload == "train":
        ips, threads, desc, ms = cpu_train_images_per_sec(params, args.classes, args.size, n, min(steps, 3), min(args.warmup, 1))
    else:
        ips, threads, desc, ms = cpu_infer_images_per_sec(params, args.classes, args.size, n, steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": metric_name(args), "value": ips,
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def metric_name(args):
    if args.workload == "train":
        return "images/sec YOLOv2-Darknet19 %dpx training step (fwd+loss+bwd+Adam)" % args.size
    return "images/sec YOLOv2-Darknet19 %dpx inference+NMS" % args.size


def workload_config(args, batch):
    if args.workload == "train":
        w = ("YOLOv2-Darknet19 %d-class %dx%d training step: forward with batch-statistics BN + 4-part loss + backward + one "
             "all-reduce of the flat gradient bucket (N > 1) + per-tensor clip + Adam, batch %d per GPU, synthetic N(0,1) images, "
             "synthetic boxes (1..8 per image), conditioned random weights" % (args.classes, args.size, args.size, batch))
        par = "dp%d (batch shards, ONE all-reduce per step on 67 M float32 gradients)" % args.gpus
        l2 = "~10 GB of activations / gradients rewritten every step (> 126 MB L2)"
    else:
        w = ("YOLOv2-Darknet19 %d-class %dx%d detection pipeline (detect.py:59-87): uint8 image -> per_image_standardization -> backbone -> "
             "head decode -> NMS (thr %.1f/%.1f, ~900 candidates per 416x416 image) -> detection list, batch %d per GPU, synthetic uniform "
             "uint8 images, conditioned random weights with a peaked final layer" % (args.classes, args.size, args.size, THRESHOLD,
                                                                                   THRESHOLD_IOU, batch))
        par = "dp%d (batch shards, no data-path collective)" % args.gpus
        l2 = "%d rotating input batches + ~2 GB activation workspace rewritten every step (> 126 MB L2)" % args.rotate
    return {"workload": w, "batch_per_gpu": batch, "global_batch": batch * args.gpus, "input": [args.size, args.size, 3],
            "classes": args.classes, "anchors": 5, "parallelism": par, "l2": l2}


# --------------------------------------------------------------------------------------------------------------------
def nms_sweep_block(peaks, points=((13, 100), (13, 1000), (13, 10000), (19, 1000), (19, 10000)), B=512, C=80):
    """BASELINE configs[4] on this GPU: grids 13x13x5 / 19x19x5, K candidates per image above 0.3, batch 512, 80 classes.
    boxes/s = B*N/t, candidates/s = B*K/t, algorithmic bytes = B*(2*4*N*C + 16*N); bit-exactness against the C
    oracle on the first 2 images of every point (the checker, outside the timed region)."""
    import torch
    from oracle.nms_c import nms_c_batch
    from yolo_tf_b200 import _lib
    L = _lib.lib()
    out = []
    for g, K in points:
        rs = np.random.RandomState(5)
        A, cells = 5, g * g
        N = cells * A
        anch = np.asarray(ANCHORS_COCO)
        gy, gx = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
        centre = np.stack([gx, gy], -1).reshape(1, cells, 1, 2) + rs.uniform(0, 1, size=(B, cells, A, 2))
        wh = anch.reshape(1, 1, A, 2) * np.exp(rs.normal(0, 0.5, size=(B, cells, A, 2)))
        lo = (centre - wh / 2).astype(np.float32).reshape(B, N, 2)
        hi = (centre + wh / 2).astype(np.float32).reshape(B, N, 2)
        conf = rs.uniform(0, 0.29, size=(B, N * C)).astype(np.float32)
        cols = np.argsort(rs.random_sample((B, N * C)).astype(np.float32), axis=1)[:, :K] if K * 8 > N * C else None
        for b in range(B):
            pick = cols[b] if cols is not None else rs.choice(N * C, size=K, replace=False)
            conf[b, pick] = rs.uniform(0.3, 1.0, size=K)
        conf = conf.reshape(B, N, C)
        d0, dlo, dhi = torch.from_numpy(conf).cuda(), torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
        work = torch.empty_like(d0)
        nbytes = L.y2_nms_workspace_bytes(B, N, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        times = []
        for rep in range(8):
            work.copy_(d0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.y2_nms(_lib.ptr(work), _lib.ptr(dlo), _lib.ptr(dhi), B, N, C, THRESHOLD, THRESHOLD_IOU, None, None, _lib.ptr(ws), nbytes,
                                _lib.current_stream()))
            e1.record()
            torch.cuda.synchronize()
            if rep >= 3:
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        ref = conf[:2].copy()
        nms_c_batch(ref, lo[:2], hi[:2], THRESHOLD, THRESHOLD_IOU)
        exact = bool(np.array_equal(work[:2].cpu().numpy().view(np.uint32), ref.view(np.uint32)))
        gbs = B * (2 * 4 * N * C + 16 * N) / ms / 1e6
        out.append({"grid": "%dx%dx5" % (g, g), "N": N, "K": K, "B": B, "C": C, "ms": ms, "boxes_per_s": B * N / ms * 1e3,
                    "candidates_per_s": B * K / ms * 1e3, "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / float(peaks["hbm_gbs"]),
                    "bit_exact_vs_c_oracle_2_images": exact})
        del d0, work, ws
    head = [p for p in out if p["N"] == 845 and p["K"] == 1000][0]
    return {"metric": "NMS boxes/sec (BASELINE configs[4]: batch 512, 80 classes, thr 0.3/0.4)", "value": head["boxes_per_s"], "unit": "boxes/s",
            "at": "13x13x5 grid, 1000 candidates per image", "hbm_peak_gbs": float(peaks["hbm_gbs"]), "points": out}


# --------------------------------------------------------------------------------------------------------------------
def run_infer(args, torch, dist, dev, world, rank, local, barrier):
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder, Model, inference
    from yolo_tf_b200.utils.postprocess import detections_device, non_max_suppress_device
    from yolo_tf_b200.utils.preprocess import per_image_standardization
    L = _lib.lib()
    B, size, C = args.batch, args.size, args.classes
    params = synthetic_checkpoint(C, 5)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    builder = Builder.from_values([str(i) for i in range(C)], size, size, ANCHORS_COCO)
    inference.PRECISION = args.precision
    if args.pair >= 0:
        _lib.check(L.y2_set_option(inference._Engine.get(dev, C, 5).h, b"pair", args.pair))
    if args.conv0_tc >= 0:
        _lib.check(L.y2_set_option(inference._Engine.get(dev, C, 5).h, b"conv0_tc", args.conv0_tc))

    rs = np.random.RandomState(100 + rank)
    host_u8 = [torch.from_numpy(synthetic_images_u8(rs, B, size)).pin_memory() for _ in range(args.rotate)]
    dev_u8 = [t.to(dev) for t in host_u8]
    cw = size // 32
    N = cw * cw * 5
    scale = (size / cw, size / cw)

    def step_device(u8):
        """detect.py:59-87 for a batch: standardise -> backbone -> decode -> NMS -> detection list."""
        x = per_image_standardization(u8)
        builder(x)
        m = builder.model
        conf, lo, hi = m.conf.view(B, N, C), m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2)
        non_max_suppress_device(conf, lo, hi, THRESHOLD, THRESHOLD_IOU, check=False)
        return detections_device(conf, lo, hi, THRESHOLD, scale), conf

    # ---------------- warm-up
    for i in range(args.warmup):
        step_device(dev_u8[i % args.rotate])
    barrier()
    _lib.check(L.y2_check_async_errors())
    # how busy is the NMS? candidates above the threshold before it, detections after it (outside the timed regions)
    x = per_image_standardization(dev_u8[0])
    builder(x)
    cands_per_image = float((builder.model.conf > THRESHOLD).sum().item()) / B
    (count, _, _, _, _), _ = step_device(dev_u8[0])
    dets_per_image = float(count.sum().item()) / B

    # ---------------- timed region 1: device-resident inputs ("value")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.y2_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_device(dev_u8[i % args.rotate])
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = L.y2_launch_count() - launches0
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # ---------------- timed region 2: end to end from pinned host memory through the public API
    copy_stream, comp_stream, out_stream = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_run(host_in, make_step, out_shapes, steps, timed):
        xbuf = [torch.empty_like(host_in[0], device=dev) for _ in range(2)]
        out_host = [[torch.empty(s, dtype=d).pin_memory() for s, d in out_shapes] for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        barrier()
        if timed:
            e0.record(copy_stream)
        for i in range(steps):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[s])
                xbuf[s].copy_(host_in[i % args.rotate], non_blocking=True)
                copied[s].record(copy_stream)
            with torch.cuda.stream(comp_stream):
                comp_stream.wait_event(copied[s])
                outs = make_step(xbuf[s])
                consumed[s].record(comp_stream)
            with torch.cuda.stream(out_stream):          # D2H of the step's result overlaps the next step's compute
                out_stream.wait_event(consumed[s])
                for dst, src in zip(out_host[s], outs):
                    src.record_stream(out_stream)
                    dst.copy_(src, non_blocking=True)
        if timed:
            e1.record(out_stream)
        barrier()
        return sum(x.numel() * x.element_size() for x in out_host[0])

    det_shapes = [((B,), torch.int32), ((B, N), torch.int32), ((B, N), torch.int32), ((B, N), torch.float32), ((B, N, 4), torch.float32)]
    e2e_run(host_u8, lambda u8: step_device(u8)[0], det_shapes, 3, False)
    # three blocks of exactly K steps each, every block timed on the device (max over ranks); the MEDIAN block is reported and all
    # three are listed: this region is driven by ~35 host-side launches per 3 ms step, so one descheduled host thread on a shared
    # box shows up as a slow block (profiles/README.md, round 2: one block at 6.4 ms/step between runs at 3.1)
    e2e_blocks = []
    for _ in range(3):
        d2h_bytes = e2e_run(host_u8, lambda u8: step_device(u8)[0], det_shapes, args.steps, True)
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_blocks.append(float(t.item()))
    e2e_ms = sorted(e2e_blocks)[1]

    # comparison variant (round 1's e2e path): float32 standardised images in, the full post-NMS score matrix + boxes out
    host_f32 = [torch.from_numpy(np.random.RandomState(200 + rank + i).normal(0, 1, size=(B, size, size, 3)).astype(np.float32)).pin_memory()
                for i in range(args.rotate)]

    def step_f32(x):
        builder(x)
        m = builder.model
        conf = m.conf.view(B, N, C)
        non_max_suppress_device(conf, m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2), THRESHOLD, THRESHOLD_IOU, check=False)
        return conf, m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2)

    f32_shapes = [((B, N, C), torch.float32), ((B, N, 2), torch.float32), ((B, N, 2), torch.float32)]
    f32_steps = max(3, args.steps // 4)
    e2e_run(host_f32, step_f32, f32_shapes, 3, False)
    f32_d2h = e2e_run(host_f32, step_f32, f32_shapes, f32_steps, True)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    f32_ms = float(t.item())
    del host_f32
    clocks = sampler.stop() if rank == 0 else None
    _lib.check(L.y2_check_async_errors())

    # ---------------- per-layer device times (CUDA events inside y2_darknet_forward), same workload
    eng = inference._Engine.get(dev, C, 5)
    _lib.check(L.y2_set_profiling(eng.h, 1))
    nl = L.y2_num_layers(eng.h)
    conv_ms, post_ms = (ctypes.c_float * nl)(), (ctypes.c_float * nl)()
    acc_conv, acc_post, reps = np.zeros(nl), np.zeros(nl), min(args.steps, 10)
    pre_ms = head_ms = 0.0
    for i in range(reps):
        h0, h1, h2, h3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        h0.record()
        x = per_image_standardization(dev_u8[i % args.rotate])
        h1.record()
        _, out = inference.darknet(x, C, 5)
        h2.record()
        m = Model(out, C, builder.anchors)
        conf, lo, hi = m.conf.view(B, N, C), m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2)
        non_max_suppress_device(conf, lo, hi, THRESHOLD, THRESHOLD_IOU, check=False)
        detections_device(conf, lo, hi, THRESHOLD, scale)
        h3.record()
        torch.cuda.synchronize()
        _lib.check(L.y2_get_layer_ms(eng.h, conv_ms, post_ms))
        acc_conv += np.array(conv_ms[:])
        acc_post += np.array(post_ms[:])
        pre_ms += h0.elapsed_time(h1)
        head_ms += h2.elapsed_time(h3)
    _lib.check(L.y2_set_profiling(eng.h, 0))
    acc_conv /= reps
    acc_post /= reps
    pre_ms /= reps
    head_ms /= reps

    if rank != 0:
        return None

    peaks, peak_kind = measured_peaks()
    # DRAM traffic of the 21 conv launches of one step, from the committed ncu --set full capture of this workload
    traffic, tpath = None, None
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "conv_step_traffic_*.json")))   # newest capture (tags sort by round)
    if tfiles and (B, size, C) == (32, 416, 80):
        tpath = tfiles[-1]
        traffic = json.load(open(tpath)).get("traffic_bytes")
    imgs = B * world * args.steps
    value = imgs / (ms_total / 1e3)
    flops_tc = conv_flops(size, size, C, 5, tensor_core_only=True) * B       # algorithmic 2*MAC of conv1..conv20+final
    tc_ms = float(acc_conv[1:].sum())
    achieved = flops_tc / (tc_ms / 1e3) / 1e12
    peak, which = pick_peak(peaks, ms_total / 1e3)
    layers = inference.layer_geometry(C, 5)
    table, hh = [], size
    for i, (name, k, cin, cout, has_bn, pool) in enumerate(layers):
        fl = 2 * hh * hh * k * k * cin * cout * B
        table.append({"layer": name, "k": k, "cin": cin, "cout": cout, "hw": hh, "conv_ms": round(float(acc_conv[i]), 4),
                      "post_ms": round(float(acc_post[i]), 4), "algorithmic_tflops": round(fl / (acc_conv[i] / 1e3) / 1e12, 2) if acc_conv[i] > 0 else None})
        if pool:
            hh //= 2
    if args.layer_report:
        with open(args.layer_report, "w") as f:
            json.dump({"batch": B, "size": size, "classes": C, "layers": table, "standardize_ms": pre_ms, "head_decode_nms_detections_ms": head_ms,
                       "tc_conv_ms": tc_ms, "conv0_ms": float(acc_conv[0]), "pool_reorg_ms": float(acc_post.sum())}, f, indent=1)
    mult = 3.0 if args.precision == 0 else 1.0
    h2d = B * size * size * 3
    return {
        "metric": metric_name(args), "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 hi/lo operands, 3 tcgen05 MMAs per product, fp32 accumulate; fp32-grade 1e-4 parity)" if args.precision == 0 else "bf16",
        "data": "synthetic", "config": workload_config(args, B),
        "nms_load": {"candidates_per_image": cands_per_image, "detections_per_image": dets_per_image},
        "e2e": {"value": imgs / (e2e_ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / args.steps, "blocks_ms_per_step": [b / args.steps for b in e2e_blocks],
                "blocks": "3 blocks of K steps; the median block is reported",
                "path": "pinned host uint8 images -> H2D (copy stream, double-buffered) -> per_image_standardization -> Builder(x) -> "
                        "non_max_suppress_device -> detections_device -> D2H of (count, box, class, score, xywh)",
                "f32_variant": {"value": B * world * f32_steps / (f32_ms / 1e3), "unit": "images/s", "steps": f32_steps,
                                "h2d_bytes_per_step": B * size * size * 3 * 4, "d2h_bytes_per_step": f32_d2h,
                                "path": "round 1's path: pinned host float32 images -> H2D -> Builder(x) -> NMS -> D2H of the full score matrix + boxes"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (21 tcgen05 conv launches per step, conv1..conv20 + final; 3x3 layers with 256-wide N tiles as CTA pairs / cta_group::2 unless --pair 0; accumulation chains capped at 32 k-blocks for fp32-grade parity)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": ("profiles/" + os.path.basename(tpath) + ": dram__bytes_read+write summed over the 21 conv launches of one step (ncu --set full)") if traffic else None,
                     "peak_source": "%s %s: the timed region lasted %.2f s" % (peak_kind, which, ms_total / 1e3),
                     "frac_of_burst_peak": achieved / float(peaks.get("bf16_tflops", peak)),
                     "frac_of_sustained_peak": achieved / float(peaks.get("bf16_tflops_sustained", peak)),
                     "algorithmic_flops_per_step": flops_tc, "kernel_ms_per_step": tc_ms,
                     "tensor_pipe_frac_incl_3x_split": mult * achieved / peak,
                     "share_of_step": {"standardize_ms": pre_ms, "tc_conv_ms": tc_ms, "conv0_pool_ms": float(acc_conv[0]),
                                       "pool_reorg_ms": float(acc_post.sum()), "decode_nms_detections_ms": head_ms}},
    }, params


def run_train(args, torch, dist, dev, world, rank, local, barrier):
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder
    from yolo_tf_b200.optimizer import AdamOptimizer, create_train_op
    from yolo_tf_b200.utils.data import transform_labels_batch
    L = _lib.lib()
    B, size, C = args.batch, args.size, args.classes
    params = synthetic_checkpoint(C, 5, dense_detections=False)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    builder = Builder.from_values([str(i) for i in range(C)], size, size, ANCHORS_VOC, hparam=HPARAM)
    train_op = create_train_op(builder, AdamOptimizer(1e-6), clip_gradient_norm=args.clip)       # train.py:127-129,160
    from yolo_tf_b200.model.yolo2 import inference
    for key, val in (("train_f16", args.train_f16), ("train_kcap", args.train_kcap)):
        if val >= 0:
            _lib.check(L.y2_set_option(inference._Engine.get(dev, C, 5).h, key.encode(), val))
    rs = np.random.RandomState(100 + rank)
    host_x = [torch.from_numpy(rs.normal(0, 1, size=(B, size, size, 3)).astype(np.float32)).pin_memory() for _ in range(2)]
    dev_x = [t.to(dev) for t in host_x]
    # synthetic boxes -> the six label tensors, by the device label encoder (utils/data/__init__.py:112-145)
    dev_lab = [list(transform_labels_batch(*synthetic_boxes(rs, B, C), C, size // 32, size // 32, device=dev)) for _ in range(2)]
    host_lab = [[t.cpu().pin_memory() for t in lab] for lab in dev_lab]

    for i in range(args.warmup):
        train_op(dev_x[i % 2], dev_lab[i % 2])
    barrier()
    _lib.check(L.y2_check_async_errors())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.y2_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        loss = train_op(dev_x[i % 2], dev_lab[i % 2])
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = L.y2_launch_count() - launches0
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # end to end: images + the six label tensors from pinned host memory every step, the loss read back
    copy_stream = torch.cuda.Stream()
    xbuf = [torch.empty_like(dev_x[0]) for _ in range(2)]
    labbuf = [[torch.empty_like(t) for t in dev_lab[0]] for _ in range(2)]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    copied, consumed = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()

    def e2e_loop(steps, timed):
        barrier()
        if timed:
            e0.record(copy_stream)
        for i in range(steps):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[s])
                xbuf[s].copy_(host_x[s], non_blocking=True)
                for d, h in zip(labbuf[s], host_lab[s]):
                    d.copy_(h, non_blocking=True)
                copied[s].record(copy_stream)
            cur.wait_event(copied[s])
            loss = train_op(xbuf[s], labbuf[s])
            consumed[s].record(cur)
            loss_host.copy_(loss, non_blocking=True)
        if timed:
            e1.record(cur)
        barrier()

    e2e_loop(2, False)
    e2e_loop(args.steps, True)
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # the collective alone, and the phases of one step (CUDA events), for the report
    flat, views = builder.backward(allreduce=False)
    ar_ms = None
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(flat)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 5
    ph = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ph[0].record()
    builder(dev_x[0], training=True)
    ph[1].record()
    builder.create_objectives(dev_lab[0])
    ph[2].record()
    flat, views = builder.backward(allreduce=True)
    ph[3].record()
    train_op.apply_gradients(flat, views)
    ph[4].record()
    torch.cuda.synchronize()
    _lib.check(L.y2_check_async_errors())
    phases = {"forward_ms": ph[0].elapsed_time(ph[1]), "loss_ms": ph[1].elapsed_time(ph[2]),
              "backward_allreduce_ms": ph[2].elapsed_time(ph[3]), "clip_adam_ms": ph[3].elapsed_time(ph[4])}
    if rank != 0:
        return None
    peaks, peak_kind = measured_peaks()
    imgs = B * world * args.steps
    value = imgs / (ms_total / 1e3)
    gf_img = train_flops_per_image(size, C) / 1e9
    achieved = value / world * gf_img / 1e3
    peak, which = pick_peak(peaks, ms_total / 1e3)
    lab_bytes = sum(t.numel() * t.element_size() for t in host_lab[0])
    return {
        "metric": metric_name(args), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 hi/lo operands, 3 tcgen05 MMAs per product, fp32 accumulate) for fwd / dgrad / wgrad GEMMs; fp32 BN, loss, Adam",
        "data": "synthetic", "config": workload_config(args, B),
        "e2e": {"value": imgs / (e2e_ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": B * size * size * 3 * 4 + lab_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
                "path": "pinned host float32 images + 6 label tensors -> H2D (copy stream, double-buffered) -> create_train_op(...)(data, labels) -> D2H of the total loss"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (forward + dgrad) and wgrad_tc_kernel: the 3 GEMM passes of the step; the step also holds the HBM-bound BN / pool / loss / Adam passes, so this is the WHOLE-STEP algorithmic rate",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": "%s %s: the timed region lasted %.2f s" % (peak_kind, which, ms_total / 1e3),
                     "frac_of_burst_peak": achieved / float(peaks.get("bf16_tflops", peak)),
                     "frac_of_sustained_peak": achieved / float(peaks.get("bf16_tflops_sustained", peak)),
                     "algorithmic_gflop_per_image": gf_img, "tensor_pipe_frac_incl_3x_split": 3.0 * achieved / peak},
        "train": {"phases_ms": phases, "allreduce_ms_alone": ar_ms, "grad_bucket_mb": flat.numel() * 4 / 1e6, "total_loss": float(loss_host.item())},
    }, params


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train"])
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 32 infer / 64 train)")
    ap.add_argument("--rotate", type=int, default=4, help="distinct input batches cycled through")
    ap.add_argument("--precision", type=int, default=0, help="0 = split-bf16 x3 (fp32-grade parity), 1 = single bf16 pass")
    ap.add_argument("--pair", type=int, default=-1, help="CTA-pair (cta_group::2) conv mode: -1 = library default, 0 off, 1 = 3x3 N=256 layers, 2 = all eligible")
    ap.add_argument("--conv0-tc", type=int, default=-1, help="conv0 on the tensor cores (y2_set_option conv0_tc): -1 = library default")
    ap.add_argument("--clip", type=float, default=1.0, help="train: per-tensor clip_by_norm (train.py:128; 0 = off)")
    ap.add_argument("--train-f16", type=int, default=-1, help="train: fp16 planes in the training forward (y2_set_option train_f16); -1 = library default (1)")
    ap.add_argument("--train-kcap", type=int, default=-1, help="train: accumulation-chain cap of the training forward in k-blocks (train_kcap); -1 = library default (8)")
    ap.add_argument("--cpu-images", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nms-sweep", action="store_true")
    ap.add_argument("--layer-report", default="", help="write the per-layer timing table (JSON) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    train = args.workload == "train"
    if args.steps is None:
        args.steps = 40 if train else 200
    if args.classes is None:
        args.classes = 20 if train else 80
    if args.batch is None:
        args.batch = 64 if train else 32
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: whatever NCCL_DEBUG level the environment asks for (its version banner
        # included) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = (run_train if train else run_infer)(args, torch, dist, dev, world, rank, local, barrier)
    if rank == 0:
        line, params = res
        if world == 1:
            peaks, _ = measured_peaks()
            if not train and not args.no_nms_sweep:
                line["nms"] = nms_sweep_block(peaks)
            if not args.no_cpu_baseline:
                if train:
                    ips, threads, desc, _ = cpu_train_images_per_sec(params, args.classes, args.size, 2, 1, 1)
                else:
                    ips, threads, desc, _ = cpu_infer_images_per_sec(params, args.classes, args.size, args.cpu_images, 2, 1)
                line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
