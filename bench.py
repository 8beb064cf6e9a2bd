#!/usr/bin/env python
"""bench.py -- images/sec of the YOLOv2-Darknet19 detection hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]                (the B200-native arm)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   (the reference's CPU path)
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): YOLOv2-Darknet19, 80 classes, 416x416, inference + head decode +
NMS, batch 32 per GPU, synthetic N(0,1) images, synthetic conditioned random weights.  One "step" =
one batch through backbone -> decode -> NMS.  Images are independent, so N GPUs = N shards of 32
(weak scaling, no data-path collective).

Prints ONE JSON line (rank 0): value = device-resident throughput; e2e = through the public API
(Builder + non_max_suppress_device) from pinned host buffers with H2D/D2H inside the timed region;
roofline = the tcgen05 conv kernel's algorithmic FLOP/s vs the measured bf16 peak; cpu_baseline =
the CPU oracle (torch-CPU conv stack + numpy decode + the reference-shaped Python NMS) on a bounded sample.
"""
import argparse
import ctypes
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
THRESHOLD, THRESHOLD_IOU = 0.3, 0.4          # detect.py:127-128


def synthetic_checkpoint(classes, num_anchors, seed=1):
    """Conditioned random weights (He-style, non-trivial BN statistics) under the TF variable names."""
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
    return p


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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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


def cpu_path_images_per_sec(params, classes, size, n_images, steps, warmup, seed=7):
    """The reference's CPU path restated (oracle/): conv stack on all host threads + decode + Python NMS.
    The ONLY place bench.py touches oracle/.  Returns (images/s, threads, description)."""
    import torch
    from oracle.darknet_oracle import darknet_oracle
    from oracle.head_oracle import decode_oracle
    from oracle.nms_oracle import nms_oracle
    rs = np.random.RandomState(seed)
    x = rs.normal(0, 1, size=(n_images, size, size, 3)).astype(np.float32)
    # be fair to the CPU: oneDNN does not scale to every core count on these shapes, so pick the
    # fastest thread count among {all, 64, 32, 16, 8} on one image before timing
    cores = os.cpu_count() or 1
    best, threads = None, cores
    for t in sorted({cores, 64, 32, 16, 8}, reverse=True):
        if t > cores:
            continue
        torch.set_num_threads(t)
        darknet_oracle(x[:1], params, classes, len(ANCHORS_COCO))          # warm
        t0 = time.perf_counter()
        darknet_oracle(x[:1], params, classes, len(ANCHORS_COCO))
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, threads = dt, t
    torch.set_num_threads(threads)

    def one_pass():
        net = darknet_oracle(x, params, classes, len(ANCHORS_COCO))
        m = decode_oracle(net, classes, ANCHORS_COCO)
        for b in range(n_images):
            conf = np.ascontiguousarray(m["conf"][b])
            nms_oracle(conf, np.ascontiguousarray(m["xy_min"][b]), np.ascontiguousarray(m["xy_max"][b]), THRESHOLD, THRESHOLD_IOU)

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    desc = ("%d step(s) x %d images of the same workload: torch-CPU fp32 conv stack (oneDNN, best of {all,64,32,16,8} = %d threads; TF1 itself is not "
            "installable, torch-CPU is the stand-in and is expected to be faster than TF-1.0 Eigen) + numpy decode + "
            "reference-shaped pure-Python NMS (1 thread)" % (steps, n_images, threads))
    return steps * n_images / dt, threads, desc, dt / steps * 1000.0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params = synthetic_checkpoint(args.classes, 5)
    n = args.cpu_images
    steps = min(args.steps, 5)
    ips, threads, desc, ms = cpu_path_images_per_sec(params, args.classes, args.size, n, steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": "images/sec YOLOv2-Darknet19 %dpx inference+NMS" % args.size, "value": ips,
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {"workload": "YOLOv2-Darknet19 %d-class %dx%d inference + head decode + NMS (thr %.1f/%.1f), batch %d per GPU, "
                        "synthetic N(0,1) images, conditioned random weights" % (args.classes, args.size, args.size, THRESHOLD,
                                                                               THRESHOLD_IOU, batch),
            "batch_per_gpu": batch, "global_batch": batch * args.gpus, "input": [args.size, args.size, 3],
            "classes": args.classes, "anchors": 5, "parallelism": "dp%d (batch shards, no data-path collective)" % args.gpus,
            "l2": "%d rotating input batches + ~2 GB activation workspace rewritten every step (> 126 MB L2)" % args.rotate}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--classes", type=int, default=80)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--rotate", type=int, default=4, help="distinct input batches cycled through")
    ap.add_argument("--precision", type=int, default=0, help="0 = split-bf16 x3 (fp32-grade parity), 1 = single bf16 pass")
    ap.add_argument("--pair", type=int, default=-1, help="CTA-pair (cta_group::2) conv mode: -1 = library default, 0 off, 1 = 3x3 N=256 layers, 2 = all eligible")
    ap.add_argument("--conv0-tc", type=int, default=-1, help="conv0 on the tensor cores (y2_set_option conv0_tc): -1 = library default")
    ap.add_argument("--cpu-images", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layer-report", default="", help="write the per-layer timing table (JSON) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder, inference
    from yolo_tf_b200.utils.postprocess import non_max_suppress_device

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
    host_in = [torch.from_numpy(rs.normal(0, 1, size=(B, size, size, 3)).astype(np.float32)).pin_memory()
               for _ in range(args.rotate)]
    dev_in = [t.to(dev) for t in host_in]
    cells = (size // 32) ** 2
    N = cells * 5

    def step_device(x):
        builder(x)
        m = builder.model
        conf = m.conf.view(B, N, C)
        non_max_suppress_device(conf, m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2), THRESHOLD, THRESHOLD_IOU, check=False)
        return conf, m.xy_min, m.xy_max

    # ---------------- warm-up
    for i in range(args.warmup):
        step_device(dev_in[i % args.rotate])
    barrier()
    _lib.check(L.y2_check_async_errors())

    # ---------------- timed region 1: device-resident inputs ("value")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.y2_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_device(dev_in[i % args.rotate])
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
    xbuf = [torch.empty_like(dev_in[0]) for _ in range(2)]
    out_host = [(torch.empty((B, N, C), dtype=torch.float32).pin_memory(), torch.empty((B, N, 2), dtype=torch.float32).pin_memory(),
                 torch.empty((B, N, 2), dtype=torch.float32).pin_memory()) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_loop(steps, timed):
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
                conf, lo, hi = step_device(xbuf[s])
                consumed[s].record(comp_stream)
            with torch.cuda.stream(out_stream):          # D2H of the step's result overlaps the next step's compute
                out_stream.wait_event(consumed[s])
                for t in (conf, lo, hi):
                    t.record_stream(out_stream)
                out_host[s][0].copy_(conf, non_blocking=True)
                out_host[s][1].copy_(lo.view(B, N, 2), non_blocking=True)
                out_host[s][2].copy_(hi.view(B, N, 2), non_blocking=True)
        if timed:
            e1.record(out_stream)
        barrier()

    e2e_loop(3, False)
    e2e_loop(args.steps, True)
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    # diagnostic: the H2D copy alone (what the copy stream must hide under the compute of the previous step)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for i in range(5):
        xbuf[0].copy_(host_in[i % args.rotate], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_ms = c0.elapsed_time(c1) / 5
    clocks = sampler.stop() if rank == 0 else None
    _lib.check(L.y2_check_async_errors())

    # ---------------- per-layer device times (CUDA events inside y2_darknet_forward), same workload
    eng = inference._Engine.get(dev, C, 5)
    _lib.check(L.y2_set_profiling(eng.h, 1))
    nl = L.y2_num_layers(eng.h)
    conv_ms, post_ms = (ctypes.c_float * nl)(), (ctypes.c_float * nl)()
    acc_conv, acc_post, reps = np.zeros(nl), np.zeros(nl), min(args.steps, 10)
    head_ms = 0.0
    for i in range(reps):
        h0, h1, h2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        builder.output = None
        h0.record()
        _, out = inference.darknet(dev_in[i % args.rotate], C, 5)
        h1.record()
        builder.output = out
        from yolo_tf_b200.model.yolo2 import Model
        m = Model(out, C, builder.anchors)
        non_max_suppress_device(m.conf.view(B, N, C), m.xy_min.view(B, N, 2), m.xy_max.view(B, N, 2), THRESHOLD, THRESHOLD_IOU, check=False)
        h2.record()
        torch.cuda.synchronize()
        _lib.check(L.y2_get_layer_ms(eng.h, conv_ms, post_ms))
        acc_conv += np.array(conv_ms[:])
        acc_post += np.array(post_ms[:])
        head_ms += h1.elapsed_time(h2)
    _lib.check(L.y2_set_profiling(eng.h, 0))
    acc_conv /= reps
    acc_post /= reps
    head_ms /= reps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    # DRAM traffic of the 21 conv launches of one step, from the committed ncu --set full capture of this workload
    traffic = None
    import glob
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "conv_step_traffic_*.json")))   # newest capture (tags sort by round)
    tpath = tfiles[-1] if tfiles else os.path.join(ROOT, "profiles", "conv_step_traffic_r1.json")
    if os.path.exists(tpath) and (B, size, C) == (32, 416, 80):
        traffic = json.load(open(tpath)).get("traffic_bytes")
    imgs = B * world * args.steps
    value = imgs / (ms_total / 1e3)
    e2e_value = imgs / (e2e_ms / 1e3)
    flops_tc = conv_flops(size, size, C, 5, tensor_core_only=True) * B       # algorithmic 2*MAC of conv1..conv20+final
    tc_ms = float(acc_conv[1:].sum())
    achieved = flops_tc / (tc_ms / 1e3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
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
            json.dump({"batch": B, "size": size, "classes": C, "layers": table, "head_decode_nms_ms": head_ms,
                       "tc_conv_ms": tc_ms, "conv0_ms": float(acc_conv[0]), "pool_reorg_ms": float(acc_post.sum())}, f, indent=1)

    line = {
        "metric": "images/sec YOLOv2-Darknet19 %dpx inference+NMS" % size, "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 hi/lo operands, 3 tcgen05 MMAs per product, fp32 accumulate; fp32-grade 1e-4 parity)" if args.precision == 0 else "bf16",
        "data": "synthetic", "config": workload_config(args, B),
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * size * size * 3 * 4,
                "d2h_bytes_per_step": B * N * (C + 4) * 4, "ms_per_step": e2e_ms / args.steps,
                "h2d_ms_alone": h2d_ms, "h2d_gbs": B * size * size * 3 * 4 / h2d_ms / 1e6,
                "path": "pinned host -> H2D (copy stream, double-buffered) -> Builder(x) -> model.conf/xy_min/xy_max -> non_max_suppress_device -> D2H"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (21 tcgen05 conv launches per step, conv1..conv20 + final; 3x3 layers with 256-wide N tiles as CTA pairs / cta_group::2 unless --pair 0; accumulation chains capped at 32 k-blocks for fp32-grade parity)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "profiles/" + os.path.basename(tpath) + ": dram__bytes_read+write summed over the 21 conv launches of one step (ncu --set full)" if traffic else None,
                     "peak_source": "%s MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" % peak_kind,
                     "algorithmic_flops_per_step": flops_tc, "kernel_ms_per_step": tc_ms,
                     "tensor_pipe_frac_incl_3x_split": (3.0 if args.precision == 0 else 1.0) * achieved / peak,
                     "share_of_step": {"tc_conv_ms": tc_ms, "conv0_pool_ms": float(acc_conv[0]), "pool_reorg_ms": float(acc_post.sum()),
                                       "decode_nms_ms": head_ms}},
    }
    if world == 1 and not args.no_cpu_baseline:
        ips, threads, desc, _ = cpu_path_images_per_sec(params, C, size, args.cpu_images, 2, 1)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": desc}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
