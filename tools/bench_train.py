"""BASELINE config 3: YOLOv2-Darknet19 20-class 416x416 TRAINING step (forward with batch-statistics BN + loss +
backward, no optimizer), batch 64 per GPU, synthetic images and boxes.  Under torchrun each rank trains its own shard
and the flat gradient bucket is averaged with ONE NCCL all-reduce per step (Builder.backward).

Prints one JSON line: images/s (max over ranks), algorithmic TFLOP/s (104.39 GFLOP per image = fwd + dgrad + wgrad,
SURVEY.md section 8d) against the measured bf16 peak, the all-reduce time, and the CPU oracle on a bounded sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402  (synthetic_checkpoint, conv_flops, measured_peaks)

ANCHORS_VOC = [[1.08, 1.19], [3.42, 4.41], [6.63, 11.38], [9.42, 5.11], [16.62, 10.52]]
HPARAM = {"prob": 1.0, "iou_best": 5.0, "iou_normal": 1.0, "coords": 1.0}


def synthetic_labels(batch, classes, cw, ch, seed):
    """Label tensors in the layout of utils/data/__init__.py:112-145 (mask, prob, coords, offset_xy_min/max, areas)."""
    rs = np.random.RandomState(seed)
    cells = cw * ch
    mask = np.zeros((batch, cells, 1), np.float32)
    prob = np.zeros((batch, cells, 1, classes), np.float32)
    coords = np.zeros((batch, cells, 1, 4), np.float32)
    lo = np.zeros((batch, cells, 1, 2), np.float32)
    hi = np.zeros((batch, cells, 1, 2), np.float32)
    for b in range(batch):
        for _ in range(rs.randint(1, 9)):
            c = rs.randint(0, classes)
            cx, cy = rs.uniform(0, 1, 2)
            w, h = rs.uniform(0.05, 0.6, 2)
            xmin, xmax = max(cx - w / 2, 0), min(cx + w / 2, 1 - 1e-6)
            ymin, ymax = max(cy - h / 2, 0), min(cy + h / 2, 1 - 1e-6)
            x, y = cw * (xmin + xmax) / 2, ch * (ymin + ymax) / 2
            ix, iy = int(np.floor(x)), int(np.floor(y))
            i = iy * cw + ix
            ww, hh = xmax - xmin, ymax - ymin
            mask[b, i] = 1
            prob[b, i, 0, c] = 1
            coords[b, i, 0] = [x - ix, y - iy, np.sqrt(ww), np.sqrt(hh)]
            lo[b, i, 0] = [x - ix - ww / 2 * cw, y - iy - hh / 2 * ch]
            hi[b, i, 0] = [x - ix + ww / 2 * cw, y - iy + hh / 2 * ch]
    areas = (hi - lo)[..., 0] * (hi - lo)[..., 1]
    return mask, prob, coords, lo, hi, areas


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--pair", type=int, default=-1, help="CTA-pair conv mode (y2_set_option 'pair'); -1 = library default")
    ap.add_argument("--timeline", default="", help="write a kernel timeline summary (torch.profiler / CUPTI) of 2 steps to this JSON")
    args = ap.parse_args()
    from yolo_tf_b200 import _lib, variables
    from yolo_tf_b200.model.yolo2 import Builder
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL banner / debug lines off stdout
        dist.init_process_group("nccl", device_id=dev)
    Bn, size, C = args.batch, args.size, args.classes
    params = B.synthetic_checkpoint(C, 5)
    store = variables.reset_default_store()
    store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
    builder = Builder.from_values([str(i) for i in range(C)], size, size, ANCHORS_VOC, hparam=HPARAM)
    if args.pair >= 0:
        from yolo_tf_b200.model.yolo2 import inference
        _lib.check(_lib.lib().y2_set_option(inference._Engine.get(dev, C, 5).h, b"pair", args.pair))
    rs = np.random.RandomState(100 + rank)
    x = torch.from_numpy(rs.normal(0, 1, size=(Bn, size, size, 3)).astype(np.float32)).to(dev)
    labels = [torch.from_numpy(t).to(dev) for t in synthetic_labels(Bn, C, size // 32, size // 32, 3 + rank)]

    def step():
        builder(x, training=True)
        builder.create_objectives(labels)
        flat, _ = builder.backward(allreduce=True)
        return flat

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    if world > 1:
        dist.barrier()
    if args.timeline:
        timeline(step, args.timeline)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.lib().y2_launch_count()
    e0.record()
    for _ in range(args.steps):
        flat = step()
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.lib().y2_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # all-reduce alone
    ar_ms = None
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(flat)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 5
    # optimizer step alone: Adam + per-tensor clip on the bucket; 28 B per parameter (+4 B for the clip norms)
    from yolo_tf_b200.optimizer import AdamOptimizer, create_train_op
    top = create_train_op(builder, AdamOptimizer(1e-6), clip_gradient_norm=1.0)
    _, views = builder.backward(allreduce=False)
    for _ in range(2):
        top.apply_gradients(flat, views)
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(10):
        top.apply_gradients(flat, views)
    o1.record()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().y2_check_async_errors())
    adam_ms = o0.elapsed_time(o1) / 10
    if rank == 0:
        peaks, kind = B.measured_peaks()
        gf_img = (3 * B.conv_flops(size, size, C, 5) - 2 * size * size * 27 * 32) / 1e9          # fwd + dgrad + wgrad, no dgrad for conv0
        imgs = Bn * world * args.steps
        ips = imgs / (ms / 1e3)
        line = {"metric": "images/sec YOLOv2-Darknet19 %dpx training step (fwd+loss+bwd)" % size, "value": ips, "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps, "batch_per_gpu": Bn, "classes": C,
                "algorithmic_gflop_per_image": gf_img, "algorithmic_tflops": ips * gf_img / 1e3,
                "frac_of_bf16_peak": ips * gf_img / 1e3 / world / float(peaks.get("bf16_tflops_sustained", 1450.3)), "peak_source": kind,
                "gpu_launches": int(launches), "allreduce_ms_alone": ar_ms,
                "adam_clip_step_ms": adam_ms, "adam_clip_algorithmic_GBs": flat.numel() * 32 / adam_ms / 1e6,
                "adam_frac_of_hbm_peak": flat.numel() * 32 / adam_ms / 1e6 / float(peaks.get("hbm_gbs", 6511.9)), "grad_bucket_mb": flat.numel() * 4 / 1e6,
                "total_loss": float(builder.objectives.total_loss())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def timeline(step, path):
    """Kernel timeline of 2 training steps: busy time per kernel, idle time between consecutive kernels."""
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in ev if "emcpy" not in e.name and "emset" not in e.name])
    span = ks[-1][1] - ks[0][0]
    busy, agg, gaps = 0.0, {}, []
    for i, (a, b, n) in enumerate(ks):
        busy += b - a
        key = n.split("(")[0].replace("void ", "").replace("y2::", "")[:48]
        agg.setdefault(key, [0, 0.0])
        agg[key][0] += 1
        agg[key][1] += b - a
        if i + 1 < len(ks):
            g = ks[i + 1][0] - b
            gaps.append((g, key, ks[i + 1][2].split("(")[0].replace("void ", "").replace("y2::", "")[:48]))
    gapsum = sum(g for g, _, _ in gaps if g > 0)
    by_pair = {}
    for g, a, b in gaps:
        if g > 0:
            by_pair.setdefault(a + " -> " + b, [0, 0.0])
            by_pair[a + " -> " + b][0] += 1
            by_pair[a + " -> " + b][1] += g
    out = {"steps": 2, "kernels": len(ks), "span_us": span, "busy_us": busy, "idle_us": gapsum,
           "per_kernel_us": {k: {"n": v[0], "us": v[1]} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
           "idle_by_pair_us": {k: {"n": v[0], "us": v[1]} for k, v in sorted(by_pair.items(), key=lambda kv: -kv[1][1])[:25]}}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("kernels", "span_us", "busy_us", "idle_us")}))
    for k, v in list(out["per_kernel_us"].items())[:14]:
        print("%-50s %4d %10.1f" % (k, v["n"], v["us"]))
    print("-- idle between:")
    for k, v in list(out["idle_by_pair_us"].items())[:14]:
        print("%-100s %4d %10.1f" % (k, v["n"], v["us"]))


if __name__ == "__main__":
    main()
