"""gpurun_out/hbm_<tag>.csv (tools/profile_hbm.sh) -> profiles/ncu_<tag>_hbm_kernels.txt: per kernel (grouped by name and, for
the NMS, by sweep point) launches, time, DRAM bytes, achieved DRAM GB/s against the measured HBM copy peak.
Usage: python tools/hbm_summary.py <tag>"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(peaks["hbm_gbs"])
rows = list(csv.reader(l for l in open(os.path.join(ROOT, "gpurun_out", "hbm_%s.csv" % TAG)) if l.startswith('"')))
col = {h: i for i, h in enumerate(rows[0])}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}
launches = OrderedDict()          # launch ID -> {name, metrics}
for r in rows[1:]:
    lid = r[col["ID"]]
    d = launches.setdefault(lid, {"name": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("y2::", ""), "m": {}})
    v = float(r[col["Metric Value"]].replace(",", ""))
    d["m"][r[col["Metric Name"]]] = v * scale.get(r[col["Metric Unit"]], 1.0)
# phases: the driver profiles detect step, NMS K=100, K=1000, K=10000, training step, in that order
phase_names = ["detection step (B=32, 416, C=80)", "NMS sweep B=512 N=845 K=100", "NMS sweep B=512 N=845 K=1000", "NMS sweep B=512 N=845 K=10000",
               "training step + Adam (B=64, 416, C=20)"]
groups, phase, seen_select = OrderedDict(), 0, 0
for lid, d in launches.items():
    if d["name"].startswith("nms_select"):
        seen_select += 1
        phase = min(seen_select - 1, 4)
    elif phase >= 1 and seen_select >= 4 and not d["name"].startswith("nms_"):
        phase = 4
    key = (phase_names[phase], d["name"])
    g = groups.setdefault(key, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "sm": [], "grid": d["m"].get("launch__grid_size")})
    g["n"] += 1
    g["t"] += d["m"].get("gpu__time_duration.sum", 0.0)
    g["rd"] += d["m"].get("dram__bytes_read.sum", 0.0)
    g["wr"] += d["m"].get("dram__bytes_write.sum", 0.0)
    g["sm"].append(d["m"].get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0))
dst = os.path.join(ROOT, "profiles", "ncu_%s_hbm_kernels.txt" % TAG)
with open(dst, "w") as f:
    f.write("# HBM-bound kernels under ncu (tools/profile_hbm.sh %s: --metrics duration + dram bytes, --clock-control none, cache flushed per\n"
            "# replay = cold-cache DRAM traffic; 1 GPU, under gpurun).  GB/s = (dram read + write) / duration; peak = %.1f GB/s\n"
            "# (MEASURED_PEAKS.json hbm_gbs).  tcgen05 GEMM kernels are excluded (profiles/ncu_*_conv_step.txt).\n" % (TAG, HBM))
    last = None
    for (ph, name), g in groups.items():
        if ph != last:
            f.write("\n## %s\n%-44s %5s %10s %10s %10s %9s %8s %7s\n" % (ph, "kernel", "n", "time_us", "read_MB", "write_MB", "GB/s", "of_peak", "sm_%"))
            last = ph
        gbs = (g["rd"] + g["wr"]) / g["t"] / 1e9 if g["t"] > 0 else 0.0
        f.write("%-44s %5d %10.1f %10.2f %10.2f %9.1f %7.1f%% %6.1f\n" % (name[:44], g["n"], g["t"] * 1e6, g["rd"] / 1e6, g["wr"] / 1e6, gbs,
                                                                       100 * gbs / HBM, sum(g["sm"]) / len(g["sm"])))
print(open(dst).read())
