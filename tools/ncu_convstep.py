"""Summarises the `--set full` capture of the 21 tcgen05 conv launches of one inference step
(gpurun_out/prof_convstep_<tag>.ncu-rep, made by tools/profile2.sh) into
  profiles/ncu_<tag>_conv_step.txt      per-launch metrics
  profiles/conv_step_traffic_<tag>.json  DRAM bytes of the step (bench.py's roofline.traffic reads the newest one)
Usage: python tools/ncu_convstep.py <tag>"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv, ARGV = [sys.argv[0], "none"], sys.argv
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ncu_summary_lib", os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_summary.py"))
src = open(spec.origin).read().split("if __name__")[0]
lib = {}
exec(compile(src, spec.origin, "exec"), lib)

TAG = ARGV[1] if len(ARGV) > 1 else "r1c"
rep = "gpurun_out/prof_convstep_%s.ncu-rep" % TAG
lib["summarize"](rep, "profiles/ncu_%s_conv_step.txt" % TAG,
                 "all 21 tcgen05 conv launches of one inference step (B=32, 416, C=80): conv1..conv20, final")
hdr, units, rows = lib["raw_rows"](rep)
col = {h: i for i, h in enumerate(hdr)}
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
tscale = {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}
rd = sum(float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]] for r in rows)
wr = sum(float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]] for r in rows)
ms = sum(float(r[col["gpu__time_duration.sum"]]) * tscale[units[col["gpu__time_duration.sum"]]] for r in rows)
out = {"source": "ncu --set full -k regex:conv_tc_kernel -c 21 of the first timed step (one detection step, B=32, 416, C=80), tag " + TAG,
       "launches": len(rows), "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
       "kernel_time_under_ncu_ms": ms}
with open("profiles/conv_step_traffic_%s.json" % TAG, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
