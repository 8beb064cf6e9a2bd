"""Turns ncu captures (gpurun_out/*.ncu-rep, launches_*.csv) into the small text summaries committed under profiles/.
Usage: python tools/ncu_summary.py <tag>"""
import csv
import subprocess
import sys
from collections import OrderedDict, defaultdict

TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summarize(rep, dst, title):
    hdr, units, rows = raw_rows(rep)
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# %s\n# source: %s (ncu --set full --clock-control none --import-source on, 1 GPU, under gpurun)\n" % (title, rep))
        for r in rows:
            f.write("\n%s  grid=%s block=%s\n" % (r[col["Kernel Name"]][:110], r[col["Grid Size"]], r[col["Block Size"]]))
            for m in METRICS:
                if m in col:
                    f.write("  %-68s %s %s\n" % (m, r[col[m]], units[col[m]]))
            rd, wr = float(r[col["dram__bytes_read.sum"]]), float(r[col["dram__bytes_write.sum"]])
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rb = rd * scale.get(units[col["dram__bytes_read.sum"]], 1)
            wb = wr * scale.get(units[col["dram__bytes_write.sum"]], 1)
            f.write("  %-68s %.3f MB\n" % ("traffic = dram read + write", (rb + wb) / 1e6))
    print("wrote", dst)


def launches(csv_path, dst):
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    order = OrderedDict()
    for r in rows[1:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        ns = float(r[col["Metric Value"]])
        agg[name][0] += 1
        agg[name][1] += ns
        order[name] = 1
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# launch list summary: %s (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:\n"
                "# compare SHARES, not absolutes). %d launches, %.3f ms total.\n" % (csv_path, len(rows) - 1, total / 1e6))
        f.write("%-60s %8s %12s %8s\n" % ("kernel", "launches", "total_us", "share"))
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.1f %7.1f%%\n" % (name[:60], n, ns / 1e3, 100 * ns / total))
    print("wrote", dst)


if __name__ == "__main__":
    summarize("gpurun_out/prof_conv_%s.ncu-rep" % TAG, "profiles/ncu_%s_conv_tc.txt" % TAG,
              "tcgen05 conv kernel: conv20 (3x3 3072->1024 @13x13, B=32), final 1x1 (1024->425), conv1 (3x3 32->64 @208x208)")
    summarize("gpurun_out/prof_hbm_%s.ncu-rep" % TAG, "profiles/ncu_%s_hbm_kernels.txt" % TAG,
              "HBM-bound kernels of one inference step (conv0+pool, max-pool, reorg, decode, NMS)")
    launches("gpurun_out/launches_%s.csv" % TAG, "profiles/launches_%s.txt" % TAG)
