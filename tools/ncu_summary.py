#!/usr/bin/env python
"""Condense ncu exports into the tables kept under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv        (from: ncu --metrics gpu__time_duration.sum --csv --log-file ...)
  python tools/ncu_summary.py ops gpurun_out/ops_raw.csv              (from: ncu -i ops.ncu-rep --page raw --csv)"""
import collections
import csv
import re
import sys

COLS = [("ms", "gpu__time_duration.sum", 1.0), ("dram_rd_GB", "dram__bytes_read.sum", 1.0), ("dram_wr_GB", "dram__bytes_write.sum", 1.0),
        ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("tensor_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue_%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1.0), ("l2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("l1_%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
        ("regs", "launch__registers_per_thread", 1.0)]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("mdt::", "")


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(short(r[k])[:70], [0, 0.0])
        a[0] += 1
        a[1] += float(r[v].replace(",", "")) / 1e3
    total = sum(a[1] for a in agg.values())
    print("launches %d, total %.0f us (cold-cache, serialised under ncu: compare SHARES)" % (len(rows) - 1, total))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("%-72s %5d launches %10.1f us %5.1f%%" % (name, n, us, 100 * us / total))


def ops(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c[1]) for c in COLS]
    kn, gs, bs = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    print("%-34s %-14s %-6s " % ("kernel", "grid", "block") + " ".join("%10s" % c[0] for c in COLS))
    for r in rows[2:]:
        vals = []
        for (label, _, _), i in zip(COLS, idx):
            x = float(r[i].replace(",", ""))
            if label == "ms" and units[i] == "us":
                x /= 1e3
            if label.endswith("GB") and units[i] == "Mbyte":
                x /= 1e3
            if label.endswith("GB") and units[i] == "Kbyte":
                x /= 1e6
            vals.append(x)
        print("%-34s %-14s %-6s " % (short(r[kn])[:34], r[gs].replace(" ", ""), r[bs].split(",")[0].strip("(")) + " ".join("%10.3f" % x for x in vals))


if __name__ == "__main__":
    {"launches": launches, "ops": ops}[sys.argv[1]](sys.argv[2])
