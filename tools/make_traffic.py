#!/usr/bin/env python
"""profiles/r02_traffic.json from an `ncu --set full` summary (tools/ncu_summary.py ops): DRAM bytes of the dominant conv launch — the first
conv_tcw_kernel row is the forward 36->36 3x3x3 conv on 2x36x128^3 (tools/ncu_ops.py runs that layer first).  bench.py reads the file.
usage: python tools/make_traffic.py profiles/r02_ncu_ops_summary.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
for line in open(src):
    f = line.split()
    if f and f[0] == "conv_tcw_kernel":
        ms, rd, wr = float(f[3]), float(f[4]), float(f[5])
        out = {"entries": [{"input_shape": [2, 36, 128, 128, 128], "kernel": "conv_tcw_kernel (fprop 36->36 k3x3x3)", "dram_bytes": (rd + wr) * 1e9,
                            "dram_read_bytes": rd * 1e9, "dram_write_bytes": wr * 1e9, "ncu_ms": ms,
                            "algorithmic_bytes": 2 * 36 * 128 ** 3 * 4 * 2 + 36 * 36 * 27 * 4,
                            "source": os.path.relpath(os.path.abspath(src), ROOT) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}]}
        json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
        print(out)
        break
else:
    raise SystemExit("no conv_tcw_kernel row in " + src)
