#!/usr/bin/env python
"""Times the pointwise (1x1x1) conv layers of cfg2 through the C-ABI on both paths (algo 4 = fp32 streaming, algo 2 = tcgen05), with the
algorithmic HBM bytes of each pass.  usage: python tools/pw_bench.py   env REPS (default 10)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402

LAYERS = [("P0_conv1", 18, 36, (128, 128, 128), True, False, True), ("final_conv", 36, 2, (128, 128, 128), False, False, False),
          ("C2.conv3", 18, 72, (32, 32, 128), True, True, False), ("C2.conv1", 72, 18, (32, 32, 128), False, True, True),
          ("C3.conv3", 36, 144, (16, 16, 64), True, True, False), ("P1_conv1", 18, 36, (64, 64, 128), True, False, False)]


def timed(fn, flush, reps):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    reps = int(os.environ.get("REPS", "10"))
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    one, zero = (1, 1, 1), (0, 0, 0)
    for name, cin, cout, sp, with_res, relu, emit in LAYERS:
        x = torch.randn(2, cin, *sp, device=dev).contiguous(memory_format=torch.channels_last_3d)
        w = torch.randn(cout, cin, 1, 1, 1, device=dev) * 0.05
        b = torch.randn(cout, device=dev)
        res = torch.randn(2, cout, *sp, device=dev).contiguous(memory_format=torch.channels_last_3d) if with_res else None
        V = x.numel() // cin
        kg = (cout + 15) // 16 * 16
        for algo in (4, 2):
            y = C.conv3d_forward(x, w, b, one, zero, relu=relu, residual=res, algo=algo)
            gy = torch.randn_like(y)
            byt_f = 4 * V * (cin + cout * (2 if with_res else 1)) + (4 * V * kg if emit else 0)
            ms = timed(lambda: C.conv3d_forward(x, w, b, one, zero, relu=relu, residual=res, algo=algo, emit_split=emit), flush, reps)
            print("%-10s %3d->%3d algo %d fprop%s %7.3f ms  %6.0f GB/s" % (name, cin, cout, algo, "+emit" if emit else "     ", ms, byt_f / ms / 1e6), flush=True)
            byt_b = 4 * V * (cin * 2 + cout * (2 if relu else 1))
            ms = timed(lambda: C.conv3d_backward(x, gy, y if relu else None, w, one, zero, True, True, False, algo=algo), flush, reps)
            print("%-10s %3d->%3d algo %d backward   %7.3f ms  %6.0f GB/s" % (name, cin, cout, algo, ms, byt_b / ms / 1e6), flush=True)
            ms = timed(lambda: C.conv3d_dgrad(gy, w, tuple(x.shape), one, zero, algo=algo), flush, reps)
            print("%-10s %3d->%3d algo %d   dgrad    %7.3f ms  %6.0f GB/s" % (name, cin, cout, algo, ms, 4 * V * (cin + cout) / ms / 1e6), flush=True)
            ms = timed(lambda: C.conv3d_wgrad(x, gy, tuple(w.shape), one, zero, True, algo=algo), flush, reps)
            print("%-10s %3d->%3d algo %d   wgrad    %7.3f ms  %6.0f GB/s" % (name, cin, cout, algo, ms, 4 * V * (cin + cout) / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
