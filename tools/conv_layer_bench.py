#!/usr/bin/env python
"""Times single conv layers of the BASELINE shapes through the C-ABI (CUDA events, L2 flushed between iterations).
usage: python tools/conv_layer_bench.py [name ...]   names: head64 p0_36 c0_18 c1_k7 (default: all)   env REPS (default 10), PASSES=012"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402

LAYERS = {
    "head64": (64, 64, 3, (1, 1, 1), 1, (32, 32, 128)),
    "p0_36": (36, 36, 3, (1, 1, 1), 1, (128, 128, 128)),
    "c0_18": (18, 18, 3, (1, 1, 1), 1, (128, 128, 128)),
    "c1_k7": (18, 18, 7, (2, 2, 1), 3, (128, 128, 128)),
    "stem": (1, 18, 3, (1, 1, 1), 1, (128, 128, 128)),
    "head64_p3": (64, 64, 3, (1, 1, 1), 1, (16, 16, 64)),
}


def main():
    names = [a for a in sys.argv[1:] if a in LAYERS] or list(LAYERS)
    reps = int(os.environ.get("REPS", "10"))
    passes = os.environ.get("PASSES", "012")
    prec = int(os.environ.get("PRECISION", "0"))
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in names:
        cin, cout, k, st, pad, sp = LAYERS[name]
        k3, p3 = C._triple(k), C._triple(pad)
        x = torch.randn(2, cin, *sp, device=dev).contiguous(memory_format=torch.channels_last_3d)
        w = torch.randn(cout, cin, *k3, device=dev) * 0.05
        y = C.conv3d_forward(x, w, None, st, p3, precision=prec)
        gy = torch.randn_like(y)
        flops = 2.0 * y.numel() * cin * np.prod(k3)
        fns = {"0": ("fprop", lambda: C.conv3d_forward(x, w, None, st, p3, relu=True, precision=prec)),
               "1": ("dgrad", lambda: C.conv3d_dgrad(gy, w, tuple(x.shape), st, p3, precision=prec)),
               "2": ("wgrad", lambda: C.conv3d_wgrad(x, gy, tuple(w.shape), st, p3, False, precision=prec))}
        for ps in passes:
            label, fn = fns[ps]
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
            ms = ts[len(ts) // 2]
            print("%-10s %-6s %8.3f ms  %7.1f TFLOP/s (algorithmic, incl. operand split/pack)" % (name, label, ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
