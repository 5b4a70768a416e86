#!/usr/bin/env python
"""Role-level cycle breakdown of the tap-stacked conv kernel (conv3d_tcw.cu) on the BASELINE layer shapes: kernel-only time (operand already
split, weights packed inside the timed call) and the counters of CTA 0 (MDT_TCW_PROF=1, include/mdt_b200.h: mdt_debug_conv_tcw_prof).
usage: python tools/tcw_prof.py [layer ...] [--once]      env PASS=0|1"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import _lib as L  # noqa: E402
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402

LAYERS = {
    "head64": (64, 64, 3, (1, 1, 1), 1, (32, 32, 128)),
    "p0_36": (36, 36, 3, (1, 1, 1), 1, (128, 128, 128)),
    "c0_18": (18, 18, 3, (1, 1, 1), 1, (128, 128, 128)),
    "c1_k7": (18, 18, 7, (2, 2, 1), 3, (128, 128, 128)),
    "bb54": (64, 54, 3, (1, 1, 1), 1, (32, 32, 128)),
}
NAMES = ["prod_wait_a", "prod_wait_b", "prod_total", "mma_wait_a", "mma_wait_b", "mma_wait_acc", "mma_total", "epi_wait", "epi_total", "tiles"]


def main():
    names = [a for a in sys.argv[1:] if a in LAYERS] or ["p0_36", "c0_18", "c1_k7", "head64"]
    once = "--once" in sys.argv
    ps = int(os.environ.get("PASS", "0"))
    lib = L.load()
    dev = "cuda:0"
    for name in names:
        cin, cout, k, st, pad, sp = LAYERS[name]
        k3, p3 = C._triple(k), C._triple(pad)
        x = torch.randn(2, cin, *sp, device=dev).contiguous(memory_format=torch.channels_last_3d)
        w = torch.randn(cout, cin, *k3, device=dev) * 0.05
        y = C.conv3d_forward(x, w, None, st, p3)          # caches the split form of x on the tensor
        gy = torch.randn_like(y)
        fn = {0: (lambda: C.conv3d_forward(x, w, None, st, p3, relu=True)), 1: (lambda: C.conv3d_dgrad(gy, w, tuple(x.shape), st, p3)),
              2: (lambda: C.conv3d_wgrad(x, gy, tuple(w.shape), st, p3, False))}[ps]
        flops = 2.0 * y.numel() * cin * np.prod(k3)
        if once:
            fn()
            torch.cuda.synchronize()
            continue
        for _ in range(2):
            fn()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        os.environ["MDT_TCW_PROF"] = "1"
        buf = (ctypes.c_ulonglong * 16)()
        L.check(lib.mdt_debug_conv_tcw_prof(buf))
        fn()
        L.check(lib.mdt_debug_conv_tcw_prof(buf))
        os.environ["MDT_TCW_PROF"] = "0"
        v = list(buf)
        tiles = max(v[9], 1)
        print("%-8s pass %d  %.3f ms  %.1f TFLOP/s | per tile (cycles): " % (name, ps, ms, flops / ms / 1e9)
              + "  ".join("%s=%d" % (n, v[i] // tiles) for i, n in enumerate(NAMES[:9])) + "  tiles=%d" % v[9], flush=True)


if __name__ == "__main__":
    main()
