#!/usr/bin/env python
"""How does the accuracy of the tensor-core weight gradient behave with the reduction length (number of voxels)?  fp64 reference on the GPU."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
for cin, cout, k in [(36, 36, 3), (18, 72, 1), (64, 64, 3)]:
    for sp in [(8, 8, 32), (16, 16, 64), (32, 32, 128), (64, 64, 128), (128, 128, 128)]:
        if cin == 64 and sp[0] > 64:
            continue
        x = torch.randn(2, cin, *sp, device=dev).contiguous(memory_format=torch.channels_last_3d)
        gy = torch.randn(2, cout, *sp, device=dev).contiguous(memory_format=torch.channels_last_3d)
        gy = gy * (torch.rand_like(gy) > 0.5)
        p = k // 2
        ws = (cout, cin, k, k, k)
        ref = torch.zeros(ws, dtype=torch.float64, device=dev)
        for n in range(2):   # fp64 reference in two halves to bound memory
            xd = x[n:n + 1].double()
            wd = torch.zeros(ws, dtype=torch.float64, device=dev, requires_grad=True)
            F.conv3d(xd, wd, None, padding=p).backward(gy[n:n + 1].double())
            ref += wd.grad
        out = {}
        for algo, name in ((1, "simt"), (2, "tc")):
            dw, _ = C.conv3d_wgrad(x, gy, ws, (1, 1, 1), (p, p, p), False, algo=algo)
            e = (dw.double() - ref)
            out[name] = (float(e.abs().max() / ref.abs().max()), float(e.norm() / ref.norm()))
        print("%2d->%2d k%d %-14s voxels %8d  simt max %.1e l2 %.1e | tc max %.1e l2 %.1e" % (cin, cout, k, sp, 2 * np.prod(sp), *out["simt"], *out["tc"]), flush=True)
