#!/usr/bin/env python
"""fprop / dgrad accuracy of the tensor-core conv on the BASELINE layer shapes (fp64 reference on the GPU), incl. the long-chain 7x7x7 conv."""
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
for cin, cout, k, st, pad, sp in [(18, 18, 7, (2, 2, 1), 3, (32, 32, 128)), (18, 18, 7, (2, 2, 1), 3, (64, 64, 128)), (64, 64, 3, (1, 1, 1), 1, (16, 16, 128)),
                                  (36, 36, 3, (1, 1, 1), 1, (32, 32, 128)), (144, 144, 3, (1, 1, 1), 1, (4, 4, 128))]:
    k3, p3 = C._triple(k), C._triple(pad)
    x = torch.randn(1, cin, *sp, device=dev).contiguous(memory_format=torch.channels_last_3d)
    w = torch.randn(cout, cin, *k3, device=dev) / np.sqrt(cin * np.prod(k3))
    xd, wd = x.double().requires_grad_(True), w.double()
    ref = F.conv3d(xd, wd, None, stride=st, padding=p3)
    gy = torch.randn_like(ref).float().contiguous(memory_format=torch.channels_last_3d)
    ref.backward(gy.double())
    res = []
    for algo in (1, 2):
        y = C.conv3d_forward(x, w, None, st, p3, algo=algo)
        dx = C.conv3d_dgrad(gy, w, tuple(x.shape), st, p3, algo=algo)
        res += [float((y.double() - ref).abs().max() / ref.abs().max()), float((dx.double() - xd.grad).abs().max() / xd.grad.abs().max())]
    print("%3d->%3d k%d s%s %-14s fprop simt %.1e tc %.1e | dgrad simt %.1e tc %.1e" % (cin, cout, k, st, sp, res[0], res[2], res[1], res[3]), flush=True)
