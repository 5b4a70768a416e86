#!/usr/bin/env python
"""Debug aid: for every distinct conv configuration met in a model forward, compare the tcgen05 kernels with the SIMT kernels
(fprop / dgrad / wgrad) on random data and print the relative error."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import _lib as L  # noqa: E402
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402
from medicaldetectiontoolkit_b200.backbone import FPN  # noqa: E402
from medicaldetectiontoolkit_b200.configs import make_cf  # noqa: E402


def main():
    dev = "cuda:0"
    patch = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 32, 16)
    cf = make_cf('retina_unet', 3, patch)
    fpn = FPN(cf, C.NDConvGenerator(3), operate_stride1=True).to(dev)
    seen = {}

    def hook(m, inp, out):
        key = (tuple(inp[0].shape), tuple(m.weight.shape), m.stride, m.padding)
        seen[key] = 1
    for m in fpn.modules():
        if isinstance(m, C.Conv3d):
            m.register_forward_hook(hook)
    with torch.no_grad():
        fpn(torch.rand(1, 1, *patch, device=dev))
    lib = L.load()
    torch.manual_seed(0)
    worst = 0
    for (xs, ws, st, pd) in sorted(seen):
        x = torch.randn(*xs, device=dev).contiguous(memory_format=torch.channels_last_3d)
        w = torch.randn(*ws, device=dev) / np.sqrt(ws[1] * ws[2] * ws[3] * ws[4])
        d = C._desc(xs, ws, st, pd, False, 0, 0)
        algos = [lib.mdt_conv3d_algo(d, ps) for ps in range(3)]
        y1 = C.conv3d_forward(x, w, None, st, pd, algo=1)
        gy = torch.randn_like(y1)
        res = []
        for ps in range(3):
            if algos[ps] != 2:
                res.append("simt")
                continue
            if ps == 0:
                a, b = C.conv3d_forward(x, w, None, st, pd, algo=2), y1
            elif ps == 1:
                a, b = C.conv3d_dgrad(gy, w, xs, st, pd, algo=2), C.conv3d_dgrad(gy, w, xs, st, pd, algo=1)
            else:
                a, b = C.conv3d_wgrad(x, gy, ws, st, pd, False, algo=2)[0], C.conv3d_wgrad(x, gy, ws, st, pd, False, algo=1)[0]
            e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            worst = max(worst, e)
            res.append("%.1e%s" % (e, " <<<<" if e > 1e-4 else ""))
        print("x%-22s w%-24s s%-10s p%-10s fprop %-12s dgrad %-12s wgrad %-12s" % (xs, ws, st, pd, *res))
    print("worst", worst)


if __name__ == "__main__":
    main()
