#!/usr/bin/env python
"""Debug aid: run the unet FPN forward/backward with the SIMT and the tcgen05 conv paths (same weights, same input) and print, per conv
module, the relative difference of its output, its input gradient and its weight gradient."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import detweights  # noqa: E402
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402
from medicaldetectiontoolkit_b200.backbone import FPN  # noqa: E402
from medicaldetectiontoolkit_b200.configs import make_cf  # noqa: E402

dev = "cuda:0"


def run(algo):
    C.DEFAULT_ALGO = algo
    cf = make_cf('retina_unet', 3, (32, 32, 16))
    fpn = detweights.fill_(FPN(cf, C.NDConvGenerator(3), operate_stride1=True)).to(dev)
    rec = {}
    for name, m in fpn.named_modules():
        if isinstance(m, C.Conv3d):
            def fh(mod, inp, out, name=name):
                rec[name + ".out"] = out.detach().clone()
                if out.requires_grad:
                    out.register_hook(lambda g, name=name: rec.__setitem__(name + ".gout", g.detach().clone()))
                if inp[0].requires_grad:
                    inp[0].register_hook(lambda g, name=name: rec.__setitem__(name + ".gin_total", g.detach().clone()))
            m.register_forward_hook(fh)
    x = torch.from_numpy(np.random.RandomState(3).rand(1, 1, 32, 32, 16).astype(np.float32)).to(dev).requires_grad_(True)
    outs = fpn(x)
    sum((o * o).mean() for o in outs).backward()
    rec["x.grad"] = x.grad.detach().clone()
    for k, p in fpn.named_parameters():
        if p.grad is not None:
            rec[k + ".grad"] = p.grad.detach().clone()
    return rec


a, b = run(1), run(0)
rows = []
for k in a:
    if k in b:
        d = float((a[k].double() - b[k].double()).abs().max() / a[k].double().abs().max().clamp_min(1e-30))
        rows.append((d, k, tuple(a[k].shape)))
for d, k, sh in sorted(rows, reverse=True)[:40]:
    print("%.2e  %-32s %s" % (d, k, sh))

# ---- is the in-network divergence a precision effect?  take the actual gradient entering C2.3.conv3 (fp32 run) and evaluate that one
# dgrad with fp64, fp32-SIMT and split-bf16 tcgen05
import torch.nn.functional as F
C.DEFAULT_ALGO = 0
cf = make_cf('retina_unet', 3, (32, 32, 16))
fpn = detweights.fill_(FPN(cf, C.NDConvGenerator(3), operate_stride1=True)).to(dev)
for name in ("C2.3.conv3", "C2.3.conv2.0", "C2.3.conv1.0"):
    mod = dict(fpn.named_modules())[name]
    gy = a[name + ".gout"] if (name + ".gout") in a else None
    if gy is None:
        continue
    w = mod.weight.detach()
    xs = tuple(a[name + ".out"].shape)
    x_shape = (xs[0], w.shape[1]) + xs[2:]
    ref = F.conv_transpose3d(gy.double(), w.double(), stride=mod.stride, padding=mod.padding)
    absref = F.conv_transpose3d(gy.double().abs(), w.double().abs(), stride=mod.stride, padding=mod.padding)
    for algo, label in ((1, "simt"), (2, "tc")):
        dx = C.conv3d_dgrad(gy, w, x_shape, mod.stride, mod.padding, algo=algo)
        print(name, label, "rel-to-max err %.2e" % float((dx.double() - ref).abs().max() / ref.abs().max()),
              " max |terms|/max|result| %.1f" % float(absref.max() / ref.abs().max()), " gy absmax %.3e absmean %.3e" % (float(gy.abs().max()), float(gy.abs().mean())))
