#!/usr/bin/env python
"""One launch of every hot-path kernel at its BASELINE.json shape, bracketed by cudaProfilerStart/Stop, for a single
`ncu --set full --profile-from-start off` capture (numbers printed by a run under ncu are not bench values):

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/ops python tools/ncu_ops.py

Ops: 3D NMS (cfg4 100k boxes + the RPN shape), 3D RoIAlign fwd/bwd (cfg3 P2 7x7x3, channels-last), anchor matching (cfg2 1.35 M x 8),
conv fprop / fused backward (dgrad + wgrad + bias) of the 36->36 and 18->18 3x3x3 layers at 2x128^3, the stem 1->18 and the 7x7x7 layer."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _oracle as O  # noqa: E402  (input generators only)
from golden_cfg import cf3d, rand_gt  # noqa: E402
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402
from medicaldetectiontoolkit_b200 import model_utils as MU  # noqa: E402
from medicaldetectiontoolkit_b200 import native_ops as NO  # noqa: E402

DEV = "cuda:0"


def main():
    jobs = []
    # NMS
    for n, thr, rounded in [(100000, 0.5, True), (6000, 0.7, False)]:
        t = torch.from_numpy(O.synth_boxes(n, 3, seed=n, rounded=rounded)).to(DEV)
        jobs.append(lambda t=t, thr=thr: NO.nms_sorted(t, thr, 3))
    # RoIAlign
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.randn(2, 36, 32, 32, 128).astype(np.float32)).to(DEV).contiguous(memory_format=torch.channels_last_3d)
    boxes, ind = O.synth_rois(1024, 3, 2, seed=8)
    tb, ti = torch.from_numpy(boxes).to(DEV), torch.from_numpy(ind).to(DEV)
    fn = NO.CropAndResizeFunction(7, 7, 3, 0)
    xg = x.clone().requires_grad_(True)
    g = torch.randn_like(fn(xg, tb, ti))

    def roi():
        y = fn(xg, tb, ti)
        torch.autograd.grad(y, xg, g)
    jobs.append(roi)
    # pyramid RoIAlign: BASELINE cfg3 (four FPN levels of a 128^3 patch, 2 x 512 proposals, pool 7x7x3), one launch forward + one backward
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import golden_inputs as GI
    fm, rois = GI.pyramid_inputs("cfg3")
    fmt = [torch.from_numpy(f).to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True) for f in fm]
    rt = torch.from_numpy(rois).to(DEV)
    hh, ww = rt[:, 2] - rt[:, 0], rt[:, 3] - rt[:, 1]
    lvl = (4 + torch.log2(torch.sqrt(hh * ww))).round().int().clamp(0, 3)
    gp = torch.randn(1024, 36, 7, 7, 3, device=DEV).contiguous(memory_format=torch.channels_last_3d)

    def pyr():
        y = NO.pyramid_roi_align(fmt, rt[:, :6].contiguous(), rt[:, 6].int(), lvl, (7, 7, 3))
        torch.autograd.grad(y, fmt, gp)
    jobs.append(pyr)
    # matching
    anchors = MU.generate_pyramid_anchors(None, cf3d((128, 128, 128)))
    gt = rand_gt(np.random.RandomState(8), 8, (128, 128, 128), 3, 4, 48).astype(np.float64)
    cls = np.random.RandomState(1).randint(1, 3, size=8).astype(np.int32)
    a, g_, c_ = torch.from_numpy(anchors).to(DEV), torch.from_numpy(gt).to(DEV), torch.from_numpy(cls).to(DEV)
    jobs.append(lambda: MU.anchor_match_device(a, g_, c_, 3, 0.01, 0.5))
    # conv layers
    for cin, cout, k, st, pad, sp in [(36, 36, 3, (1, 1, 1), 1, (128, 128, 128)), (18, 18, 3, (1, 1, 1), 1, (128, 128, 128)), (1, 18, 3, (1, 1, 1), 1, (128, 128, 128)),
                                      (18, 18, 7, (2, 2, 1), 3, (128, 128, 128)), (64, 64, 3, (1, 1, 1), 1, (32, 32, 128))]:
        k3, p3 = C._triple(k), C._triple(pad)
        xc = torch.randn(2, cin, *sp, device=DEV).contiguous(memory_format=torch.channels_last_3d)
        w = torch.randn(cout, cin, *k3, device=DEV) * 0.05
        b = torch.zeros(cout, device=DEV)
        y = C.conv3d_forward(xc, w, b, st, p3, relu=True)
        gy = torch.randn_like(y)

        def conv(xc=xc, w=w, b=b, st=st, p3=p3, y=y, gy=gy):
            C.conv3d_forward(xc, w, b, st, p3, relu=True)
            need_dx = xc.shape[1] > 1
            if C.conv3d_backward(xc, gy, y, w, st, p3, need_dx, True, False) is None:    # stem: no fused path
                C.conv3d_wgrad(xc, gy, tuple(w.shape), st, p3, True)
                if need_dx:
                    C.conv3d_dgrad(gy, w, tuple(xc.shape), st, p3)
        jobs.append(conv)
    for j in jobs:   # warm-up: plans, tensor maps, allocator
        j()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for j in jobs:
        j()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ncu_ops: done")


if __name__ == "__main__":
    main()
