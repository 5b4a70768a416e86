#!/usr/bin/env python
"""Loss-side kernels (csrc/loss_ops.cu) next to the torch-op formulation of the same reference formulas, on the cfg2 shapes:
segmentation loss on 2 x 2 x 128^3 logits (forward + backward), SHEM class loss on 1 347 840 anchors x 3 classes (forward + backward).
CUDA-event times (L2 flushed between iterations), launches per call, achieved GB/s on the algorithmic bytes.  -> profiles/r02_loss_bench.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import _lib as L  # noqa: E402
from medicaldetectiontoolkit_b200 import model_utils as mutils  # noqa: E402
from medicaldetectiontoolkit_b200 import native_ops  # noqa: E402
from medicaldetectiontoolkit_b200 import retina_unet as RU  # noqa: E402

DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, reps=10):
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
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    lib = L.load()
    out = {"gpu": torch.cuda.get_device_name(0)}
    torch.manual_seed(0)
    logits = (torch.randn(2, 2, 128, 128, 128, device=DEV) * 2).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    seg = (torch.rand(2, 1, 128, 128, 128, device=DEV) < 0.05).to(torch.uint8)

    def seg_fused():
        d, c = native_ops.seg_loss(logits, seg)
        ((1 - d) * 0.5 + c * 0.5).backward()
        logits.grad = None

    def seg_torch():
        s = seg.long()
        ohe = F.one_hot(s[:, 0], 2).movedim(-1, 1).float()
        d = mutils.batch_dice(F.softmax(logits, dim=1), ohe)
        c = F.cross_entropy(logits, s[:, 0])
        ((1 - d) * 0.5 + c * 0.5).backward()
        logits.grad = None

    vox = 2 * 128 ** 3
    alg = vox * (2 * 4 + 1) * 2 + vox * 2 * 4            # forward + backward reads of logits and labels, backward write of d(logits)
    for name, fn in (("fused", seg_fused), ("torch_ops", seg_torch)):
        c0 = lib.mdt_launch_count()
        us = timed(fn)
        out["seg_loss_2x2x128^3_fwd_bwd_" + name] = {"us": us, "gbs_on_algorithmic_bytes": alg / us / 1e3, "lib_launches_per_call": (lib.mdt_launch_count() - c0) / 12}
    A = 1347840
    cls = (torch.randn(A, 3, device=DEV) * 0.5).requires_grad_(True)
    match = torch.full((A,), -1, dtype=torch.int32, device=DEV)
    match[torch.randperm(A, device=DEV)[: A // 5]] = 0
    pos = torch.sort(torch.randperm(A, device=DEV)[:3])[0]
    match[pos] = 1

    def shem(fused):
        def run():
            RU.FUSED_LOSSES = fused
            try:
                loss, _ = RU.compute_class_loss(match, cls, shem_poolsize=20, max_pos=3, pos_ids=pos)
                loss.backward()
                cls.grad = None
            finally:
                RU.FUSED_LOSSES = True
        return run

    alg2 = A * 16 + A * 12                              # forward: logits + matches once; backward: zero-fill of d(logits)
    for name, fn in (("fused", shem(True)), ("torch_ops", shem(False))):
        c0 = lib.mdt_launch_count()
        us = timed(fn)
        out["shem_class_loss_1.35M_anchors_fwd_bwd_" + name] = {"us": us, "gbs_on_algorithmic_bytes": alg2 / us / 1e3,
                                                                "lib_launches_per_call": (lib.mdt_launch_count() - c0) / 12}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
