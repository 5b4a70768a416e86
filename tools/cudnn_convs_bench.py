#!/usr/bin/env python
"""Same-GPU comparison arm: the SAME Retina U-Net train step (this package's model, matching, losses, NMS) with every conv
replaced by stock `torch.nn.Conv3d` (cuDNN), fp32 without TF32 (the precision class of the tcgen05 split-bf16 path: ~1e-6) and with
TF32 allowed (~1e-3, below the 1e-4 parity bar) — the ratio that says what the hand-written conv kernels are worth on a B200.
usage: python tools/cudnn_convs_bench.py [--steps 10] [--warmup 3]"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import retina_unet  # noqa: E402
from medicaldetectiontoolkit_b200.configs import make_cf, synthetic_batch  # noqa: E402


class TorchConvGenerator(object):
    """utils/model_utils.py:732-781 for norm=None on stock modules (same nesting / state-dict keys as the reference)"""

    def __init__(self, dim):
        self.dim = dim

    def __call__(self, c_in, c_out, ks, pad=0, stride=1, norm=None, relu='relu'):
        conv = (nn.Conv2d if self.dim == 2 else nn.Conv3d)(c_in, c_out, kernel_size=ks, padding=pad, stride=stride)
        if relu is not None:
            conv = nn.Sequential(conv, nn.ReLU(inplace=True))
        return conv


def run(kind, steps, warmup):
    dev = torch.device("cuda:0")
    cf = make_cf('retina_unet', 3, (128, 128, 128), batch_size=2)
    torch.manual_seed(0)
    np.random.seed(1000)
    saved = retina_unet.NDConvGenerator
    if kind != "mdt":
        retina_unet.NDConvGenerator = TorchConvGenerator
        torch.backends.cudnn.allow_tf32 = kind == "cudnn_tf32"
        torch.backends.cuda.matmul.allow_tf32 = kind == "cudnn_tf32"
        torch.backends.cudnn.benchmark = True
    try:
        net = retina_unet.net(cf, None).to(dev)
        if kind != "mdt":
            net = net.to(memory_format=torch.channels_last_3d)
    finally:
        retina_unet.NDConvGenerator = saved
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    batches = [synthetic_batch(cf, 2, seed=i) for i in range(2)]
    for b in batches:
        b['data'] = torch.from_numpy(b['data']).to(dev)
        b['seg'] = torch.from_numpy(b['seg']).to(dev)

    def step(b):
        res = net.train_forward(b, monitor_anchors=False)
        opt.zero_grad(set_to_none=True)
        res['torch_loss'].backward()
        opt.step()

    for i in range(warmup):
        step(batches[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"convs": kind, "ms_per_step": ms, "patches_per_s": 2000.0 / ms, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--kinds", nargs="*", default=["cudnn_fp32", "cudnn_tf32", "mdt"])
    a = ap.parse_args()
    out = []
    for k in a.kinds:
        torch.cuda.reset_peak_memory_stats()
        try:
            out.append(run(k, a.steps, a.warmup))
        except Exception as ex:   # e.g. out of memory with the library's workspace sizes
            out.append({"convs": k, "error": repr(ex)[:300]})
        torch.cuda.empty_cache()
        print(json.dumps(out[-1]), flush=True)
