#!/usr/bin/env python
"""Per-layer conv timing of one Retina U-Net train step (CUDA events around every conv call on the launching stream).
python tools/conv_profile.py [--patch 128 128 128] [--batch 2] > gpurun_out/conv_profile.txt"""
import argparse
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import conv as C  # noqa: E402
from medicaldetectiontoolkit_b200 import retina_unet  # noqa: E402
from medicaldetectiontoolkit_b200.configs import make_cf, synthetic_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, nargs=3, default=[128, 128, 128])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--precision", type=int, default=0)
    args = ap.parse_args()
    C.DEFAULT_PRECISION = args.precision
    dev = torch.device("cuda:0")
    cf = make_cf('retina_unet', 3, tuple(args.patch), batch_size=args.batch)
    torch.manual_seed(0)
    np.random.seed(0)
    net = retina_unet.net(cf, None).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    batch = synthetic_batch(cf, args.batch, seed=0)

    def step():
        res = net.train_forward(batch, monitor_anchors=False)
        opt.zero_grad()
        res['torch_loss'].backward()
        opt.step()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    C.EVENT_LOG = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    log = C.EVENT_LOG
    C.EVENT_LOG = None
    agg = defaultdict(lambda: [0.0, 0, 0.0])
    for a, b, tag in log:
        ps, xs, ws, st, algo = tag
        od = [(xs[2 + i] + 0) for i in range(3)]
        mult = 2.0 if ps == 3 else 1.0
        flops = mult * 2.0 * xs[0] * np.prod([xs[2 + i] // st[i] for i in range(3)]) * ws[0] * ws[1] * ws[2] * ws[3] * ws[4]
        key = (("fprop", "dgrad", "wgrad", "bwd")[ps], xs[1], ws[0], ws[2:], st, xs[2:], {2: "TC", 4: "PW"}.get(algo, "SIMT"))
        agg[key][0] += a.elapsed_time(b)
        agg[key][1] += 1
        agg[key][2] += flops
    total = sum(v[0] for v in agg.values())
    print("step %.2f ms, conv calls %d, conv total %.2f ms" % (e0.elapsed_time(e1), len(log), total))
    print("%-6s %4s %4s %-10s %-10s %-16s %-5s %5s %9s %8s %8s" % ("pass", "cin", "cout", "k", "stride", "in-spatial", "algo", "calls", "ms", "share", "TFLOP/s"))
    for key, (ms, n, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print("%-6s %4d %4d %-10s %-10s %-16s %-5s %5d %9.3f %7.1f%% %8.1f" % (key[0], key[1], key[2], "x".join(map(str, key[3])), "x".join(map(str, key[4])),
                                                                         "x".join(map(str, key[5])), key[6], n, ms, 100 * ms / total, fl / ms / 1e9))


if __name__ == "__main__":
    main()
