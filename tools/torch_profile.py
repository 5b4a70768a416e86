#!/usr/bin/env python
"""Top CUDA kernels of one Retina U-Net train step (torch.profiler, device time), to find the non-conv time."""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medicaldetectiontoolkit_b200 import retina_unet  # noqa: E402
from medicaldetectiontoolkit_b200.configs import make_cf, synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
cf = make_cf('retina_unet', 3, (128, 128, 128))
torch.manual_seed(0)
np.random.seed(0)
net = retina_unet.net(cf, None).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
batch = synthetic_batch(cf, 2, seed=0)


def step():
    res = net.train_forward(batch, monitor_anchors=False)
    opt.zero_grad()
    res['torch_loss'].backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))

# ---- where does the GPU idle?  gaps between consecutive device activities (kernels + memcpys) of the profiled step, largest first
try:
    from torch.autograd import DeviceType
    dev_evts = [e for e in prof.events() if e.device_type == DeviceType.CUDA]
    dev_evts.sort(key=lambda e: e.time_range.start)
    gaps, busy_end, t0 = [], None, None
    for e in dev_evts:
        s, t = e.time_range.start, e.time_range.end
        if busy_end is None:
            busy_end, t0 = t, s
            prev = e
            continue
        if s > busy_end:
            gaps.append((s - busy_end, prev.name[:60], e.name[:60], busy_end - t0))
        if t > busy_end:
            busy_end, prev = t, e
    total_gap = sum(g[0] for g in gaps)
    print("\ndevice timeline: span %.2f ms, idle %.2f ms in %d gaps (> 20 us: %.2f ms)" % ((busy_end - t0) / 1e3, total_gap / 1e3, len(gaps),
                                                                                         sum(g[0] for g in gaps if g[0] > 20) / 1e3))
    print("largest gaps: us | at ms | after kernel -> before kernel")
    for g in sorted(gaps, reverse=True)[:25]:
        print("%8.1f | %7.2f | %s -> %s" % (g[0], g[3] / 1e3, g[1], g[2]))
    # idle time per millisecond bucket of the step
    buckets = {}
    for g in gaps:
        buckets[int(g[3] // 2000)] = buckets.get(int(g[3] // 2000), 0.0) + g[0]
    print("idle us per 2-ms bucket of the step:", " ".join("%d:%.0f" % (k * 2, v) for k, v in sorted(buckets.items())))
except Exception as ex:   # the event API differs between torch versions; the table above is the primary output
    print("gap analysis unavailable:", ex)
