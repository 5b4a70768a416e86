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

# (sum of "Self CUDA" vs the step time from bench.py gives the GPU idle share; a per-gap analysis over prof.events() took > 5 min here - dropped)
