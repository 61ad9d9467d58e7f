#!/usr/bin/env python
"""torch.profiler table of three training steps (4 scenes x 40 k points, forward + backward): where a step's GPU time goes."""
import sys, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools')]
import bench_train
from spacap3d_b200.scenes import make_scene
dev = torch.device("cuda", 0)
model = bench_train.build_model(1, dev)
pc = torch.from_numpy(np.stack([make_scene(5000 + i, 40000) for i in range(4)], 0)).to(dev)
for _ in range(3):
    model.zero_grad(set_to_none=True); bench_train.step(model, pc, 1)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        model.zero_grad(set_to_none=True); bench_train.step(model, pc, 1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=90))
