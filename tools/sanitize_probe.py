#!/usr/bin/env python
"""Small invocations of the kernels with cross-thread / cross-CTA communication, for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
Covers fps_cluster_kernel (DSMEM st.async + mbarrier exchange), fps_bucket_kernel (named barriers, per-warp queues),
fps_morton_sort_kernel, sa_inline_kernel / sa_fused_pipe_kernel (mbarrier pipelines, tcgen05 / TMEM), pm_linear_kernel
(TMA + tcgen05, also looping over tiles with the double-buffered accumulator), group_points_grad (list build with atomic
cursors + point-owned gather),
group_points_kernel (bulk TMA staging), three_interpolate_kernel, bqg_* (grid ball query).  Sizes are small because the
tools slow execution down ~100x; every result is also checked against the CPU oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from spacap3d_b200 import _ext  # noqa: E402
from spacap3d_b200.pointnet2_modules import PointnetSAModuleVotes  # noqa: E402
from spacap3d_b200.scenes import make_scene_xyz  # noqa: E402

dev = "cuda:0"
xyz_np = np.stack([make_scene_xyz(11, 9000), make_scene_xyz(12, 9000, with_replacement=True)], 0)
xyz = torch.from_numpy(xyz_np).to(dev)
want = oracle.furthest_point_sampling(xyz_np, 96)
for algo in (_ext.FPS_CLUSTER, _ext.FPS_BUCKET):
    with _ext.launch_options(fps_algo=algo):
        idx, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, 96)
    assert np.array_equal(idx.cpu().numpy(), want), algo
bq = _ext.ball_query(new_xyz, xyz, 0.4, 16)
assert np.array_equal(bq.cpu().numpy(), oracle.ball_query(new_xyz.cpu().numpy(), xyz_np, 0.4, 16))
feats = torch.randn(2, 12, 9000, device=dev)
grouped = _ext.group_points(feats, bq)
assert np.array_equal(grouped.cpu().numpy(), oracle.group_points(feats.cpu().numpy(), bq.cpu().numpy()))
d2, i3 = _ext.three_nn(xyz[:, :256].contiguous(), new_xyz)
w = torch.rand(2, 256, 3, device=dev)
up = _ext.three_interpolate(torch.randn(2, 20, 96, device=dev), i3, w)
assert torch.isfinite(up).all()
torch.manual_seed(0)
sa = PointnetSAModuleVotes(npoint=128, radius=0.4, nsample=16, mlp=[128, 128, 128, 256], use_xyz=True,
                           normalize_xyz=True).to(dev).eval()
sa1 = PointnetSAModuleVotes(npoint=128, radius=0.3, nsample=64, mlp=[1, 64, 64, 128], use_xyz=True,
                            normalize_xyz=True).to(dev).eval()
with torch.no_grad():
    x1, f1, _ = sa1(xyz[:, :4096].contiguous(), torch.randn(2, 1, 4096, device=dev))
    x2, f2, _ = sa(x1, f1)
assert torch.isfinite(f2).all()
X = _ext.split_half(torch.randn(256, 128, device=dev))
W = _ext.split_half(torch.randn(64, 128, device=dev) / 11.0)
hi, lo = _ext.pm_linear(X, W, torch.zeros(64, device=dev), _ext.PM_HIDDEN, 256)
ref = torch.relu((X[0].double() + X[1].double()) @ (W[0].double() + W[1].double()).t())
assert ((hi.double() + lo.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
with _ext.launch_options(pm_n_tile=64, pm_tiles_per_cta=2):          # two row tiles per CTA: accumulator ring
    hi2, lo2 = _ext.pm_linear(X, W, torch.zeros(64, device=dev), _ext.PM_HIDDEN, 256)
assert torch.equal(hi, hi2) and torch.equal(lo, lo2)
g = torch.randn(2, 12, 96, 16, device=dev)                          # N = 9000 > 8192: the large-cloud backward
gg = _ext.group_points_grad(g, bq, 9000)
want_gg = oracle.group_points_grad(g.cpu().numpy(), bq.cpu().numpy(), 9000)
assert np.allclose(gg.cpu().numpy(), want_gg, rtol=1e-5, atol=1e-5)
torch.cuda.synchronize()
print("sanitize probe OK")
