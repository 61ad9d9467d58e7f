import sys, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from spacap3d_b200 import _ext
from spacap3d_b200.scenes import make_scene_xyz
B = 8
xyz = torch.from_numpy(np.stack([make_scene_xyz(100 + i, 40000) for i in range(B)], 0)).cuda()
_, l1 = _ext.furthest_point_sampling_with_xyz(xyz, 2048)
_, l2 = _ext.furthest_point_sampling_with_xyz(l1, 1024)
idx = _ext.ball_query(l2, l1, 0.4, 32)
g = torch.randn(B, 128, 1024, 32, device="cuda")
for _ in range(3):
    out = _ext.group_points_grad(g, idx, 2048)
torch.cuda.synchronize()
cnt = torch.bincount(idx[0].flatten().long(), minlength=2048)
print("list len: max", int(cnt.max()), "mean", float(cnt.float().mean()), ">32:", int((cnt > 32).sum()), "sum>32", int(cnt[cnt > 32].sum()))
