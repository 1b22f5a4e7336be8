"""wide grouping backward (C2 SA2, K = 128, C = 320) and three_interpolate forward, three calls each, for ncu"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import ops, synthetic  # noqa: E402

DEV = "cuda:0"
B = 16
xyz = torch.from_numpy(synthetic.surface_clouds(B, 2048, 1002)).to(DEV)
g = torch.Generator(device=DEV).manual_seed(0)
idx1, pos1 = ops.furthest_point_sample(xyz, 512, return_xyz=True)
idx2, pos2 = ops.furthest_point_sample(pos1, 128, return_xyz=True)
x1 = torch.randn((B, 320, 512), device=DEV, generator=g, requires_grad=True)
ridx = ops.ball_query(0.8, 128, pos1, pos2, return_dist=False)[0]
out = ops.group_features(x1, pos1, pos2, ridx)
gout = torch.randn_like(out)
for _ in range(3):
    torch.autograd.grad(out, x1, gout, retain_graph=True)
dist, i3, w3 = ops.three_nn(xyz, pos1, return_weight=True)
f3 = torch.randn((B, 128, 512), device=DEV, generator=g)
for _ in range(3):
    ops.three_interpolate(f3, i3, w3)
torch.cuda.synchronize()
