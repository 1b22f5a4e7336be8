"""Three forward+backward passes of the C3 layer-0b KPConv (32->32, ~50k points) for an ncu launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import kpconv, ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
pos_h, batch_h = synthetic.indoor_spheres()
pos, batch = torch.from_numpy(pos_h).to(dev), torch.from_numpy(batch_h).to(dev)
nb = ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch, return_dist=False)[0]
x = torch.randn((pos.shape[0], 32), device=dev, requires_grad=True)
W = (torch.randn((15, 32, 32), device=dev) * 0.1).requires_grad_(True)
kp = torch.from_numpy(kpconv.load_kernels(0.06, random_rotation=False)).to(dev)
dout = torch.randn((pos.shape[0], 32), device=dev)
for _ in range(3):
    out = ops.KPConv_ops(pos, pos, nb, x, kp, W, 0.04, "linear", "sum")
    torch.autograd.grad(out, (x, W), dout)
torch.cuda.synchronize()
