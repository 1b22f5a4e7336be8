"""One unary layer of KPConv level 0 (50 614 x 64 -> 128) forward + backward on the tcgen05 row GEMM, three times, for ncu:
    ncu --set full -k regex:rowgemm_tc -c 2 -o out python bench_tools/rowgemm_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import ops  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((50614, 64), device=dev, generator=g, requires_grad=True)
w = (torch.randn((128, 64), device=dev, generator=g) / 8).requires_grad_(True)
dy = torch.randn((50614, 128), device=dev, generator=g)
for _ in range(3):
    torch.autograd.grad(ops.linear(x, w), (x, w), dy)
torch.cuda.synchronize()
