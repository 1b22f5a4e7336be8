"""Three eager training steps of the KPConv encoder stack with a precomputed pyramid (for an ncu launch list):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python bench_tools/stack_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import kpconv, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
pos_h, batch_h = synthetic.indoor_spheres(seed=1003)
pos, batch = torch.from_numpy(pos_h).to(dev), torch.from_numpy(batch_h).to(dev)
torch.manual_seed(5)
enc = kpconv.KPConvEncoder(5, max_nn_cap=34).to(dev)
head = torch.nn.Linear(enc.out_channels, 13).to(dev)
opt = torch.optim.SGD(list(enc.parameters()) + list(head.parameters()), lr=0.01, momentum=0.98)
feats = torch.cat([torch.ones(pos.shape[0], 1, device=dev), torch.randn(pos.shape[0], 4, device=dev)], 1)
ms = kpconv.MultiScaleTransform(enc.strategies())(pos, batch)
print("points per level:", [m.pos.shape[0] for m in ms])
for it in range(3):
    opt.zero_grad(set_to_none=True)
    logits = head(enc(pos, feats, batch, precomputed=ms).x)
    loss = torch.nn.functional.cross_entropy(logits, torch.arange(logits.shape[0], device=dev) % 13)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
