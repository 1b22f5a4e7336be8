"""Eager training steps of the KPConvPaper segmentation network (C3 / C5 shape) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python bench_tools/stack_profile.py
With the argument `geometry` the pyramid is rebuilt inside every step (the searches show up in the list)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import segmentation  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(5)
tr = segmentation.KPConvSegmentationTrainer(feat_dim=4, num_classes=13, device=dev)
rebuild = len(sys.argv) > 1 and sys.argv[1] == "geometry"
tr.build_pyramid()
print("points per level:", [m.pos.shape[0] for m in tr.pyramid[0]])
for it in range(3):
    loss = tr.step(rebuild_geometry=rebuild)
torch.cuda.synchronize()
print("loss", loss)
