"""FPS on the C4 shape (4096 of 120 000 points) for timing / ncu:  python bench_tools/fps_profile.py [clouds] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import ops, synthetic  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scans = torch.from_numpy(synthetic.lidar_scans(B=8)[:B]).to(DEV)
g = torch.Generator(device=DEV).manual_seed(0)
shuffled = scans[:, torch.randperm(scans.shape[1], device=DEV, generator=g)].contiguous()
for name, t in (("scan order", scans), ("shuffled", shuffled)):
    ops.furthest_point_sample(t, 4096)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.furthest_point_sample(t, 4096)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("FPS %d x 120000 -> 4096, %s: %.3f ms  (%.3f us per round)" % (B, name, ms, ms * 1e3 / 4095), flush=True)
