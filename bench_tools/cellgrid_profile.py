"""three calls each of the cell-list searches on the C3 / C4 shapes, for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python bench_tools/cellgrid_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import ops, synthetic  # noqa: E402

DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
ops.GRID_SEARCH_MIN_POINTS = 0
if which in ("all", "c3"):
    pos_h, batch_h = synthetic.indoor_spheres()
    pos, batch = torch.from_numpy(pos_h).to(DEV), torch.from_numpy(batch_h).to(DEV)
    for _ in range(3):
        ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch, return_dist=False)
        ops.knn_table(pos[::4].contiguous(), pos, 1, batch[::4].contiguous(), batch, return_dist=False)
if which in ("all", "c4"):
    scans = torch.from_numpy(synthetic.lidar_scans(B=8)).to(DEV)
    qi = ops.furthest_point_sample(scans, 4096)
    q = scans.gather(1, qi.unsqueeze(-1).expand(-1, -1, 3)).reshape(-1, 3).contiguous()
    x = scans.reshape(-1, 3)
    bx = torch.arange(8, device=DEV).repeat_interleave(120000)
    by = torch.arange(8, device=DEV).repeat_interleave(4096)
    for _ in range(3):
        ops.knn_table(x, q, 16, bx, by, return_dist=False)
torch.cuda.synchronize()
