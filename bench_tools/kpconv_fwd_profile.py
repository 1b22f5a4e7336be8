"""fused KPConv forward on every encoder layer shape (3 calls each, inference mode) for an ncu launch list, or with an
argument the index of ONE layer (for `ncu --set full -k regex:kpconv_fwd_tc`):
    python bench_tools/kpconv_fwd_profile.py [layer-index] [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import kpconv, ops, synthetic  # noqa: E402

DEV = "cuda:0"
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
m_override = int(sys.argv[2]) if len(sys.argv) > 2 else None
ops.KPCONV_SPLIT_MAX_POINTS = 0
pos_h, batch_h = synthetic.indoor_spheres()
s_pos, s_batch, grid = torch.from_numpy(pos_h).to(DEV), torch.from_numpy(batch_h).to(DEV), 0.04
g = torch.Generator(device=DEV).manual_seed(0)
li = -1
for L in synthetic.C3_LAYERS:
    if L["strided"]:
        q_pos, q_batch = kpconv.grid_subsample(s_pos, s_batch, grid * 2)
    else:
        q_pos, q_batch = s_pos, s_batch
    if L["cin"] % 32 == 0:
        li += 1
        if only is None or only == li:
            M = m_override or L["M"]
            nb = ops.ball_query(2.5 * grid, M, s_pos, q_pos, mode="partial_dense", batch_x=s_batch, batch_y=q_batch,
                                return_dist=False)[0]
            x = torch.randn((s_pos.shape[0], L["cin"]), device=DEV, generator=g)
            W = torch.randn((15, L["cin"], L["cout"]), device=DEV, generator=g) * 0.05
            kp = torch.from_numpy(kpconv.load_kernels(1.5 * grid, random_rotation=False)).to(DEV)
            with torch.no_grad():
                for _ in range(3):
                    ops.KPConv_ops(q_pos, s_pos, nb, x, kp, W, grid, "linear", "sum")
            torch.cuda.synchronize()
            print("layer %d: %d->%d n=%d M=%d" % (li, L["cin"], L["cout"], q_pos.shape[0], M), flush=True)
    if L["strided"]:
        s_pos, s_batch, grid = q_pos, q_batch, grid * 2
