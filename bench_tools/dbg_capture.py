import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_points3d_b200 import ops, synthetic, _cabi
DEV = "cuda:0"
B = 16
xyz = torch.from_numpy(synthetic.surface_clouds(B, 2048, 1002)).to(DEV)
idx1 = ops.furthest_point_sample(xyz, 512)
pos1 = xyz.gather(1, idx1.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
ridx = ops.ball_query(0.1, 32, xyz, pos1, return_dist=False)[0]
gout = torch.randn((B, 6, 512, 32), device=DEV)
def direct():
    return ops._group_bwd(gout, ridx, B, 3, 2048, 512, 32, 6 * 512 * 32, 3)
for _ in range(3):
    direct()
torch.cuda.synchronize()
print("eager ok")
for mode in ["direct"]:
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            r = direct()
        g.replay(); torch.cuda.synchronize()
        print(mode, "capture ok", float(r.sum()))
    except Exception as e:
        print(mode, "capture FAILED:", str(e)[:300])
        print("p3d last error:", _cabi.lib().p3d_last_error())
