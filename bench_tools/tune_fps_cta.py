import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from torch_points3d_b200 import ops, synthetic
B, N, S = %d, %d, %d
xyz = torch.from_numpy(synthetic.surface_clouds(B, N, 1)).cuda()
for _ in range(3): ops.furthest_point_sample(xyz, S)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    r = ops.furthest_point_sample(xyz, S)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print("%%.1f" %% (e0.elapsed_time(e1) / 20 * 1000))
'''
def run(B, N, S, force):
    env = dict(os.environ)
    if force: env["P3D_FPS_CTA"] = force
    out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, B, N, S)], env=env, capture_output=True, text=True)
    return out.stdout.strip() or out.stderr.strip()[-80:]
for (B, N, S, vs) in [(16, 2048, 512, ["", "128,16", "256,8", "512,4", "256,16", "1024,2"]),
                      (16, 1024, 512, ["", "64,16", "128,8", "256,4", "512,2"]),
                      (16, 512, 128, ["", "64,8", "128,4", "256,4"]),
                      (8, 4096, 1024, ["", "256,16", "512,8", "1024,4"]),
                      (4, 8192, 1024, ["", "512,16", "1024,8"])]:
    print("B=%d N=%d S=%d us:" % (B, N, S), " ".join("%s=%s" % (v or "default", run(B, N, S, v)) for v in vs), flush=True)
