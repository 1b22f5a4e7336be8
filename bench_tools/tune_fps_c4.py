"""Cluster FPS variants at the C4 shape (120k points) and the C2 shapes with small clusters."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from torch_points3d_b200 import ops, synthetic
B, N, S = %d, %d, %d
xyz = torch.from_numpy(synthetic.surface_clouds(B, N, 1)).cuda()
for _ in range(2): ops.furthest_point_sample(xyz, S)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ops.furthest_point_sample(xyz, S)
e1.record(); torch.cuda.synchronize()
print("%%.1f" %% (e0.elapsed_time(e1) / 3 * 1000))
'''
def run(B, N, S, force):
    env = dict(os.environ)
    if force: env["P3D_FPS_FORCE"] = force
    out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, B, N, S)], env=env, capture_output=True, text=True)
    return out.stdout.strip() or out.stderr.strip()[-100:]
for B in (1, 4, 8):
    row = ["%s=%s" % (f or "default", run(B, 120000, 512, f)) for f in
           ("", "1024,8,16", "512,16,16", "256,32,16", "1024,16,8", "512,32,8")]
    print("B=%d N=120000 S=512 us:" % B, " ".join(row), flush=True)
for (B, N, S) in [(16, 2048, 512), (16, 1024, 512)]:
    row = ["%s=%s" % (f or "default", run(B, N, S, f)) for f in
           ("", "256,4,2", "128,8,2", "256,4,4", "128,4,4", "128,8,4", "256,8,4", "256,8,8")]
    print("B=%d N=%d S=%d us:" % (B, N, S), " ".join(row), flush=True)
