"""Tuning sweep for the single-CTA FPS variants (threads x points-per-thread).
Run on the GPU box: python bench_tools/tune_fps.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from torch_points3d_b200 import ops, synthetic
B, N, S = %d, %d, %d
xyz = torch.from_numpy(synthetic.surface_clouds(B, N, 1)).cuda()
for _ in range(3): ops.furthest_point_sample(xyz, S)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.furthest_point_sample(xyz, S)
e1.record(); torch.cuda.synchronize()
print("%%.2f" %% (e0.elapsed_time(e1) / 20 * 1000))
'''

for (B, N, S) in [(16, 2048, 512), (16, 1024, 512), (16, 512, 128), (8, 4096, 1024)]:
    row = []
    for T, P in [(0, 0), (32, 16), (32, 32), (64, 8), (64, 16), (64, 32), (128, 4), (128, 8), (128, 16), (256, 2),
                 (256, 4), (256, 8), (512, 2), (512, 4), (1024, 2), (1024, 4)]:
        if T and T * P < N:
            continue
        env = dict(os.environ)
        if T:
            env["P3D_FPS_FORCE"] = "%d,%d" % (T, P)
        out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, B, N, S)], env=env, capture_output=True, text=True)
        row.append("%s=%s" % ("default" if not T else "%dx%d" % (T, P), out.stdout.strip() or out.stderr[-80:]))
    print("B=%d N=%d S=%d us:" % (B, N, S), " ".join(row), flush=True)
