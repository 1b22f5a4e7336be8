"""Device time of the wide gather / scatter kernels on the C2 shapes (PointNet++ MSG, B = 16): grouping backward at
SA2 (K = 128, C = 320), three_interpolate forward / backward at the two FP levels; algorithmic bytes / time against
the measured HBM peak.  L2 is flushed between repetitions (a 256 MB write).
    python bench_tools/gather_times.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import ops, synthetic  # noqa: E402

DEV = "cuda:0"
PEAK = 6551.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbps"]
except Exception:
    pass


def timed(fn, reps=10):
    """median device time of one CUDA-graph replay of fn (no host launch gaps), L2 flushed before each replay"""
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=DEV)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()  # noqa: F841
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = 16
    xyz = torch.from_numpy(synthetic.surface_clouds(B, 2048, 1002)).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    idx1, pos1 = ops.furthest_point_sample(xyz, 512, return_xyz=True)
    idx2, pos2 = ops.furthest_point_sample(pos1, 128, return_xyz=True)
    rows = []
    for K, r in ((128, 0.8), (64, 0.4)):
        x1 = torch.randn((B, 320, 512), device=DEV, generator=g, requires_grad=True)
        ridx = ops.ball_query(r, K, pos1, pos2, return_dist=False)[0]
        gout = torch.randn((B, 320, 128, K), device=DEV, generator=g)
        by = (320 * 128 * K + 320 * 512) * B * 4
        tf = timed(lambda: ops.grouping_operation(x1.detach(), ridx))
        rows.append(("group fwd SA2 K=%d C=320" % K, tf, by))
        # the backward waits on the inverse index built beside the forward: both in one graph, forward subtracted
        tfb = timed(lambda: torch.autograd.grad(ops.grouping_operation(x1, ridx), x1, gout))
        rows.append(("group bwd SA2 K=%d C=320 (fwd+bwd - fwd)" % K, tfb - tf, by))
    for C, m, n, known, unknown in ((128, 512, 2048, pos1, xyz), (256, 128, 512, pos2, pos1)):
        dist, i3, w3 = ops.three_nn(unknown, known, return_weight=True)
        f3 = torch.randn((B, C, m), device=DEV, generator=g, requires_grad=True)
        g3 = torch.randn((B, C, n), device=DEV, generator=g)
        by = (C * n + C * m) * B * 4 + n * 3 * 12 * B
        tf = timed(lambda: ops.three_interpolate(f3.detach(), i3, w3))
        rows.append(("three_interpolate fwd C=%d %d->%d" % (C, m, n), tf, by))
        tfb = timed(lambda: torch.autograd.grad(ops.three_interpolate(f3, i3, w3), f3, g3))
        rows.append(("three_interpolate bwd C=%d %d->%d (fwd+bwd - fwd)" % (C, m, n), tfb - tf, by))
    for name, t, by in rows:
        print("%-58s %8.1f us  %7.1f MB  %6.0f GB/s  %.2f of HBM peak" % (name, t, by / 1e6, by / t / 1e3, by / t / 1e3 / PEAK))


if __name__ == "__main__":
    main()
