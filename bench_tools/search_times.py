"""cell-list vs sweep kernels on the C3 / C4 search shapes (CUDA-graph replay, us per call):
    python bench_tools/search_times.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import kpconv, ops, synthetic  # noqa: E402

DEV = "cuda:0"


def timed(fn, reps=10):
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
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def both(name, fn):
    ops.GRID_SEARCH_MIN_POINTS = 0
    a = timed(fn)
    ops.GRID_SEARCH_MIN_POINTS = 1 << 40
    b = timed(fn)
    ops.GRID_SEARCH_MIN_POINTS = 4096
    print("%-52s cell list %9.1f us   sweep %9.1f us" % (name, a, b), flush=True)


def main():
    pos_h, batch_h = synthetic.indoor_spheres()
    pos, batch = torch.from_numpy(pos_h).to(DEV), torch.from_numpy(batch_h).to(DEV)
    s_pos, s_batch = pos, batch
    for lv, (grid, r, M) in enumerate([(0.04, 0.1, 34), (0.08, 0.2, 34), (0.16, 0.4, 34), (0.32, 0.8, 34)]):
        both("ball query level %d: n=%d r=%.2f M=%d" % (lv, s_pos.shape[0], r, M),
             lambda: ops.ball_query(r, M, s_pos, s_pos, mode="partial_dense", batch_x=s_batch, batch_y=s_batch,
                                    return_dist=False))
        q_pos, q_batch = kpconv.grid_subsample(s_pos, s_batch, grid * 2)
        both("  strided query n=%d of %d" % (q_pos.shape[0], s_pos.shape[0]),
             lambda: ops.ball_query(r, M, s_pos, q_pos, mode="partial_dense", batch_x=s_batch, batch_y=q_batch,
                                    return_dist=False))
        both("  upsample knn k=1: %d queries among %d" % (s_pos.shape[0], q_pos.shape[0]),
             lambda: ops.knn_table(q_pos, s_pos, 1, q_batch, s_batch, return_dist=False))
        s_pos, s_batch = q_pos, q_batch
    scans = torch.from_numpy(synthetic.lidar_scans(B=8)).to(DEV)
    qi = ops.furthest_point_sample(scans, 4096)
    q = scans.gather(1, qi.unsqueeze(-1).expand(-1, -1, 3)).reshape(-1, 3).contiguous()
    x = scans.reshape(-1, 3)
    bx = torch.arange(8, device=DEV).repeat_interleave(120000)
    by = torch.arange(8, device=DEV).repeat_interleave(4096)
    both("C4 knn k=16: 8 x 4096 queries among 8 x 120000", lambda: ops.knn_table(x, q, 16, bx, by, return_dist=False))
    both("C4 ball query r=0.5 M=16 (same queries)",
         lambda: ops.ball_query(0.5, 16, x, q, mode="partial_dense", batch_x=bx, batch_y=by, return_dist=False))


if __name__ == "__main__":
    main()
