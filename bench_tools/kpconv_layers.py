"""Forward and forward+backward device time of KPConv_ops on every layer shape of the KPConvPaper encoder
(conf/models/segmentation/kpconv.yaml:14-57 on the C3 spheres), fused tcgen05 path against the small-point-set path:
    python bench_tools/kpconv_layers.py"""
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


def main():
    pos_h, batch_h = synthetic.indoor_spheres()
    pos, batch = torch.from_numpy(pos_h).to(DEV), torch.from_numpy(batch_h).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    s_pos, s_batch, grid = pos, batch, 0.04
    print("%-34s %10s %10s %12s %12s %10s" % ("layer", "fused fwd", "split fwd", "fused f+b", "split f+b", "rel diff"))
    for L in synthetic.C3_LAYERS:
        if L["cin"] % 32:
            continue
        if L["strided"]:
            q_pos, q_batch = kpconv.grid_subsample(s_pos, s_batch, grid * 2)
        else:
            q_pos, q_batch = s_pos, s_batch
        r = 2.5 * grid
        nb = ops.ball_query(r, L["M"], s_pos, q_pos, mode="partial_dense", batch_x=s_batch, batch_y=q_batch,
                            return_dist=False)[0]
        n, N0 = q_pos.shape[0], s_pos.shape[0]
        x = torch.randn((N0, L["cin"]), device=DEV, generator=g, requires_grad=True)
        W = (torch.randn((15, L["cin"], L["cout"]), device=DEV, generator=g) * 0.05).requires_grad_(True)
        kp = torch.from_numpy(kpconv.load_kernels(1.5 * grid, random_rotation=False)).to(DEV)
        dout = torch.randn((n, L["cout"]), device=DEV, generator=g)

        def fwd(split):
            ops.KPCONV_SPLIT_MAX_POINTS = (1 << 40) if split else 0
            with torch.no_grad():
                return ops.KPConv_ops(q_pos, s_pos, nb, x, kp, W, grid, "linear", "sum")

        def fb(split):
            ops.KPCONV_SPLIT_MAX_POINTS = (1 << 40) if split else 0
            out = ops.KPConv_ops(q_pos, s_pos, nb, x, kp, W, grid, "linear", "sum")
            return torch.autograd.grad(out, (x, W), dout)

        saved_min = ops.KPCONV_SPLIT_MIN_WEIGHTS
        ops.KPCONV_SPLIT_MIN_WEIGHTS = 0
        t = [timed(lambda: fwd(False)), timed(lambda: fwd(True)), timed(lambda: fb(False)), timed(lambda: fb(True))]
        a, b = fwd(False), fwd(True)
        err = float((a - b).abs().max() / b.abs().max())
        ops.KPCONV_SPLIT_MIN_WEIGHTS = saved_min
        print("L%d %4d->%-4d n=%6d N0=%6d M=%2d %8.1f us %8.1f us %10.1f us %10.1f us %10.2e" %
              (L["level"], L["cin"], L["cout"], n, N0, L["M"], t[0], t[1], t[2], t[3], err), flush=True)
        if L["strided"]:
            s_pos, s_batch, grid = q_pos, q_batch, grid * 2
    ops.KPCONV_SPLIT_MAX_POINTS = 0


if __name__ == "__main__":
    main()
