"""Forward and forward + backward device time of the unary layers (ops.linear: tcgen05 row GEMM of csrc/rowgemm_tc.cu for
y = x W^T and dx = dy W, library GEMM for the weight gradient) against torch.nn.functional.linear (library fp32 GEMMs) on
the layer shapes of the KPConvPaper segmentation network at the C3 / C5 point counts:
    python bench_tools/linear_times.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import ops  # noqa: E402

DEV = "cuda:0"
SHAPES = [(50614, 64, 16), (50614, 32, 64), (50614, 16, 64), (50614, 64, 128), (50614, 192, 64),
          (15039, 128, 32), (15039, 32, 128), (15039, 128, 256), (15039, 384, 128), (2580, 256, 64), (2580, 64, 256),
          (2580, 256, 512), (2580, 768, 256), (464, 512, 1024), (464, 1536, 512), (114, 1024, 1024),
          (50614, 32, 480), (15039, 64, 960), (2580, 128, 1920), (464, 256, 3840), (114, 512, 7680)]


def timed(fn, reps=20):
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
    gen = torch.Generator(device=DEV).manual_seed(0)
    print("%-22s %12s %12s %12s %12s %10s" % ("n, Cin -> Cout", "own fwd", "torch fwd", "own f+b", "torch f+b", "rel diff"))
    tot = [0.0, 0.0, 0.0, 0.0]
    for n, ci, co in SHAPES:
        x = torch.randn((n, ci), device=DEV, generator=gen, requires_grad=True)
        w = (torch.randn((co, ci), device=DEV, generator=gen) / ci ** 0.5).requires_grad_(True)
        dy = torch.randn((n, co), device=DEV, generator=gen)
        with torch.no_grad():
            a, b = ops.linear(x, w), torch.nn.functional.linear(x, w)
        rel = float((a - b).abs().max() / b.abs().max())
        with torch.no_grad():
            f_own = timed(lambda: ops.linear(x, w))
            f_lib = timed(lambda: torch.nn.functional.linear(x, w))
        fb_own = timed(lambda: torch.autograd.grad(ops.linear(x, w), (x, w), dy))
        fb_lib = timed(lambda: torch.autograd.grad(torch.nn.functional.linear(x, w), (x, w), dy))
        for i, v in enumerate((f_own, f_lib, fb_own, fb_lib)):
            tot[i] += v
        print("%6d %5d -> %-5d %9.1f us %9.1f us %9.1f us %9.1f us %10.2e" % (n, ci, co, f_own, f_lib, fb_own, fb_lib, rel))
    print("sums: own fwd %.1f us, library fwd %.1f us; own f+b %.1f us, library f+b %.1f us" % tuple(tot))


if __name__ == "__main__":
    main()
