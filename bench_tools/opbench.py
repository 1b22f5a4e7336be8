"""Per-operator timings (CUDA events, L2-exceeding or rotating buffers where it matters).
Run on the GPU box:  python bench_tools/opbench.py [filter]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_points3d_b200 import _cabi, kpconv, ops, synthetic  # noqa: E402

DEV = "cuda:0"
FILTER = sys.argv[1] if len(sys.argv) > 1 else ""


def timed(name, fn, reps=20, bytes_=None):
    """GPU-side time: the call is captured in a CUDA graph and replayed (no Python/launch overhead)"""
    if FILTER and FILTER not in name:
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    eager = fn
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            keep = eager()  # noqa: F841
        fn = graph.replay
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("   (graph capture failed for %s: %s)" % (name, str(e)[:80]))
        fn = eager
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1000
    extra = "  %.0f GB/s" % (bytes_ / us / 1e3) if bytes_ else ""
    print("%-58s %9.1f us%s" % (name, us, extra), flush=True)


def main():
    B = 16
    xyz = torch.from_numpy(synthetic.surface_clouds(B, 2048, 1002)).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    timed("fps 16x2048->512", lambda: ops.furthest_point_sample(xyz, 512))
    idx1 = ops.furthest_point_sample(xyz, 512)
    pos1 = xyz.gather(1, idx1.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    timed("fps 16x512->128", lambda: ops.furthest_point_sample(pos1, 128))
    idx2 = ops.furthest_point_sample(pos1, 128)
    pos2 = pos1.gather(1, idx2.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    x0 = torch.randn((B, 3, 2048), device=DEV, generator=g, requires_grad=True)
    x1 = torch.randn((B, 320, 512), device=DEV, generator=g, requires_grad=True)
    for (p, q, x, r, k) in [(xyz, pos1, x0, 0.1, 32), (xyz, pos1, x0, 0.2, 64), (xyz, pos1, x0, 0.4, 128),
                            (pos1, pos2, x1, 0.4, 64), (pos1, pos2, x1, 0.8, 128)]:
        tag = "N=%d S=%d r=%.1f K=%d C=%d" % (p.shape[1], q.shape[1], r, k, x.shape[1])
        timed("ball_query " + tag, lambda: ops.ball_query(r, k, p, q, return_dist=False))
        ridx = ops.ball_query(r, k, p, q, return_dist=False)[0]
        C, S = x.shape[1], q.shape[1]
        nbytes = B * (4 * C * p.shape[1] + 8 * S * k + 4 * (C + 3) * S * k)
        timed("group_features fwd " + tag, lambda: ops.group_features(x, p, q, ridx), bytes_=nbytes)
        out = ops.group_features(x, p, q, ridx)
        gout = torch.randn_like(out)
        timed("group_features bwd " + tag,
              lambda: ops._group_bwd(gout, ridx, B, C, p.shape[1], S, k, (C + 3) * S * k, 3),
              bytes_=B * (4 * C * S * k + 8 * S * k + 4 * C * p.shape[1]))
    for (unk, kn, C) in [(pos1, pos2, 256), (xyz, pos1, 128)]:
        tag = "n=%d m=%d C=%d" % (unk.shape[1], kn.shape[1], C)
        timed("three_nn " + tag, lambda: ops.three_nn(unk, kn))
        dist, idx = ops.three_nn(unk, kn)
        w = torch.rand_like(dist)
        f = torch.randn((B, C, kn.shape[1]), device=DEV, generator=g, requires_grad=True)
        timed("three_interpolate fwd " + tag, lambda: ops.three_interpolate(f, idx, w),
              bytes_=B * 4 * C * (kn.shape[1] + unk.shape[1]))
        out = ops.three_interpolate(f, idx, w)
        gout = torch.randn_like(out)
        def ibwd():
            d = torch.empty((B, C, kn.shape[1]), device=DEV)
            nb = _cabi.lib().p3d_three_interp_bwd_ws_bytes(B, C, kn.shape[1], unk.shape[1])
            ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
            ops._call(torch.device(DEV), "p3d_three_interp_bwd_f32", gout.data_ptr(), idx.data_ptr(), w.data_ptr(), B, C,
                      kn.shape[1], unk.shape[1], d.data_ptr(), ws.data_ptr(), nb, 0)
            return d
        timed("three_interpolate bwd " + tag, ibwd, bytes_=B * 4 * C * (kn.shape[1] + unk.shape[1]))

    # C4: LiDAR scans
    if not FILTER or "C4" in FILTER:
        scans = torch.from_numpy(synthetic.lidar_scans(B=8)).to(DEV)
        timed("C4 fps 8x120000->4096", lambda: ops.furthest_point_sample(scans, 4096), reps=3)
        qi = ops.furthest_point_sample(scans, 4096)
        q = scans.gather(1, qi.unsqueeze(-1).expand(-1, -1, 3)).reshape(-1, 3).contiguous()
        x = scans.reshape(-1, 3)
        bx = torch.arange(8, device=DEV).repeat_interleave(120000)
        by = torch.arange(8, device=DEV).repeat_interleave(4096)
        timed("C4 knn k=16 8x(4096 q, 120000 x)", lambda: ops.knn_table(x, q, 16, bx, by, return_dist=False), reps=5)

    # C3: KPConv
    if not FILTER or "C3" in FILTER or "kpconv" in FILTER:
        pos_h, batch_h = synthetic.indoor_spheres()
        pos, batch = torch.from_numpy(pos_h).to(DEV), torch.from_numpy(batch_h).to(DEV)
        n = pos.shape[0]
        timed("C3 ball_query partial n=%d r=0.1 M=34" % n,
              lambda: ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch,
                                     return_dist=False))
        nb = ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch,
                            return_dist=False)[0]
        kp = torch.from_numpy(kpconv.load_kernels(0.06, random_rotation=False)).to(DEV)
        for cin, cout in [(32, 32), (64, 64), (128, 128)]:
            x = torch.randn((n, cin), device=DEV, generator=g, requires_grad=True)
            W = (torch.randn((15, cin, cout), device=DEV, generator=g) * 0.1).requires_grad_(True)
            flops = n * (2 * 15 * 34 * cin + 2 * 15 * cin * cout)
            for prec in ("fp32", "tensor"):
                def fwd():
                    with torch.no_grad():
                        return ops.KPConv_ops(pos, pos, nb, x, kp, W, 0.04, "linear", "sum", precision=prec)
                timed("C3 kpconv fwd %s %d->%d n=%d (%.1f GFLOP)" % (prec, cin, cout, n, flops / 1e9), fwd)
            with torch.no_grad():
                a = ops.KPConv_ops(pos, pos, nb, x, kp, W, 0.04, "linear", "sum", precision="fp32")
                b = ops.KPConv_ops(pos, pos, nb, x, kp, W, 0.04, "linear", "sum", precision="tensor")
            print("      tensor vs fp32: max|diff|/max|out| = %.3e  (identical bits: %s)" % (
                float((a - b).abs().max() / a.abs().max()), bool(torch.equal(a, b))))
            gout = torch.randn((n, cout), device=DEV, generator=g)
            xd, Wd = x.detach(), W.detach()

            def kbwd():
                dfeat, dW = torch.zeros_like(xd), torch.zeros_like(Wd)
                nbytes = _cabi.lib().p3d_kpconv_bwd_ws_bytes(n, 34, 15, cin, cout)
                ws = torch.empty(nbytes // 4, dtype=torch.float32, device=DEV)
                ops._call(torch.device(DEV), "p3d_kpconv_bwd_f32", pos.data_ptr(), pos.data_ptr(), nb.data_ptr(),
                          xd.data_ptr(), kp.data_ptr(), Wd.data_ptr(), gout.data_ptr(), n, n, 34, 15, cin, cout, 0.04,
                          1, 0, dfeat.data_ptr(), dW.data_ptr(), ws.data_ptr(), nbytes)
                return dfeat, dW
            timed("C3 kpconv bwd %d->%d" % (cin, cout), kbwd, reps=5)


if __name__ == "__main__":
    main()
