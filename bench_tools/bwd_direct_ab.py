"""Narrow grouping backward (bwd_direct_kernel, the first set-abstraction level of C2: C = 3, N = 2048, S = 512,
K = 32 / 64 / 128 on real ball-query tables): the library in the tree against another build of it, through the C ABI.
Checks that both builds give the same bits and the float64 scatter sum within 1e-5, then times 50 back-to-back
launches of each with CUDA events.
    python bench_tools/bwd_direct_ab.py [other_libp3d_b200.so]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_points3d_b200 import _cabi, ops, synthetic  # noqa: E402

DEV = "cuda:0"


def load(path):
    h = ctypes.CDLL(path)
    for name in ("p3d_group_bwd_ws_bytes", "p3d_group_bwd_f32"):
        fn = getattr(h, name)
        fn.restype, fn.argtypes = _cabi.SIGNATURES[name]
    return h


def run(h, dout, idx, N, ws):
    B, C, S, K = dout.shape
    dfeat = torch.empty((B, C, N), dtype=torch.float32, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    rc = h.p3d_group_bwd_f32(dout.data_ptr(), idx.data_ptr(), B, C, N, S, K, dfeat.data_ptr(), C * S * K, 0,
                             ws.data_ptr(), ws.numel(), 0, st)
    assert rc == 0, rc
    return dfeat


def timed(h, dout, idx, N, ws, reps=50):
    for _ in range(5):
        run(h, dout, idx, N, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run(h, dout, idx, N, ws)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    libs = {"tree": load(_cabi.LIB_PATH)}
    if len(sys.argv) > 1:
        libs["other"] = load(sys.argv[1])
    g = torch.Generator(device=DEV).manual_seed(0)
    out = {}
    cases = [(16, 3, 2048, 512, K, r) for K, r in ((32, 0.1), (64, 0.2), (128, 0.4))] + [(5, 6, 1000, 300, 48, 0.3)]
    for B, C, N, S, K, r in cases:
        xyz = torch.from_numpy(synthetic.surface_clouds(B, N, 1002)).to(DEV)
        _, pos1 = ops.furthest_point_sample(xyz, S, return_xyz=True)
        idx = ops.ball_query(r, K, xyz, pos1, return_dist=False)[0].contiguous()
        dout = torch.randn((B, C, S, K), device=DEV, generator=g)
        ref = torch.zeros((B, C, N), dtype=torch.float64, device=DEV)
        ref.scatter_add_(2, idx.reshape(B, 1, S * K).expand(B, C, S * K), dout.double().reshape(B, C, S * K))
        row = {}
        res = {}
        for name, h in libs.items():
            nbytes = h.p3d_group_bwd_ws_bytes(B, C, N, S, K)
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=DEV)
            res[name] = run(h, dout, idx, N, ws)
            again = run(h, dout, idx, N, ws)
            row[name + "_deterministic"] = bool(torch.equal(res[name], again))
            row[name + "_relerr"] = float((res[name].double() - ref).abs().max() / ref.abs().max())
            row[name + "_us"] = round(timed(h, dout, idx, N, ws), 2)
        if "other" in res:
            row["same_bits"] = bool(torch.equal(res["tree"], res["other"]))
        out["B%d C%d N%d S%d K%d" % (B, C, N, S, K)] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
