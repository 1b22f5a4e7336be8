"""CPU ORACLE -- test infrastructure, never a product path.

numpy front-end over oracle/p3d_oracle.c (see that file's header for the
reference citations and the pinned / unpinned status of every operator).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None

_F = ctypes.POINTER(ctypes.c_float)
_L = ctypes.POINTER(ctypes.c_int64)


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_radius_csr.restype = ctypes.c_int64
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_F)


def _l(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_L)


def batch_to_ptr(batch, num_clouds=None):
    """sorted batch vector -> CSR ptr [C+1] (int64)."""
    batch = np.asarray(batch, dtype=np.int64)
    if num_clouds is None:
        num_clouds = int(batch.max()) + 1 if batch.size else 0
    counts = np.bincount(batch, minlength=num_clouds).astype(np.int64)
    ptr = np.zeros(num_clouds + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr


# ------------------------------------------------------------------ FPS ----
def fps_dense(xyz, npoint):
    xyz, p = _f(xyz)
    B, N, _ = xyz.shape
    if npoint > N:
        raise ValueError("npoint > N")
    idx = np.zeros((B, npoint), dtype=np.int64)
    lib().orc_fps_dense(p, B, N, npoint, idx.ctypes.data_as(_L))
    return idx


def fps_counts(ptr, ratio):
    """per-cloud sample count = ceil(ratio * N_c) (torch_cluster 1.5.x)."""
    n = np.diff(np.asarray(ptr, dtype=np.int64))
    return np.ceil(np.float32(ratio) * n.astype(np.float32)).astype(np.int64)


def fps_csr(xyz, ptr, counts, start=None):
    xyz, p = _f(xyz)
    ptr, pp = _l(ptr)
    C = ptr.shape[0] - 1
    out_ptr = np.zeros(C + 1, dtype=np.int64)
    np.cumsum(np.asarray(counts, dtype=np.int64), out=out_ptr[1:])
    idx = np.zeros(int(out_ptr[-1]), dtype=np.int64)
    ps = None
    if start is not None:
        start, ps = _l(start)
    lib().orc_fps_csr(p, pp, C, out_ptr.ctypes.data_as(_L), idx.ctypes.data_as(_L), ps)
    return idx


# ----------------------------------------------------------- ball query ----
def ball_query_dense(radius, nsample, x, y, strict=True):
    x, px = _f(x)
    y, py = _f(y)
    B, N, _ = x.shape
    S = y.shape[1]
    idx = np.zeros((B, S, nsample), dtype=np.int64)
    d2 = np.zeros((B, S, nsample), dtype=np.float32)
    lib().orc_ball_query_dense(px, py, B, N, S, ctypes.c_float(radius), nsample, int(strict),
                               idx.ctypes.data_as(_L), d2.ctypes.data_as(_F))
    return idx, d2


def ball_query_csr(radius, nsample, x, y, ptr_x, ptr_y, strict=False, shadow=None):
    x, px = _f(x)
    y, py = _f(y)
    ptr_x, ppx = _l(ptr_x)
    ptr_y, ppy = _l(ptr_y)
    C = ptr_x.shape[0] - 1
    n = y.shape[0]
    if shadow is None:
        shadow = x.shape[0]
    idx = np.zeros((n, nsample), dtype=np.int64)
    d2 = np.zeros((n, nsample), dtype=np.float32)
    lib().orc_ball_query_csr(px, py, ppx, ppy, C, ctypes.c_float(radius), nsample, int(strict),
                             ctypes.c_int64(shadow), idx.ctypes.data_as(_L), d2.ctypes.data_as(_F))
    return idx, d2


def radius_csr(x, y, r, ptr_x, ptr_y, max_num_neighbors=32, strict=False):
    x, px = _f(x)
    y, py = _f(y)
    ptr_x, ppx = _l(ptr_x)
    ptr_y, ppy = _l(ptr_y)
    C = ptr_x.shape[0] - 1
    cap = y.shape[0] * max_num_neighbors
    row = np.zeros(cap, dtype=np.int64)
    col = np.zeros(cap, dtype=np.int64)
    e = lib().orc_radius_csr(px, py, ppx, ppy, C, ctypes.c_float(r), max_num_neighbors, int(strict),
                             row.ctypes.data_as(_L), col.ctypes.data_as(_L))
    return np.stack([row[:e], col[:e]])


# ------------------------------------------------------------- grouping ----
def group_fwd(feat, idx):
    feat, pf = _f(feat)
    idx, pi = _l(idx)
    B, C, N = feat.shape
    _, S, K = idx.shape
    out = np.zeros((B, C, S, K), dtype=np.float32)
    lib().orc_group_fwd(pf, pi, B, C, N, S, K, out.ctypes.data_as(_F))
    return out


def group_bwd(dout, idx, N):
    dout, pg = _f(dout)
    idx, pi = _l(idx)
    B, C, S, K = dout.shape
    d = np.zeros((B, C, N), dtype=np.float32)
    lib().orc_group_bwd(pg, pi, B, C, N, S, K, d.ctypes.data_as(_F))
    return d


# ------------------------------------------------ three_nn / interpolate ----
def three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    if m < 3:
        raise ValueError("three_nn needs at least 3 known points")
    dist = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int64)
    lib().orc_three_nn(pu, pk, B, n, m, dist.ctypes.data_as(_F), idx.ctypes.data_as(_L))
    return dist, idx


def three_interpolate_fwd(feat, idx, weight):
    feat, pf = _f(feat)
    idx, pi = _l(idx)
    weight, pw = _f(weight)
    B, C, m = feat.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().orc_three_interp_fwd(pf, pi, pw, B, C, m, n, out.ctypes.data_as(_F))
    return out


def three_interpolate_bwd(dout, idx, weight, m):
    dout, pg = _f(dout)
    idx, pi = _l(idx)
    weight, pw = _f(weight)
    B, C, n = dout.shape
    d = np.zeros((B, C, m), dtype=np.float32)
    lib().orc_three_interp_bwd(pg, pi, pw, B, C, m, n, d.ctypes.data_as(_F))
    return d


# ------------------------------------------------------------------ knn ----
def knn_csr(x, y, k, ptr_x, ptr_y):
    """returns (col [My,k] int64 with -1 padding, d2 [My,k])."""
    x, px = _f(x)
    y, py = _f(y)
    ptr_x, ppx = _l(ptr_x)
    ptr_y, ppy = _l(ptr_y)
    C = ptr_x.shape[0] - 1
    My = y.shape[0]
    col = np.zeros((My, k), dtype=np.int64)
    d2 = np.zeros((My, k), dtype=np.float32)
    lib().orc_knn_csr(px, py, ppx, ppy, C, k, col.ctypes.data_as(_L), d2.ctypes.data_as(_F))
    return col, d2


def knn_edges(x, y, k, ptr_x, ptr_y):
    """torch_cluster.knn layout: [2, E] with row = y index, col = x index."""
    col, _ = knn_csr(x, y, k, ptr_x, ptr_y)
    row = np.repeat(np.arange(col.shape[0], dtype=np.int64), k).reshape(col.shape)
    keep = col >= 0
    return np.stack([row[keep], col[keep]])


# --------------------------------------------------------------- KPConv ----
INFLUENCE = {"constant": 0, "linear": 1, "gaussian": 2}
AGGREGATION = {"sum": 0, "closest": 1}


def kpconv_fwd(q, s, nb, x, Kp, W, extent, influence="linear", aggregation="sum"):
    q, pq = _f(q)
    s, ps = _f(s)
    nb, pn = _l(nb)
    x, px = _f(x)
    Kp, pk = _f(Kp)
    W, pw = _f(W)
    n, M = nb.shape
    N0, Cin = x.shape
    K, _, Cout = W.shape
    out = np.zeros((n, Cout), dtype=np.float32)
    lib().orc_kpconv_fwd(pq, ps, pn, px, pk, pw, ctypes.c_int64(n), ctypes.c_int64(N0), M, K, Cin, Cout,
                         ctypes.c_float(extent), INFLUENCE[influence], AGGREGATION[aggregation],
                         out.ctypes.data_as(_F))
    return out


def kpconv_bwd(q, s, nb, x, Kp, W, dout, extent, influence="linear", aggregation="sum"):
    q, pq = _f(q)
    s, ps = _f(s)
    nb, pn = _l(nb)
    x, px = _f(x)
    Kp, pk = _f(Kp)
    W, pw = _f(W)
    dout, pg = _f(dout)
    n, M = nb.shape
    N0, Cin = x.shape
    K, _, Cout = W.shape
    dx = np.zeros((N0, Cin), dtype=np.float32)
    dW = np.zeros((K, Cin, Cout), dtype=np.float32)
    lib().orc_kpconv_bwd(pq, ps, pn, px, pk, pw, pg, ctypes.c_int64(n), ctypes.c_int64(N0), M, K, Cin, Cout,
                         ctypes.c_float(extent), INFLUENCE[influence], AGGREGATION[aggregation],
                         dx.ctypes.data_as(_F), dW.ctypes.data_as(_F))
    return dx, dW


def neighbor_max(x, nb):
    x, px = _f(x)
    nb, pn = _l(nb)
    N0, C = x.shape
    n, M = nb.shape
    out = np.zeros((n, C), dtype=np.float32)
    lib().orc_neighbor_max(px, pn, ctypes.c_int64(n), ctypes.c_int64(N0), M, C, out.ctypes.data_as(_F), None)
    return out


# ------------------------------------------------------- GridSampling3D ----
# numpy restatement of GridSampling3D._process (core/data_transform/grid_transform.py:110-126) and
# group_data (:33-81).  The arithmetic of the three third-party calls is restated from their published
# sources (not under /root/reference; "parity unpinned" except for the reference tests named below):
#   torch_cluster.grid_cluster 1.5.9 (cpu/grid_cpu.cpp): pos -= start; num_voxels = ((end - start) / size).long()
#       + 1; strides = exclusive cumprod (x fastest); cluster = sum((pos / size).long() * strides)
#   torch_geometric 1.7.2 voxel_grid: appends batch (cast to pos.dtype) as a 4th coordinate with size 1
#   torch_geometric 1.7.2 consecutive_cluster: unique(sorted, return_inverse); perm = empty.scatter_(0, inv, arange)
#       (sequential on the CPU: the LAST point of a voxel wins)
#   torch_scatter 2.0.8 scatter_mean: sequential fp32 scatter_add, then true_divide by the count
# Pinned by the reference's own tests: majority label (test/test_grid_sampling.py:30-36), idempotence of a
# second pass (:38-58), int32 quantised coordinates (:60-67), MultiScaleTransform consistency
# (test/test_transform.py:67-93).
def grid_cluster(pos, size, batch=None):
    """returns (cluster [N] int64 consecutive ids, unique_pos_indices [n] int64, coords [N,3] float32)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    coords = np.rint(pos / np.float32(size)).astype(np.float32)  # torch.round: half to even
    if batch is None:
        full = coords
    else:
        full = np.concatenate([coords, np.asarray(batch).astype(np.float32)[:, None]], axis=1)
    if full.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), coords
    start, end = full.min(0), full.max(0)
    num_voxels = (end - start).astype(np.float32).astype(np.int64) + 1
    strides = np.concatenate([[1], np.cumprod(num_voxels)[:-1]]).astype(np.int64)
    key = ((full - start[None, :]).astype(np.float32).astype(np.int64) * strides[None, :]).sum(1)
    _, inv = np.unique(key, return_inverse=True)
    perm = np.empty(int(inv.max()) + 1, dtype=np.int64)
    perm[inv] = np.arange(inv.shape[0], dtype=np.int64)  # numpy assigns in order: last write wins
    return inv.astype(np.int64), perm, coords


def scatter_mean(x, cluster, n):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((n,) + x.shape[1:], dtype=np.float32)
    np.add.at(out, cluster, x)  # unbuffered, in index order, fp32
    cnt = np.bincount(cluster, minlength=n).astype(np.float32)
    cnt = np.maximum(cnt, 1)
    return out / cnt.reshape((n,) + (1,) * (x.ndim - 1))


def label_vote(labels, cluster, n):
    labels = np.asarray(labels, dtype=np.int64)
    lmin = labels.min()
    hist = np.zeros((n, int(labels.max() - lmin) + 1), dtype=np.int64)
    np.add.at(hist, (cluster, labels - lmin), 1)
    return hist.argmax(1) + lmin


# --------------------------------------------------- deformable KPConv ----
# numpy restatement of KPConv_deform_ops (torch_points3d/modules/KPConv/convolution_ops.py:110-235) and of its
# autograd gradients.  PINNED by tests/golden/kpdeform_*.npz, generated by running the reference function and
# torch autograd in this container (tests/golden/make_golden.py).  The topk re-compaction of :158-185 is restated
# as what it computes: neighbours in range of no kernel point contribute nothing.
def _deform_tables(q, s, nb, Kp, offsets, extent, influence):
    q, s, Kp, offsets = (np.asarray(t, np.float32) for t in (q, s, Kp, offsets))
    N0 = s.shape[0]
    s1 = np.concatenate([s, np.full((1, 3), 1e6, np.float32)])
    nbc = np.where((nb < 0) | (nb >= N0), N0, nb)
    rel = s1[nbc] - q[:, None, :]
    kpd = offsets + Kp[None]
    diff = rel[:, :, None, :] - kpd[:, None, :, :]
    sqd = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    thr = np.float32(float(extent) ** 2)
    in_range = (sqd < thr).any(2)
    ext = np.float32(extent)
    if influence == "constant":
        w = (sqd < thr).astype(np.float32)
        slope = np.zeros_like(w)
    elif influence == "linear":
        r = np.sqrt(sqd)
        w = np.maximum(np.float32(1) - r / ext, np.float32(0))
        with np.errstate(divide="ignore"):
            slope = np.where(w > 0, np.float32(-1) / (np.float32(2) * ext * r), np.float32(0)).astype(np.float32)
    elif influence == "gaussian":
        den = np.float32(2.0 * (float(extent) * 0.3) ** 2 + 1e-9)
        w = np.exp(-sqd / den).astype(np.float32)
        slope = -w / den
    else:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    w = w * in_range[:, :, None]
    slope = slope * in_range[:, :, None]
    return nbc, diff, sqd, w.astype(np.float32), slope.astype(np.float32), kpd


def kpconv_deform_fwd(q, s, nb, x, Kp, offsets, mods, W, extent, influence="linear"):
    """returns (out [n,Cout], sq_distances [n,M,K], deformed_K_points [n,K,3])"""
    x, W = np.asarray(x, np.float32), np.asarray(W, np.float32)
    nbc, _, sqd, w, _, kpd = _deform_tables(q, s, nb, Kp, offsets, extent, influence)
    x1 = np.concatenate([x, np.zeros((1, x.shape[1]), np.float32)])
    wf = np.einsum("nmk,nmc->nkc", w, x1[nbc])
    if mods is not None:
        wf = wf * np.asarray(mods, np.float32)[:, :, None]
    return np.einsum("nkc,kco->no", wf, W).astype(np.float32), sqd, kpd


def kpconv_deform_bwd(q, s, nb, x, Kp, offsets, mods, W, extent, influence, dout, dsq=None, ddef=None):
    """gradients of sum(out*dout) + sum(sq_distances*dsq) + sum(deformed*ddef): (dx, dW, doffsets, dmods)"""
    x, W, dout = np.asarray(x, np.float32), np.asarray(W, np.float32), np.asarray(dout, np.float32)
    nbc, diff, sqd, w, slope, _ = _deform_tables(q, s, nb, Kp, offsets, extent, influence)
    N0 = x.shape[0]
    x1 = np.concatenate([x, np.zeros((1, x.shape[1]), np.float32)])
    fnb = x1[nbc]
    wf0 = np.einsum("nmk,nmc->nkc", w, fnb)
    m = np.ones(wf0.shape[:2], np.float32) if mods is None else np.asarray(mods, np.float32)
    dwfm = np.einsum("no,kco->nkc", dout, W)
    dW = np.einsum("nkc,no->kco", wf0 * m[:, :, None], dout)
    dmods = (dwfm * wf0).sum(2)
    dwf0 = dwfm * m[:, :, None]
    dx1 = np.zeros_like(x1)
    np.add.at(dx1, nbc, np.einsum("nmk,nkc->nmc", w, dwf0))
    gw = np.einsum("nkc,nmc->nmk", dwf0, fnb)
    dd2 = gw * slope
    if dsq is not None:
        dd2 = dd2 + np.asarray(dsq, np.float32)
    doff = np.einsum("nmk,nmkd->nkd", dd2, -2.0 * diff)
    if ddef is not None:
        doff = doff + np.asarray(ddef, np.float32)
    return dx1[:N0], dW, doff.astype(np.float32), (None if mods is None else dmods)


# ------------------------------------------------ MESSAGE_PASSING operators ----
# numpy restatement of what torch_geometric 1.7.2 executes around the per-edge MLP of PointConv (message: cat([x_j,
# pos_j - pos_i]); aggregate: scatter max, 0 for targets without edges) and RandlaKernel
# (modules/RandLANet/modules.py:34-40; aggregate: scatter add).  torch_geometric / torch_scatter are not under
# /root/reference and not installable here: "parity unpinned", the spec is the published operator definition.
def edge_message(x, pos_src, pos_dst, src, dst):
    rel = np.asarray(pos_src, np.float32)[src] - np.asarray(pos_dst, np.float32)[dst]
    return rel if x is None else np.concatenate([np.asarray(x, np.float32)[src], rel], axis=1)


def edge_encoding(pos_src, pos_dst, src, dst):
    pi, pj = np.asarray(pos_dst, np.float32)[dst], np.asarray(pos_src, np.float32)[src]
    v = pi - pj
    d = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])[:, None]
    return np.concatenate([pi, pj, v, d], axis=1).astype(np.float32)


def segment_max(msg, dst, T):
    msg = np.asarray(msg, np.float32)
    out = np.full((T, msg.shape[1]), -np.inf, np.float32)
    np.maximum.at(out, dst, msg)
    out[np.isinf(out)] = 0
    return out


def segment_sum(msg, dst, T):
    msg = np.asarray(msg, np.float32)
    out = np.zeros((T, msg.shape[1]), np.float32)
    np.add.at(out, dst, msg)  # in edge order
    return out
