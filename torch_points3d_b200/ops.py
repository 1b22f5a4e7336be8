"""Torch-facing operators over the C ABI (include/p3d_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every operator below is one
(or a few) launches of hand-written sm_100a kernels in libp3d_b200.so on
`torch.cuda.current_stream()`.  There is no CPU path and no PyTorch re-implementation to fall
back to: CPU tensors raise `RuntimeError`.

Signatures and error behaviour follow the reference's operator surface:
  torch-points-kernels 0.7.0  (furthest_point_sample, ball_query, grouping_operation, three_nn,
                               three_interpolate, knn)   -- see compat/tp.py
  torch-cluster 1.5.9 via torch_geometric.nn (fps, knn, radius, knn_interpolate) -- compat/cluster.py
  torch_points3d/modules/KPConv/convolution_ops.py:19-107 (KPConv_ops)
"""
import math
from typing import Optional, Tuple

import torch

from . import _cabi

INFLUENCE = {"constant": 0, "linear": 1, "gaussian": 2}
AGGREGATION = {"sum": 0, "closest": 1}
PRECISION = {"fp32": 0, "tensor": 1}
# KPConv forward arithmetic: "tensor" = kernel-point contraction on tcgen05 (bf16 hi+lo split operands,
# fp32 accumulate; <= 2e-5 of max|out| against the fp32 oracle) where the shape allows it,
# "fp32" = exact fp32 SIMT kernel everywhere (<= 1e-5).
KPCONV_PRECISION = "tensor"
# keep the fp32 weighted features of a tensor-path forward for the backward pass (n*K*Cin floats per layer) instead
# of recomputing them there
SAVE_KPCONV_WF = True


# ------------------------------------------------------------------ plumbing ---------------
def _require_cuda(name: str, *tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "%s: expected CUDA tensors (torch_points3d_b200 has no CPU fallback; the CPU restatement "
                "lives in oracle/ and is test infrastructure only)" % name)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("%s: tensors are on different devices (%s vs %s)" % (name, dev, t.device))
    if dev is None:
        raise RuntimeError("%s: no tensor given" % name)
    return dev


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous() if t.dtype == torch.float32 else t.to(torch.float32).contiguous()


def _i64(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous() if t.dtype == torch.int64 else t.to(torch.int64).contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _call(dev: torch.device, name: str, *args) -> None:
    with torch.cuda.device(dev):
        _cabi.call(name, *args, _stream())


def batch_to_ptr(batch: torch.Tensor, num_clouds: Optional[int] = None) -> torch.Tensor:
    """sorted batch vector -> CSR ptr [C+1] (device, int64)."""
    if num_clouds is None:
        num_clouds = int(batch[-1].item()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=num_clouds)
    ptr = torch.zeros(num_clouds + 1, dtype=torch.int64, device=batch.device)
    torch.cumsum(counts, 0, out=ptr[1:])
    return ptr


# ----------------------------------------------------------------------- FPS ---------------
def furthest_point_sample(xyz: torch.Tensor, npoint: int, return_xyz: bool = False):
    """tp.furthest_point_sample(xyz[B,N,3], npoint) -> idx[B,npoint]
    (reference call site core/spatial_ops/sampling.py:100).  Start index 0, lowest-index
    tie-break.  Returns int64 (tpk's CUDA path returned int32 and every caller cast it with
    `.long()`, core/base_conv/dense.py:75).  `return_xyz=True` also returns the sampled
    coordinates [B,npoint,3] written by the same kernel (the `pos.gather(1, idx)` of dense.py:75-76)."""
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise ValueError("furthest_point_sample expects xyz of shape [B, N, 3]")
    B, N, _ = xyz.shape
    if npoint > N:
        raise ValueError("caanot sample %i points from an input set of %i points" % (npoint, N))
    dev = _require_cuda("furthest_point_sample", xyz)
    xyz = _f32(xyz)
    idx = torch.empty((B, npoint), dtype=torch.int64, device=dev)
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=dev) if return_xyz else None
    ws_bytes = _cabi.lib().p3d_fps_dense_ws_bytes(B, N)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev) if ws_bytes else None
    _call(dev, "p3d_fps_dense_f32", xyz.data_ptr(), B, N, npoint, idx.data_ptr(), 1, _ptr(new_xyz), _ptr(ws),
          ws_bytes)
    return (idx, new_xyz) if return_xyz else idx


def fps(x: torch.Tensor, batch: Optional[torch.Tensor] = None, ratio: float = 0.5,
        random_start: bool = True) -> torch.Tensor:
    """torch_geometric.nn.fps == torch_cluster.fps (reference call site sampling.py:63):
    per cloud ceil(ratio * N_c) samples, GLOBAL int64 indices, concatenated in cloud order."""
    if x.dim() != 2 or x.shape[1] != 3:
        raise ValueError("fps expects x of shape [N, 3]")
    dev = _require_cuda("fps", x, batch)
    x = _f32(x)
    n_total = x.shape[0]
    if batch is None:
        ptr = torch.tensor([0, n_total], dtype=torch.int64, device=dev)
    else:
        if batch.numel() != n_total:
            raise ValueError("fps: batch must hold one entry per point")
        ptr = batch_to_ptr(_i64(batch))
    sizes = ptr[1:] - ptr[:-1]
    counts = torch.ceil(sizes.to(torch.float32) * ratio).to(torch.int64)  # torch_cluster 1.5.x: ceil
    counts = torch.minimum(counts, sizes)
    out_ptr = torch.zeros_like(ptr)
    torch.cumsum(counts, 0, out=out_ptr[1:])
    # the output length and the largest cloud are shapes: one small D2H, like torch_cluster's own
    host = torch.stack([out_ptr[-1], sizes.max() if sizes.numel() else out_ptr[-1]]).tolist()
    total_out, max_n = int(host[0]), int(host[1])
    idx = torch.empty(total_out, dtype=torch.int64, device=dev)
    start = None
    if random_start:
        start = (torch.rand(sizes.shape[0], device=dev) * sizes.to(torch.float32)).to(torch.int64)
        start = torch.minimum(start, torch.clamp(sizes - 1, min=0))
    ws_bytes = _cabi.lib().p3d_fps_csr_ws_bytes(n_total, max_n)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev) if ws_bytes else None
    _call(dev, "p3d_fps_csr_f32", x.data_ptr(), ptr.data_ptr(), ptr.shape[0] - 1, out_ptr.data_ptr(),
          idx.data_ptr(), _ptr(start), max_n, _ptr(ws), ws_bytes)
    return idx


# ---------------------------------------------------------------- ball query ---------------
def ball_query(radius: float, nsample: int, x: torch.Tensor, y: torch.Tensor, mode: str = "dense",
               batch_x: Optional[torch.Tensor] = None, batch_y: Optional[torch.Tensor] = None,
               sort: bool = False, return_dist: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """tp.ball_query (reference call sites neighbour_finder.py:35-37 and :164).

    dense:          x[B,N,3], y[B,S,3] -> idx[B,S,nsample] (padded with the first hit, 0 when the
                    ball is empty), dist2[B,S,nsample] (-1 padded); strict d2 < r2.
    partial_dense:  x[N,3], y[n,3], batch_x, batch_y -> idx[n,nsample] padded with the shadow
                    index x.shape[0], dist2 padded with -1; d2 <= r2.
    `return_dist=False` skips the dist2 store (every reference call site discards it)."""
    if mode is None:
        raise Exception('The mode should be defined within ["partial_dense | dense"]')
    mode = mode.lower()
    if sort:
        raise NotImplementedError("ball_query(sort=True) is only defined for the reference's CPU path")
    if mode == "partial_dense":
        if batch_x is None or batch_y is None:
            raise Exception("batch_x and batch_y should be provided")
        if x.dim() != 2 or y.dim() != 2:
            raise AssertionError("partial_dense ball_query expects x[N,3] and y[n,3]")
        dev = _require_cuda("ball_query", x, y, batch_x, batch_y)
        x, y, batch_x, batch_y = _f32(x), _f32(y), _i64(batch_x), _i64(batch_y)
        n = y.shape[0]
        idx = torch.empty((n, nsample), dtype=torch.int64, device=dev)
        dist2 = torch.empty((n, nsample), dtype=torch.float32, device=dev) if return_dist else None
        _ball_query_batch(dev, x, y, batch_x, batch_y, float(radius), int(nsample), 0, x.shape[0], idx, dist2)
        return idx, dist2
    if mode == "dense":
        if batch_x is not None or batch_y is not None:
            raise Exception("batch_x and batch_y should not be provided")
        if x.dim() != 3 or y.dim() != 3:
            raise AssertionError("dense ball_query expects x[B,N,3] and y[B,S,3]")
        dev = _require_cuda("ball_query", x, y)
        x, y = _f32(x), _f32(y)
        B, N, _ = x.shape
        S = y.shape[1]
        idx = torch.empty((B, S, nsample), dtype=torch.int64, device=dev)
        dist2 = torch.empty((B, S, nsample), dtype=torch.float32, device=dev) if return_dist else None
        _call(dev, "p3d_ball_query_dense_f32", x.data_ptr(), y.data_ptr(), B, N, S, float(radius), int(nsample), 1,
              idx.data_ptr(), _ptr(dist2))
        return idx, dist2
    raise Exception("unrecognized mode {}".format(mode))


# Batched searches switch to the cell-list kernels (csrc/cellgrid.cu: same results bit for bit) once the support set is
# large enough for the table build (a memset + 3 small launches, ~20 us) to pay: below this many support points the plain
# sweep is used.  Measured on B200 (profiles/r2_search_times.md): 50 k points in 8 clouds 87 vs 310 us, 15 k points in 8
# clouds 48 vs 39 us -- the sweep's cost grows with points PER CLOUD, which the host does not know without a device
# read, so the switch is on the total; lower it for batches of few large clouds.
GRID_SEARCH_MIN_POINTS = 24576
GRID_BALL_MAX_NSAMPLE = 64
GRID_KNN_MAX_K = 32


def _grid_ws(dev, nx):
    nbytes = _cabi.lib().p3d_cellgrid_ws_bytes(nx)
    return torch.empty(nbytes, dtype=torch.uint8, device=dev), nbytes  # caching-allocator blocks are 512-byte aligned


def _ball_query_batch(dev, x, y, batch_x, batch_y, radius, nsample, strict, shadow, idx, dist2):
    nx = x.shape[0]
    if nx >= GRID_SEARCH_MIN_POINTS and nsample <= GRID_BALL_MAX_NSAMPLE and radius > 0:
        ws, nbytes = _grid_ws(dev, nx)
        _call(dev, "p3d_ball_query_batch_grid_f32", x.data_ptr(), y.data_ptr(), _ptr(batch_x), _ptr(batch_y), nx,
              y.shape[0], radius, nsample, strict, shadow, idx.data_ptr(), _ptr(dist2), ws.data_ptr(), nbytes)
    else:
        _call(dev, "p3d_ball_query_batch_f32", x.data_ptr(), y.data_ptr(), _ptr(batch_x), _ptr(batch_y), nx,
              y.shape[0], radius, nsample, strict, shadow, idx.data_ptr(), _ptr(dist2))


def ball_query_multi(radii, nsamples, x: torch.Tensor, y: torch.Tensor, strict: bool = True):
    """every scale of a multi-scale-grouping layer in one sweep over the candidates: returns [idx_s [B,S,nsample_s]]
    with idx_s == ball_query(radii[s], nsamples[s], x, y)[0] bit for bit (DenseRadiusNeighbourFinder is called
    once per scale on the same pair, core/base_conv/dense.py:78-82)."""
    import ctypes
    dev = _require_cuda("ball_query_multi", x, y)
    if x.dim() != 3 or y.dim() != 3 or x.shape[0] != y.shape[0]:
        raise AssertionError("ball_query_multi expects x[B,N,3] and y[B,S,3]")
    if len(radii) != len(nsamples) or not 1 <= len(radii) <= 4:
        raise ValueError("ball_query_multi: 1..4 scales, one radius and one nsample per scale")
    x, y = _f32(x), _f32(y)
    B, N, _ = x.shape
    S = y.shape[1]
    ns = len(radii)
    out = [torch.empty((B, S, int(k)), dtype=torch.int64, device=dev) for k in nsamples]
    h_r = (ctypes.c_float * ns)(*[float(r) for r in radii])
    h_k = (ctypes.c_int * ns)(*[int(k) for k in nsamples])
    h_i = (ctypes.c_void_p * ns)(*[t.data_ptr() for t in out])
    _call(dev, "p3d_ball_query_dense_multi_f32", x.data_ptr(), y.data_ptr(), B, N, S, ns, ctypes.cast(h_r, ctypes.c_void_p),
          ctypes.cast(h_k, ctypes.c_void_p), 1 if strict else 0, ctypes.cast(h_i, ctypes.c_void_p))
    return out


def radius(x: torch.Tensor, y: torch.Tensor, r: float, batch_x: Optional[torch.Tensor] = None,
           batch_y: Optional[torch.Tensor] = None, max_num_neighbors: int = 32, num_workers: int = 1) -> torch.Tensor:
    """torch_geometric.nn.radius == torch_cluster.radius (reference call sites
    neighbour_finder.py:33,141): edge list [2,E], row = index into y, col = index into x,
    at most max_num_neighbors per query in candidate index order; strict d2 < r2."""
    dev = _require_cuda("radius", x, y, batch_x, batch_y)
    x, y = _f32(x.view(-1, 1) if x.dim() == 1 else x), _f32(y.view(-1, 1) if y.dim() == 1 else y)
    if x.shape[1] != 3:
        raise ValueError("radius: only 3-D points are supported")
    if (batch_x is None) != (batch_y is None):
        if batch_x is None:
            batch_x = torch.zeros(x.shape[0], dtype=torch.int64, device=dev)
        else:
            batch_y = torch.zeros(y.shape[0], dtype=torch.int64, device=dev)
    n = y.shape[0]
    shadow = -1
    idx = torch.empty((n, max_num_neighbors), dtype=torch.int64, device=dev)
    _ball_query_batch(dev, x, y, None if batch_x is None else _i64(batch_x), None if batch_y is None else _i64(batch_y),
                      float(r), int(max_num_neighbors), 1, shadow, idx, None)
    keep = idx >= 0
    row = torch.arange(n, device=dev).view(-1, 1).expand_as(idx)[keep]
    return torch.stack([row, idx[keep]], dim=0)


# ----------------------------------------------------------------------- knn ---------------
def knn_table(x: torch.Tensor, y: torch.Tensor, k: int, batch_x: Optional[torch.Tensor] = None,
              batch_y: Optional[torch.Tensor] = None, return_dist: bool = True):
    """k nearest supports (rows of x) of every query (row of y) inside its cloud:
    col[Ny,k] global indices ascending in d2 (ties -> lower index), -1 where the cloud has fewer
    than k points; d2[Ny,k] likewise."""
    dev = _require_cuda("knn", x, y, batch_x, batch_y)
    x, y = _f32(x), _f32(y)
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != 3 or y.shape[1] != 3:
        raise ValueError("knn expects x[N,3] and y[M,3]")
    if (batch_x is None) != (batch_y is None):
        if batch_x is None:
            batch_x = torch.zeros(x.shape[0], dtype=torch.int64, device=dev)
        else:
            batch_y = torch.zeros(y.shape[0], dtype=torch.int64, device=dev)
    bx = None if batch_x is None else _i64(batch_x)
    by = None if batch_y is None else _i64(batch_y)
    col = torch.empty((y.shape[0], k), dtype=torch.int64, device=dev)
    d2 = torch.empty((y.shape[0], k), dtype=torch.float32, device=dev) if return_dist else None
    if x.shape[0] >= GRID_SEARCH_MIN_POINTS and k <= GRID_KNN_MAX_K:
        ws, nbytes = _grid_ws(dev, x.shape[0])
        _call(dev, "p3d_knn_batch_grid_f32", x.data_ptr(), y.data_ptr(), _ptr(bx), _ptr(by), x.shape[0], y.shape[0],
              int(k), col.data_ptr(), _ptr(d2), ws.data_ptr(), nbytes)
    else:
        _call(dev, "p3d_knn_batch_f32", x.data_ptr(), y.data_ptr(), _ptr(bx), _ptr(by), x.shape[0], y.shape[0], int(k),
              col.data_ptr(), _ptr(d2))
    return col, d2


def knn_edges(x: torch.Tensor, y: torch.Tensor, k: int, batch_x: Optional[torch.Tensor] = None,
              batch_y: Optional[torch.Tensor] = None, cosine: bool = False, num_workers: int = 1) -> torch.Tensor:
    """torch_geometric.nn.knn == torch_cluster.knn (reference call sites neighbour_finder.py:47,
    interpolate.py:25): [2,E] int64, row = y index, col = x index, ascending distance per y."""
    if cosine:
        raise NotImplementedError("knn(cosine=True) is not on the reference's point-cloud path")
    col, _ = knn_table(x, y, k, batch_x, batch_y, return_dist=False)
    n = col.shape[0]
    row = torch.arange(n, device=col.device).view(-1, 1).expand_as(col)
    if batch_x is None and batch_y is None and x.shape[0] >= k:
        return torch.stack([row.reshape(-1), col.reshape(-1)], dim=0)  # every row is full: no sync
    keep = col >= 0
    return torch.stack([row[keep], col[keep]], dim=0)


def knn_dense(pos_support: torch.Tensor, pos: torch.Tensor, k: int):
    """tp.knn(pos_support[B,N,3], pos[B,M,3], k) -> (idx[B,M,k], dist2[B,M,k]) (tpk: CPU only)."""
    if pos_support.dim() != 3 or pos.dim() != 3:
        raise AssertionError("knn expects pos_support[B,N,3] and pos[B,M,3]")
    dev = _require_cuda("knn", pos_support, pos)
    xs, ys = _f32(pos_support), _f32(pos)
    B, N, _ = xs.shape
    M = ys.shape[1]
    if k > N:
        raise ValueError("knn: k (%d) larger than the number of support points (%d)" % (k, N))
    idx = torch.empty((B, M, k), dtype=torch.int64, device=dev)
    d2 = torch.empty((B, M, k), dtype=torch.float32, device=dev)
    _call(dev, "p3d_knn_dense_f32", xs.data_ptr(), ys.data_ptr(), B, N, M, int(k), idx.data_ptr(), d2.data_ptr())
    return idx, d2


def three_nn(unknown: torch.Tensor, known: torch.Tensor, return_weight: bool = False):
    """tp.three_nn(unknown[B,n,3], known[B,m,3]) -> (dist[B,n,3] = sqrt(d2), idx[B,n,3])
    (reference call site core/base_conv/dense.py:136).  `return_weight=True` appends the normalised
    inverse-distance weights of dense.py:137-139, computed by the same kernel."""
    if unknown.dim() != 3 or known.dim() != 3:
        raise AssertionError("three_nn expects unknown[B,n,3] and known[B,m,3]")
    if known.shape[1] < 3:
        raise ValueError("Not enough points. It should be at least 3 but got %i" % known.shape[1])
    dev = _require_cuda("three_nn", unknown, known)
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist = torch.empty((B, n, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((B, n, 3), dtype=torch.int64, device=dev)
    weight = torch.empty((B, n, 3), dtype=torch.float32, device=dev) if return_weight else None
    _call(dev, "p3d_three_nn_f32", unknown.data_ptr(), known.data_ptr(), B, n, m, dist.data_ptr(), idx.data_ptr(),
          _ptr(weight))
    return (dist, idx, weight) if return_weight else (dist, idx)


# ------------------------------------------------------------------ grouping ---------------
def _workspace(nbytes: int, dev: torch.device) -> Optional[torch.Tensor]:
    return torch.empty(nbytes, dtype=torch.uint8, device=dev) if nbytes else None


# The backward of grouping / three_interpolate needs the inverse of the index tensor, which depends on the
# indices only.  With PREPARE_INVERSE_IN_FORWARD the forward pass builds it on a side stream (it overlaps the
# bandwidth-bound gather and whatever the caller runs next) and the backward only waits on an event.
PREPARE_INVERSE_IN_FORWARD = True
_side_streams = {}


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=dev)
    return st


class _Prepared:
    """inverse index being built on a side stream: workspace + the event the backward waits on"""

    def __init__(self, name, dev, ws_bytes, idx, *args):
        self.ws_bytes = ws_bytes
        self.ws = _workspace(ws_bytes, dev)
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(cur)  # the indices (and the fresh workspace) are produced on the current stream
        with torch.cuda.stream(side):
            with torch.cuda.device(dev):
                _cabi.call(name, idx.data_ptr(), *args, self.ws.data_ptr(), ws_bytes, side.cuda_stream)
            self.event = torch.cuda.Event()
            self.event.record(side)
        if not torch.cuda.is_current_stream_capturing():
            # if no backward ever joins (a forward under grad mode whose graph is dropped), the allocator must still
            # wait for the side stream before it recycles the workspace OR the index tensor the kernel is reading
            self.ws.record_stream(side)
            idx.record_stream(side)

    def join(self, dev):
        torch.cuda.current_stream(dev).wait_event(self.event)


class PreparedInverse:
    """inverse index of a neighbour table, built ahead of time on the caller's stream (`prepare_grouping`,
    `prepare_three_interpolate`).  It depends on the indices only, so a pipeline can build it together with the
    neighbour search -- e.g. for the NEXT batch while the current one is in its feature stage -- and hand it to
    `group_features` / `grouping_operation` / `three_interpolate` through their `prepared=` argument."""

    def __init__(self, ws, ws_bytes):
        self.ws, self.ws_bytes = ws, ws_bytes

    def join(self, dev):  # same-stream (or caller-synchronised) producer: nothing to wait for
        pass


def prepare_grouping(idx: torch.Tensor, C: int, N: int) -> Optional[PreparedInverse]:
    """inverse index for the backward of grouping `C` channels of an `N`-point table through idx[B,S,K]
    (None when the shape uses the workspace-free path)"""
    dev = _require_cuda("prepare_grouping", idx)
    idx = _i64(idx)
    B, S, K = idx.shape
    ws_bytes = _cabi.lib().p3d_group_bwd_ws_bytes(B, C, N, S, K)
    if not ws_bytes:
        return None
    ws = _workspace(ws_bytes, dev)
    _call(dev, "p3d_group_bwd_prepare", idx.data_ptr(), B, C, N, S, K, ws.data_ptr(), ws_bytes)
    return PreparedInverse(ws, ws_bytes)


def prepare_three_interpolate(idx: torch.Tensor, C: int, m: int) -> Optional[PreparedInverse]:
    """inverse index for the backward of three_interpolate(features[B,C,m], idx[B,n,3], ...)"""
    dev = _require_cuda("prepare_three_interpolate", idx)
    idx = _i64(idx)
    B, n, _ = idx.shape
    ws_bytes = _cabi.lib().p3d_three_interp_bwd_ws_bytes(B, C, m, n)
    if not ws_bytes:
        return None
    ws = _workspace(ws_bytes, dev)
    _call(dev, "p3d_three_interp_bwd_prepare", idx.data_ptr(), B, C, m, n, ws.data_ptr(), ws_bytes)
    return PreparedInverse(ws, ws_bytes)


def _prepare_group(idx, B, C, N, S, K):
    if not PREPARE_INVERSE_IN_FORWARD:
        return None
    ws_bytes = _cabi.lib().p3d_group_bwd_ws_bytes(B, C, N, S, K)
    if not ws_bytes:
        return None
    return _Prepared("p3d_group_bwd_prepare", idx.device, ws_bytes, idx, B, C, N, S, K)


def _group_bwd(grad_out, idx, B, C, N, S, K, batch_stride, c_off, prepared=None):
    dev = grad_out.device
    dfeat = torch.empty((B, C, N), dtype=torch.float32, device=dev)
    if prepared is not None:
        prepared.join(dev)
        ws, ws_bytes, flag = prepared.ws, prepared.ws_bytes, 1
    else:
        ws_bytes = _cabi.lib().p3d_group_bwd_ws_bytes(B, C, N, S, K)
        ws, flag = _workspace(ws_bytes, dev), 0
    _call(dev, "p3d_group_bwd_f32", grad_out.data_ptr(), idx.data_ptr(), B, C, N, S, K, dfeat.data_ptr(),
          batch_stride, c_off, _ptr(ws), ws_bytes, flag)
    return dfeat


class _Grouping(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx, prepared=None):
        dev = _require_cuda("grouping_operation", features, idx)
        features, idx = _f32(features), _i64(idx)
        B, C, N = features.shape
        _, S, K = idx.shape
        out = torch.empty((B, C, S, K), dtype=torch.float32, device=dev)
        _call(dev, "p3d_group_fwd_f32", features.data_ptr(), idx.data_ptr(), B, C, N, S, K, out.data_ptr(),
              C * S * K, 0)
        ctx.save_for_backward(idx)
        ctx.N = N
        if prepared is None and ctx.needs_input_grad[0]:
            prepared = _prepare_group(idx, B, C, N, S, K)
        ctx.prepared = prepared
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = _f32(grad_out)
        B, C, S, K = grad_out.shape
        dfeat = _group_bwd(grad_out, idx, B, C, ctx.N, S, K, C * S * K, 0, ctx.prepared)
        return dfeat, None, None


def grouping_operation(features: torch.Tensor, idx: torch.Tensor, prepared=None) -> torch.Tensor:
    """tp.grouping_operation(features[B,C,N], idx[B,S,K]) -> [B,C,S,K]
    (reference call sites modules/pointnet2/dense.py:38,45; modules/RSConv/dense.py:147,...)."""
    if features.dim() != 3 or idx.dim() != 3 or features.shape[0] != idx.shape[0]:
        raise AssertionError("grouping_operation expects features[B,C,N] and idx[B,S,K]")
    return _Grouping.apply(features, idx, prepared)


class _GroupXyzFeat(torch.autograd.Function):
    """PointNetMSGDown._prepare_features (modules/pointnet2/dense.py:36-54) in two launches that
    write straight into the concatenated [B, 3+C, S, K] tensor: no transpose of pos, no
    grouped_pos temporary, no torch.cat copy."""

    @staticmethod
    def forward(ctx, x, pos, new_pos, idx, norm, use_xyz, prepared=None):
        dev = _require_cuda("group_features", pos, new_pos, idx, x)
        pos, new_pos, idx = _f32(pos), _f32(new_pos), _i64(idx)
        B, N, _ = pos.shape
        _, S, K = idx.shape
        C = 0 if x is None else x.shape[1]
        c_xyz = 3 if use_xyz else 0
        out = torch.empty((B, c_xyz + C, S, K), dtype=torch.float32, device=dev)
        stride = (c_xyz + C) * S * K
        if use_xyz and x is not None:  # both halves in one launch
            x = _f32(x)
            _call(dev, "p3d_group_xyz_feat_f32", pos.data_ptr(), new_pos.data_ptr(), x.data_ptr(), idx.data_ptr(), B,
                  C, N, S, K, float(norm), out.data_ptr())
        elif use_xyz:
            _call(dev, "p3d_group_xyz_f32", pos.data_ptr(), new_pos.data_ptr(), idx.data_ptr(), B, N, S, K,
                  float(norm), out.data_ptr(), stride)
        else:
            x = _f32(x)
            _call(dev, "p3d_group_fwd_f32", x.data_ptr(), idx.data_ptr(), B, C, N, S, K, out.data_ptr(), stride,
                  c_xyz)
        ctx.save_for_backward(idx)
        ctx.dims = (B, C, N, S, K, c_xyz)
        if prepared is None and C > 0 and ctx.needs_input_grad[0]:
            prepared = _prepare_group(idx, B, C, N, S, K)
        ctx.prepared = prepared
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, N, S, K, c_xyz = ctx.dims
        dx = None
        if C > 0 and ctx.needs_input_grad[0]:
            grad_out = _f32(grad_out)
            dx = _group_bwd(grad_out, idx, B, C, N, S, K, (c_xyz + C) * S * K, c_xyz, ctx.prepared)
        return dx, None, None, None, None, None, None


def group_features(x, pos, new_pos, idx, normalize_by: Optional[float] = None, use_xyz: bool = True, prepared=None):
    """fused `_prepare_features`: cat([(pos[idx] - new_pos) / r, x[idx]], dim=1) -> [B,3+C,S,K].
    Positions carry no gradient (as in the reference, where pos never requires grad)."""
    if x is None and not use_xyz:
        raise AssertionError("Cannot have not features and not use xyz as a feature!")
    norm = 1.0 if normalize_by is None else float(normalize_by)
    return _GroupXyzFeat.apply(x, pos, new_pos, idx, norm, use_xyz, prepared)


# --------------------------------------------------------- three_interpolate ---------------
class _ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx, weight, prepared=None):
        dev = _require_cuda("three_interpolate", features, idx, weight)
        features, idx, weight = _f32(features), _i64(idx), _f32(weight)
        B, C, m = features.shape
        n = idx.shape[1]
        out = torch.empty((B, C, n), dtype=torch.float32, device=dev)
        _call(dev, "p3d_three_interp_fwd_f32", features.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, C, m, n,
              out.data_ptr())
        ctx.save_for_backward(idx, weight)
        ctx.m = m
        ctx.prepared = prepared
        if prepared is None and PREPARE_INVERSE_IN_FORWARD and ctx.needs_input_grad[0]:
            ws_bytes = _cabi.lib().p3d_three_interp_bwd_ws_bytes(B, C, m, n)
            if ws_bytes:
                ctx.prepared = _Prepared("p3d_three_interp_bwd_prepare", dev, ws_bytes, idx, B, C, m, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        grad_out = _f32(grad_out)
        B, C, n = grad_out.shape
        dfeat = torch.empty((B, C, ctx.m), dtype=torch.float32, device=grad_out.device)
        if ctx.prepared is not None:
            ctx.prepared.join(grad_out.device)
            ws, ws_bytes, flag = ctx.prepared.ws, ctx.prepared.ws_bytes, 1
        else:
            ws_bytes = _cabi.lib().p3d_three_interp_bwd_ws_bytes(B, C, ctx.m, n)
            ws, flag = _workspace(ws_bytes, grad_out.device), 0
        _call(grad_out.device, "p3d_three_interp_bwd_f32", grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), B,
              C, ctx.m, n, dfeat.data_ptr(), _ptr(ws), ws_bytes, flag)
        return dfeat, None, None, None


def three_interpolate(features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, prepared=None) -> torch.Tensor:
    """tp.three_interpolate(features[B,C,m], idx[B,n,3], weight[B,n,3]) -> [B,C,n]
    (reference call site core/base_conv/dense.py:140); differentiable w.r.t. features."""
    if features.dim() != 3 or idx.dim() != 3 or weight.shape != idx.shape or idx.shape[2] != 3:
        raise AssertionError("three_interpolate expects features[B,C,m], idx[B,n,3], weight[B,n,3]")
    return _ThreeInterpolate.apply(features, idx, weight, prepared)


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k: int = 3, num_workers: int = 1):
    """torch_geometric.nn.knn_interpolate (reference call sites message_passing.py:173,
    interpolate.py:70; formula restated by the reference at interpolate.py:25-30,55-57):
    y = sum_k w x[col] / sum_k w with w = 1 / clamp(d2, 1e-16) over the k nearest pos_x."""
    col, d2 = knn_table(pos_x, pos_y, k, batch_x, batch_y, return_dist=True)
    valid = col >= 0
    w = torch.where(valid, 1.0 / torch.clamp(d2, min=1e-16), torch.zeros_like(d2))
    feats = x[col.clamp(min=0)]  # [Ny, k, C]
    num = (feats * w.unsqueeze(-1)).sum(1)
    return num / w.sum(1, keepdim=True)


# -------------------------------------------------------------------- KPConv ---------------
# ------------------------------------------------------------ unary layers of the KPConv blocks ------------
def _rowgemm(A: torch.Tensor, B: torch.Tensor, sbk: int, sbj: int, Kd: int, N: int) -> torch.Tensor:
    """C[R, N] = A[R, Kd] . B with B(k, j) = B.flatten()[k * sbk + j * sbj] on the tcgen05 row GEMM (csrc/rowgemm_tc.cu)"""
    dev = A.device
    R = A.shape[0]
    C = torch.empty((R, N), dtype=torch.float32, device=dev)
    ws_bytes = _cabi.lib().p3d_rowgemm_ws_bytes(Kd, N)
    ws = _workspace(ws_bytes, dev)
    _call(dev, "p3d_rowgemm_f32", A.data_ptr(), A.stride(0), B.data_ptr(), sbk, sbj, R, Kd, N, C.data_ptr(), N, _ptr(ws),
          ws_bytes)
    return C


def rowgemm_supported(Kd: int, N: int) -> bool:
    return N % 16 == 0 and N >= 16 and Kd >= 1


# Tables with fewer rows than this keep the library GEMMs in ops.linear: with 128-row tiles a level of a few thousand
# points gives each SM one or two tiles and the kernel's fill / drain is the whole run (bench_tools/linear_times.py:
# 2580 x 256 -> 64 forward 16.6 us against 8.5 us; 50 614 x 64 -> 128 forward + backward 82 us against 183 us)
LINEAR_MIN_ROWS = 8192


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        _require_cuda("linear", x, weight)
        x, weight = _f32(x), _f32(weight)
        Cout, Cin = weight.shape
        if x.shape[1] != Cin:
            raise ValueError("linear: x is [n, %d] but weight is %s" % (x.shape[1], tuple(weight.shape)))
        ctx.save_for_backward(x, weight)
        if x.shape[0] < LINEAR_MIN_ROWS or not rowgemm_supported(Cin, Cout):
            return x @ weight.t()
        return _rowgemm(x, weight, 1, Cin, Cin, Cout)          # B(k, j) = W[j][k]

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _f32(dy)
        Cout, Cin = weight.shape
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if x.shape[0] < LINEAR_MIN_ROWS or not rowgemm_supported(Cout, Cin):
                dx = dy @ weight
            else:
                dx = _rowgemm(dy, weight, Cin, 1, Cout, Cin)   # B(k, j) = W[k][j]
        if ctx.needs_input_grad[1]:
            n = x.shape[0]
            if n >= LINEAR_MIN_ROWS and Cin % 16 == 0:
                # dW[Cout, Cin] = dy^T . x: reduction over the points cut over CTAs, deterministic partial sums
                dw = torch.zeros((Cout, Cin), dtype=torch.float32, device=dy.device)
                ws_bytes = _cabi.lib().p3d_rowgemm_dw_ws_bytes(n, Cout, Cin)
                ws = _workspace(ws_bytes, dy.device)
                _call(dy.device, "p3d_rowgemm_dw_f32", dy.data_ptr(), x.data_ptr(), n, Cout, Cin, dw.data_ptr(), _ptr(ws),
                      ws_bytes)
            else:
                dw = dy.t() @ x
        return dx, dw


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x @ weight.T (+ bias) on a [n, Cin] point-feature table: the unary layers of the KPConv blocks
    (modules/KPConv/blocks.py:176-194) and the decoder MLPs (base_modules.py:35-45).  For tables of at least
    LINEAR_MIN_ROWS rows the forward, the input gradient and the weight gradient (a reduction over all points with a
    tiny output: the library runs it as a split-K launch of a few dozen CTAs) run on the tcgen05 row GEMM
    (csrc/rowgemm_tc.cu; widths that are multiples of 16); smaller tables and other widths use the library."""
    y = _Linear.apply(x, weight)
    return y if bias is None else y + bias


class Linear(torch.nn.Linear):
    """torch.nn.Linear with the same parameters and state_dict keys; [n, Cin] fp32 CUDA tables go through `linear`,
    anything else through the stock implementation"""

    def forward(self, x):
        if x.dim() == 2 and x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32:
            return linear(x, self.weight, self.bias)
        return super().forward(x)


class _KPConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, support, neighbors, feat, k_points, k_values, extent, influence, aggregation,
                precision, save_wf=False):
        dev = _require_cuda("KPConv_ops", query, support, neighbors, feat, k_points, k_values)
        query, support, feat = _f32(query), _f32(support), _f32(feat)
        k_points, k_values, neighbors = _f32(k_points), _f32(k_values), _i64(neighbors)
        n, M = neighbors.shape
        N0, Cin = feat.shape
        K, _, Cout = k_values.shape
        out = torch.empty((n, Cout), dtype=torch.float32, device=dev)
        ws_bytes = _cabi.lib().p3d_kpconv_fwd_ws_bytes(n, K, Cin, Cout) if precision else 0
        ws = _workspace(ws_bytes, dev)
        wf = None
        if save_wf and ws_bytes and M <= 40 and n > 0 and ctx.needs_input_grad[5]:
            # training on the tensor path: keep the weighted features the kernel builds anyway (n*K*Cin floats)
            wf = torch.empty((n, K, Cin), dtype=torch.float32, device=dev)
            _call(dev, "p3d_kpconv_fwd_wf_f32", query.data_ptr(), support.data_ptr(), neighbors.data_ptr(),
                  feat.data_ptr(), k_points.data_ptr(), k_values.data_ptr(), n, N0, M, K, Cin, Cout, float(extent),
                  influence, aggregation, out.data_ptr(), wf.data_ptr(), ws.data_ptr(), ws_bytes)
        else:
            _call(dev, "p3d_kpconv_fwd_f32", query.data_ptr(), support.data_ptr(), neighbors.data_ptr(),
                  feat.data_ptr(), k_points.data_ptr(), k_values.data_ptr(), n, N0, M, K, Cin, Cout, float(extent),
                  influence, aggregation, precision, out.data_ptr(), _ptr(ws), ws_bytes)
        ctx.wf = wf
        ctx.save_for_backward(query, support, neighbors, feat, k_points, k_values)
        ctx.cfg = (float(extent), influence, aggregation)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        query, support, neighbors, feat, k_points, k_values = ctx.saved_tensors
        extent, influence, aggregation = ctx.cfg
        dev = grad_out.device
        grad_out = _f32(grad_out)
        n, M = neighbors.shape
        N0, Cin = feat.shape
        K, _, Cout = k_values.shape
        need_x, need_w = ctx.needs_input_grad[3], ctx.needs_input_grad[5]
        dfeat = torch.zeros_like(feat) if need_x else None
        dW = torch.zeros_like(k_values) if need_w else None
        if need_x or need_w:
            ws_bytes = _cabi.lib().p3d_kpconv_bwd_ws_bytes(n, M, K, Cin, Cout)
            ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.float32, device=dev)
            _call(dev, "p3d_kpconv_bwd_wf_f32", query.data_ptr(), support.data_ptr(), neighbors.data_ptr(),
                  feat.data_ptr(), k_points.data_ptr(), k_values.data_ptr(), grad_out.data_ptr(), _ptr(ctx.wf), n, N0,
                  M, K, Cin, Cout, extent, influence, aggregation, _ptr(dfeat), _ptr(dW), ws.data_ptr(), ws_bytes)
            ctx.wf = None
        return None, None, None, dfeat, None, dW, None, None, None, None, None


class _KPConvSplit(torch.autograd.Function):
    """KPConv_ops as (influence-weighted neighbour aggregation kernel) + (plain library GEMM): used for the deep
    encoder levels, a few hundred query points against 15 x 512 x 512 weights, where the contraction is a
    weight-streaming GEMM (exact fp32)."""

    @staticmethod
    def forward(ctx, query, support, neighbors, feat, k_points, k_values, extent, influence, aggregation):
        dev = _require_cuda("KPConv_ops", query, support, neighbors, feat, k_points, k_values)
        query, support, feat = _f32(query), _f32(support), _f32(feat)
        k_points, k_values, neighbors = _f32(k_points), _f32(k_values), _i64(neighbors)
        n, M = neighbors.shape
        N0, Cin = feat.shape
        K, _, Cout = k_values.shape
        wf = torch.empty((n, K * Cin), dtype=torch.float32, device=dev)
        _call(dev, "p3d_kpconv_wf_f32", query.data_ptr(), support.data_ptr(), neighbors.data_ptr(), feat.data_ptr(),
              k_points.data_ptr(), n, N0, M, K, Cin, float(extent), influence, aggregation, wf.data_ptr())
        ctx.save_for_backward(query, support, neighbors, k_points, k_values, wf)
        ctx.cfg = (float(extent), influence, aggregation, N0, Cin)
        return wf @ k_values.reshape(K * Cin, Cout)

    @staticmethod
    def backward(ctx, grad_out):
        query, support, neighbors, k_points, k_values, wf = ctx.saved_tensors
        extent, influence, aggregation, N0, Cin = ctx.cfg
        grad_out = _f32(grad_out)
        n, M = neighbors.shape
        K, _, Cout = k_values.shape
        dW = dfeat = None
        if ctx.needs_input_grad[5]:
            dW = (wf.t() @ grad_out).reshape(K, Cin, Cout)
        if ctx.needs_input_grad[3]:
            dwf = (grad_out @ k_values.reshape(K * Cin, Cout).t()).contiguous()
            dfeat = torch.zeros((N0, Cin), dtype=torch.float32, device=grad_out.device)
            _call(grad_out.device, "p3d_kpconv_scatter_f32", query.data_ptr(), support.data_ptr(), neighbors.data_ptr(),
                  k_points.data_ptr(), dwf.data_ptr(), n, N0, M, K, Cin, extent, influence, aggregation,
                  dfeat.data_ptr())
        return None, None, None, dfeat, None, dW, None, None, None


# Opt-in (default off): query-point count below which KPConv_ops runs as aggregation kernel + library GEMM (see
# _KPConvSplit).  The fused tcgen05 kernel spreads small point sets over the SMs by output channels and by slabs of the
# contraction and is the faster forward at every level of the KPConvPaper encoder (profiles/r2_kpconv_layers.md).
KPCONV_SPLIT_MAX_POINTS = 0
KPCONV_SPLIT_MIN_WEIGHTS = 15 * 64 * 64


def KPConv_ops(query_points, support_points, neighbors_indices, features, K_points, K_values, KP_extent,
               KP_influence, aggregation_mode, precision=None):
    """KPConv_ops (torch_points3d/modules/KPConv/convolution_ops.py:19-107), same positional
    signature.  neighbors_indices[n,M] int64; entries == support_points.shape[0] (or -1, as
    gathering.py:10 allows) are shadow neighbours.  Differentiable w.r.t. features and K_values
    (K_points is a non-trainable Parameter in the reference, kernels.py:66-68)."""
    if KP_influence not in INFLUENCE:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    if aggregation_mode not in AGGREGATION:
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    if K_values.dim() != 3 or K_values.shape[0] != K_points.shape[0] or K_values.shape[1] != features.shape[1]:
        raise AssertionError("K_values must be [n_kernel_points, Cin, Cout]")
    if (precision is None and neighbors_indices.shape[0] <= KPCONV_SPLIT_MAX_POINTS
            and K_values.numel() >= KPCONV_SPLIT_MIN_WEIGHTS):
        return _KPConvSplit.apply(query_points, support_points, neighbors_indices, features, K_points, K_values,
                                  float(KP_extent), INFLUENCE[KP_influence], AGGREGATION[aggregation_mode])
    prec = precision or KPCONV_PRECISION
    cin, cout = features.shape[1], K_values.shape[2]
    if (prec == "tensor" and cin % 32 != 0 and cin < 32 and cout % 32 == 0 and cout <= 512
            and neighbors_indices.shape[1] <= 40 and features.is_cuda):
        # first layer of a network (a handful of raw channels): zero-pad Cin to 32 so the layer runs on the tcgen05
        # path and the sparse backward kernels instead of the generic fp32 ones; the padding contributes exact zeros
        pad = 32 - cin
        features = torch.nn.functional.pad(features, (0, pad))
        K_values = torch.nn.functional.pad(K_values, (0, 0, 0, pad))
    # (grad mode is always off inside Function.forward, so "is this a training forward" is decided here)
    save_wf = SAVE_KPCONV_WF and prec == "tensor" and torch.is_grad_enabled() and K_values.requires_grad
    return _KPConv.apply(query_points, support_points, neighbors_indices, features, K_points, K_values,
                         float(KP_extent), INFLUENCE[KP_influence], AGGREGATION[aggregation_mode], PRECISION[prec],
                         save_wf)



# --------------------------------------------------------- deformable KPConv ---------------
class _KPConvDeform(torch.autograd.Function):
    """outputs (out [n,Cout], sq_distances [n,M,K] or empty, kpmin [n,K] or empty)"""

    @staticmethod
    def forward(ctx, query, support, neighbors, feat, k_points, offsets, mods, k_values, extent, influence, want_sqd,
                want_kpmin):
        dev = _require_cuda("KPConv_deform_ops", query, support, neighbors, feat, k_points, offsets, mods, k_values)
        query, support, feat, offsets = _f32(query), _f32(support), _f32(feat), _f32(offsets)
        k_points, k_values, neighbors = _f32(k_points), _f32(k_values), _i64(neighbors)
        mods = None if mods is None else _f32(mods)
        n, M = neighbors.shape
        N0, Cin = feat.shape
        K, _, Cout = k_values.shape
        out = torch.empty((n, Cout), dtype=torch.float32, device=dev)
        wf = torch.empty((n, K, Cin), dtype=torch.float32, device=dev)
        sqd = torch.empty((n, M, K) if want_sqd else (0,), dtype=torch.float32, device=dev)
        kpmin = torch.empty((n, K) if want_kpmin else (0,), dtype=torch.float32, device=dev)
        argmin = torch.empty((n, K), dtype=torch.int32, device=dev) if want_kpmin else None
        _call(dev, "p3d_kpconv_deform_fwd_f32", query.data_ptr(), support.data_ptr(), neighbors.data_ptr(),
              feat.data_ptr(), k_points.data_ptr(), offsets.data_ptr(), _ptr(mods), k_values.data_ptr(), n, N0, M, K,
              Cin, Cout, float(extent), influence, out.data_ptr(), wf.data_ptr(), sqd.data_ptr() if want_sqd else None,
              kpmin.data_ptr() if want_kpmin else None, _ptr(argmin))
        ctx.save_for_backward(query, support, neighbors, feat, k_points, offsets, mods, k_values, wf, argmin)
        ctx.cfg = (float(extent), influence)
        ctx.set_materialize_grads(False)
        return out, sqd, kpmin

    @staticmethod
    def backward(ctx, g_out, g_sqd, g_kpmin):
        query, support, neighbors, feat, k_points, offsets, mods, k_values, wf, argmin = ctx.saved_tensors
        extent, influence = ctx.cfg
        dev = query.device
        n, M = neighbors.shape
        N0, Cin = feat.shape
        K, _, Cout = k_values.shape
        need = ctx.needs_input_grad
        g_out = torch.zeros((n, Cout), dtype=torch.float32, device=dev) if g_out is None else _f32(g_out)
        g_sqd = None if g_sqd is None or g_sqd.numel() == 0 else _f32(g_sqd)
        g_kpmin = None if g_kpmin is None or g_kpmin.numel() == 0 else _f32(g_kpmin)
        dfeat = torch.zeros_like(feat) if need[3] else None
        dW = torch.zeros_like(k_values) if need[7] else None
        doff = torch.empty_like(offsets) if need[5] else None
        dmod = torch.empty_like(mods) if (mods is not None and need[6]) else None
        ws_bytes = _cabi.lib().p3d_kpconv_deform_bwd_ws_bytes(n, K, Cin)
        ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.float32, device=dev)
        _call(dev, "p3d_kpconv_deform_bwd_f32", query.data_ptr(), support.data_ptr(), neighbors.data_ptr(),
              feat.data_ptr(), k_points.data_ptr(), offsets.data_ptr(), _ptr(mods), k_values.data_ptr(), wf.data_ptr(),
              g_out.data_ptr(), _ptr(g_sqd), _ptr(g_kpmin), _ptr(argmin), n, N0, M, K, Cin, Cout, extent, influence,
              _ptr(dfeat), _ptr(dW), _ptr(doff), _ptr(dmod), ws.data_ptr(), ws_bytes)
        return None, None, None, dfeat, None, doff, dmod, dW, None, None, None, None


def _check_deform(features, K_points, offsets, modulations, K_values, KP_influence, aggregation_mode):
    if KP_influence not in INFLUENCE:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    if aggregation_mode == "closest":
        # convolution_ops.py:196 calls torch.argmin(..., output_type=torch.long): the reference cannot run this mode
        raise TypeError("argmin() received an invalid combination of arguments - got output_type (the reference's "
                        "deformable 'closest' aggregation does not run; only 'sum' is offered)")
    if aggregation_mode != "sum":
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    K = K_points.shape[0]
    if K_values.dim() != 3 or K_values.shape[0] != K or K_values.shape[1] != features.shape[1]:
        raise AssertionError("K_values must be [n_kernel_points, Cin, Cout]")
    if offsets.dim() != 3 or offsets.shape[1] != K or offsets.shape[2] != 3:
        raise AssertionError("offsets must be [n_points, n_kernel_points, 3]")
    if modulations is not None and tuple(modulations.shape) != tuple(offsets.shape[:2]):
        raise AssertionError("modulations must be [n_points, n_kernel_points]")


def KPConv_deform_ops(query_points, support_points, neighbors_indices, features, K_points, offsets, modulations,
                      K_values, KP_extent, KP_influence, aggregation_mode):
    """KPConv_deform_ops (torch_points3d/modules/KPConv/convolution_ops.py:110-235), same positional signature and
    return value: (output_features [n,Cout], sq_distances [n,M,K], deformed_K_points [n,K,3]).  Differentiable
    w.r.t. features, offsets, modulations and K_values; sq_distances and deformed_K_points carry gradients to
    offsets like the reference's (they feed the regularisers of modules/KPConv/losses.py)."""
    _check_deform(features, K_points, offsets, modulations, K_values, KP_influence, aggregation_mode)
    out, sqd, _ = _KPConvDeform.apply(query_points, support_points, neighbors_indices, features, K_points, offsets,
                                      modulations, K_values, float(KP_extent), INFLUENCE[KP_influence], True, False)
    return out, sqd, torch.add(offsets, K_points)


def KPConv_deform_fitting(query_points, support_points, neighbors_indices, features, K_points, offsets, modulations,
                          K_values, KP_extent, KP_influence, aggregation_mode):
    """Same operator for callers that only reduce sq_distances the way fitting_loss does (losses.py:4-16):
    returns (output_features, kpmin [n,K] = sq_distances.min(dim=1)[0], deformed_K_points) without ever writing the
    [n,M,K] tensor; kpmin is differentiable w.r.t. offsets through the first argmin, like torch.min."""
    _check_deform(features, K_points, offsets, modulations, K_values, KP_influence, aggregation_mode)
    out, _, kpmin = _KPConvDeform.apply(query_points, support_points, neighbors_indices, features, K_points, offsets,
                                        modulations, K_values, float(KP_extent), INFLUENCE[KP_influence], False, True)
    return out, kpmin, torch.add(offsets, K_points)


class _NeighborMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, neighbors):
        dev = _require_cuda("neighbor_max", x, neighbors)
        x, neighbors = _f32(x), _i64(neighbors)
        N0, C = x.shape
        n, M = neighbors.shape
        out = torch.empty((n, C), dtype=torch.float32, device=dev)
        arg = torch.empty((n, C), dtype=torch.int64, device=dev)
        _call(dev, "p3d_neighbor_max_f32", x.data_ptr(), neighbors.data_ptr(), n, N0, M, C, out.data_ptr(),
              arg.data_ptr())
        ctx.save_for_backward(arg)
        ctx.N0 = N0
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (arg,) = ctx.saved_tensors
        n, C = grad_out.shape
        grad_out = _f32(grad_out)
        dx = torch.empty((ctx.N0, C), dtype=torch.float32, device=grad_out.device)
        _call(grad_out.device, "p3d_neighbor_max_bwd_f32", grad_out.data_ptr(), arg.data_ptr(), n, ctx.N0, C,
              dx.data_ptr())
        return dx, None


def neighbor_max(x: torch.Tensor, neighbors: torch.Tensor) -> torch.Tensor:
    """strided-shortcut max-pool of ResnetBBlock (modules/KPConv/blocks.py:206-210):
    max over the neighbour rows of cat(x, zeros(1,C))."""
    return _NeighborMax.apply(x, neighbors)


# ------------------------------------------------------------ GridSampling3D ---------------
class VoxelClusters:
    """result of `grid_cluster`: consecutive voxel ids plus the segment list the mean / vote kernels consume"""

    def __init__(self, cluster, unique_pos_indices, order, seg_start, n, mean_pos, batch, coords):
        self.cluster, self.unique_pos_indices, self.order, self.seg_start = cluster, unique_pos_indices, order, seg_start
        self.n, self.mean_pos, self.batch, self.coords = n, mean_pos, batch, coords


def grid_cluster(pos: torch.Tensor, size: float, batch: Optional[torch.Tensor] = None, want_mean: bool = True,
                 want_coords: bool = False) -> VoxelClusters:
    """`consecutive_cluster(voxel_grid(torch.round(pos / size), batch, 1))` of GridSampling3D._process
    (core/data_transform/grid_transform.py:110-120) in one cooperative kernel, together with the per-voxel
    mean position (group_data mode="mean"), the voxel's batch id and, on request, the int32 quantised
    coordinates.  One D2H read of the voxel count sizes the outputs (the reference's torch.unique does the same)."""
    dev = _require_cuda("grid_cluster", pos, batch)
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise ValueError("grid_cluster expects pos of shape [N, 3]")
    pos = _f32(pos)
    N = pos.shape[0]
    b = None if batch is None else _i64(batch)
    cluster = torch.empty(N, dtype=torch.int64, device=dev)
    last = torch.empty(N, dtype=torch.int64, device=dev)
    order = torch.empty(N, dtype=torch.int32, device=dev)
    seg = torch.empty(N + 1, dtype=torch.int32, device=dev)
    mean = torch.empty(N, 3, dtype=torch.float32, device=dev) if want_mean else None
    ob = torch.empty(N, dtype=torch.int64, device=dev) if b is not None else None
    co = torch.empty(N, 3, dtype=torch.int32, device=dev) if want_coords else None
    n_dev = torch.empty(1, dtype=torch.int64, device=dev)
    ws_bytes = _cabi.lib().p3d_grid_cluster_ws_bytes(N)
    ws = _workspace(ws_bytes, dev)
    _call(dev, "p3d_grid_cluster_f32", pos.data_ptr(), _ptr(b), N, float(size), cluster.data_ptr(), last.data_ptr(),
          order.data_ptr(), seg.data_ptr(), _ptr(mean), _ptr(ob), _ptr(co), n_dev.data_ptr(), ws.data_ptr(), ws_bytes)
    n = int(n_dev.item())
    return VoxelClusters(cluster, last[:n], order, seg[:n + 1], n, None if mean is None else mean[:n],
                         None if ob is None else ob[:n].to(batch.dtype), None if co is None else co[:n])


def voxels_from_cluster(cluster: torch.Tensor, unique_pos_indices: Optional[torch.Tensor] = None) -> VoxelClusters:
    """the segment list of consecutive cluster ids that did not come from `grid_cluster` (group_data called with a
    caller's own `cluster`, grid_transform.py:33): points sorted by (cluster, point index), first slot per cluster"""
    cluster = _i64(cluster)
    n = int(cluster.max().item()) + 1 if cluster.numel() else 0
    order = torch.sort(cluster, stable=True).indices.to(torch.int32)
    seg = torch.zeros(n + 1, dtype=torch.int32, device=cluster.device)
    if n:
        seg[1:] = torch.bincount(cluster, minlength=n).cumsum(0).to(torch.int32)
    return VoxelClusters(cluster, unique_pos_indices, order, seg, n, None, None, None)


def segment_mean(x: torch.Tensor, vc: VoxelClusters) -> torch.Tensor:
    """`scatter_mean(x, cluster, dim=0)` (grid_transform.py:76) over the voxels of `vc`; sums run in ascending
    point index in fp32, which is what the sequential CPU scatter_add of the reference produces."""
    dev = _require_cuda("segment_mean", x, vc.order)
    if x.shape[0] != vc.cluster.shape[0]:
        raise ValueError("segment_mean: x has %d rows, the clustering %d" % (x.shape[0], vc.cluster.shape[0]))
    shape = x.shape[1:]
    x2 = _f32(x).reshape(x.shape[0], -1)
    C = x2.shape[1]
    out = torch.empty(vc.n, C, dtype=torch.float32, device=dev)
    _call(dev, "p3d_segment_mean_f32", x2.data_ptr(), vc.order.data_ptr(), vc.seg_start.data_ptr(), None, vc.n, C,
          out.data_ptr())
    return out.reshape((vc.n,) + tuple(shape))


def segment_vote(labels: torch.Tensor, vc: VoxelClusters) -> torch.Tensor:
    """majority label per voxel (grid_transform.py:70-74); ties go to the lowest label"""
    dev = _require_cuda("segment_vote", labels, vc.cluster)
    lab = _i64(labels)
    if lab.numel() == 0:
        return lab.clone()
    lmin, lmax = int(lab.min().item()), int(lab.max().item())
    L = lmax - lmin + 1
    hist = torch.zeros(vc.n, L, dtype=torch.int32, device=dev)
    out = torch.empty(vc.n, dtype=torch.int64, device=dev)
    _call(dev, "p3d_segment_vote_i64", lab.data_ptr(), vc.cluster.data_ptr(), lab.shape[0], vc.n, lmin, L,
          hist.data_ptr(), out.data_ptr())
    return out.to(labels.dtype)


# ------------------------------------------------- MESSAGE_PASSING operators ---------------
class _EdgeMessage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos_src, pos_dst, src, dst):
        dev = _require_cuda("edge_message", x, pos_src, pos_dst, src, dst)
        pos_src, pos_dst, src, dst = _f32(pos_src), _f32(pos_dst), _i64(src), _i64(dst)
        x = None if x is None else _f32(x)
        C = 0 if x is None else x.shape[1]
        E = src.shape[0]
        out = torch.empty((E, C + 3), dtype=torch.float32, device=dev)
        _call(dev, "p3d_edge_message_f32", _ptr(x), pos_src.data_ptr(), pos_dst.data_ptr(), src.data_ptr(),
              dst.data_ptr(), E, C, out.data_ptr())
        ctx.save_for_backward(src)
        ctx.shape = None if x is None else tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (src,) = ctx.saved_tensors
        if ctx.shape is None or not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        grad_out = _f32(grad_out)
        dx = torch.zeros(ctx.shape, dtype=torch.float32, device=grad_out.device)
        _call(grad_out.device, "p3d_edge_message_bwd_f32", grad_out.data_ptr(), src.data_ptr(), src.shape[0],
              ctx.shape[1], dx.data_ptr())
        return dx, None, None, None, None


def edge_message(x: Optional[torch.Tensor], pos_src: torch.Tensor, pos_dst: torch.Tensor, src: torch.Tensor,
                 dst: torch.Tensor) -> torch.Tensor:
    """PointConv.message before its local_nn (torch_geometric 1.7.2): cat([x_j, pos_j - pos_i], dim=1) for every
    edge, j = src (support index), i = dst (query index); x may be None.  Differentiable w.r.t. x."""
    return _EdgeMessage.apply(x, pos_src, pos_dst, src, dst)


def edge_encoding(pos_src: torch.Tensor, pos_dst: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """RandlaKernel.message relative point position encoding (modules/RandLANet/modules.py:34-40):
    cat([pos_i, pos_j, pos_i - pos_j, |pos_i - pos_j|]) -> [E, 10]"""
    dev = _require_cuda("edge_encoding", pos_src, pos_dst, src, dst)
    pos_src, pos_dst, src, dst = _f32(pos_src), _f32(pos_dst), _i64(src), _i64(dst)
    out = torch.empty((src.shape[0], 10), dtype=torch.float32, device=dev)
    _call(dev, "p3d_edge_encoding_f32", pos_src.data_ptr(), pos_dst.data_ptr(), src.data_ptr(), dst.data_ptr(),
          src.shape[0], out.data_ptr())
    return out


class _SegmentReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msg, dst, T, mode):
        dev = _require_cuda("segment_reduce", msg, dst)
        msg, dst = _f32(msg), _i64(dst)
        E, C = msg.shape
        out = torch.empty((T, C), dtype=torch.float32, device=dev)
        arg = torch.empty((T, C), dtype=torch.int64, device=dev) if mode == 0 else None
        _call(dev, "p3d_segment_reduce_f32", msg.data_ptr(), dst.data_ptr(), E, T, C, mode, out.data_ptr(), _ptr(arg))
        ctx.save_for_backward(dst, arg)
        ctx.cfg = (E, T, C, mode)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        dst, arg = ctx.saved_tensors
        E, T, C, mode = ctx.cfg
        grad_out = _f32(grad_out)
        dmsg = torch.empty((E, C), dtype=torch.float32, device=grad_out.device)
        _call(grad_out.device, "p3d_segment_reduce_bwd_f32", grad_out.data_ptr(), dst.data_ptr(), _ptr(arg), E, T, C,
              mode, dmsg.data_ptr())
        return dmsg, None, None, None


def segment_max(msg: torch.Tensor, dst: torch.Tensor, num_targets: int) -> torch.Tensor:
    """aggr="max" of MessagePassing over edges sorted by target: [E,C] -> [T,C]; a target without edges gets 0
    (torch_scatter's fill value)."""
    return _SegmentReduce.apply(msg, dst, int(num_targets), 0)


def segment_sum(msg: torch.Tensor, dst: torch.Tensor, num_targets: int) -> torch.Tensor:
    """aggr="add" of MessagePassing over edges sorted by target, summed in edge order (deterministic)."""
    return _SegmentReduce.apply(msg, dst, int(num_targets), 1)
