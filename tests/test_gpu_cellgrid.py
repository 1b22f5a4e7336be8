"""GPU parity of the cell-list searches (csrc/cellgrid.cu) against the oracle's brute-force restatements: the same
tables bit for bit as the sweep kernels -- ball query: first `nsample` hits in candidate index order with their fp32
squared distances, shadow padding; knn: k nearest by (d2, index) -- on shapes chosen to reach every branch: ragged and
empty clouds, clouds below the direct-sweep threshold next to large ones, duplicated points, more hits than the
per-warp buffer holds (selection of the lowest indices), exact-boundary distances under both membership rules, sparse
outliers (ring growth and the sweep fallback of knn), coordinates far from the origin (periodic lattice keys).
Reference semantics: neighbour_finder.py:34-37,46-47; SURVEY.md A.3 / A.6."""
import numpy as np
import pytest
import torch

import oracle
from torch_points3d_b200 import _cabi, ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module", autouse=True)
def _force_grid_path():
    import __graft_entry__
    __graft_entry__.build()
    old = ops.GRID_SEARCH_MIN_POINTS
    ops.GRID_SEARCH_MIN_POINTS = 0
    before = _cabi.launch_count()
    yield
    ops.GRID_SEARCH_MIN_POINTS = old
    assert _cabi.launch_count() > before


def _ragged(rng, sizes, scale=1.0, dup=0.05, offset=0.0):
    pts, batch = [], []
    for b, n in enumerate(sizes):
        p = (rng.random((n, 3)) * scale + offset).astype(np.float32)
        if n > 4 and dup > 0:
            k = max(1, int(n * dup))
            p[rng.integers(0, n, k)] = p[rng.integers(0, n, k)]
        pts.append(p)
        batch.append(np.full(n, b, dtype=np.int64))
    return np.concatenate(pts), np.concatenate(batch)


def _queries(rng, x, bx, sizes_q, jitter):
    qs, bq = [], []
    for b, n in enumerate(sizes_q):
        own = x[bx == b]
        if len(own) == 0 or n == 0:
            continue
        sel = own[rng.integers(0, len(own), n)] + rng.normal(0, jitter, (n, 3)).astype(np.float32)
        qs.append(sel.astype(np.float32))
        bq.append(np.full(n, b, dtype=np.int64))
    return np.concatenate(qs), np.concatenate(bq)


def _ptr(batch, num):
    return np.concatenate([[0], np.cumsum(np.bincount(batch, minlength=num))])


@pytest.mark.parametrize("seed,sizes,r,M", [
    (0, [3000, 1, 500, 0, 2200], 0.08, 16),        # ragged, a single-point cloud, an empty cloud, a small (swept) cloud
    (1, [6000], 0.05, 34),
    (2, [2500, 2500], 0.30, 40),                   # hundreds of hits per query: buffer compaction, lowest indices kept
    (3, [4000, 900], 0.02, 8),                     # mostly empty balls
    (4, [1500, 1500, 1500], 0.11, 64),
    (5, [5000], 1.5, 33),                          # every point is a hit
])
def test_ball_query_grid_bit_exact(seed, sizes, r, M):
    rng = np.random.default_rng(seed)
    x, bx = _ragged(rng, sizes)
    y, by = _queries(rng, x, bx, [max(1, s // 3) for s in sizes], 0.01)
    nc = len(sizes)
    want_i, want_d = oracle.ball_query_csr(r, M, x, y, _ptr(bx, nc), _ptr(by, nc))
    got_i, got_d = ops.ball_query(r, M, T(x), T(y), mode="partial_dense", batch_x=T(bx), batch_y=T(by))
    assert np.array_equal(N(got_i), want_i)
    assert np.array_equal(N(got_d), want_d)
    # without distances, and as an edge list (radius: strict membership, neighbour_finder.py:33)
    assert np.array_equal(N(ops.ball_query(r, M, T(x), T(y), mode="partial_dense", batch_x=T(bx), batch_y=T(by),
                                           return_dist=False)[0]), want_i)
    want_e = oracle.radius_csr(x, y, r, _ptr(bx, nc), _ptr(by, nc), max_num_neighbors=M, strict=True)
    assert np.array_equal(N(ops.radius(T(x), T(y), r, T(bx), T(by), max_num_neighbors=M)), want_e)


def test_ball_query_grid_boundary_and_far_coordinates():
    # lattice points: many distances are EXACTLY r^2 -> the strict / non-strict rules differ on them
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(12), indexing="ij"), -1).reshape(-1, 3)
    x = (g * 0.25).astype(np.float32)
    bx = np.zeros(len(x), dtype=np.int64)
    ptr = np.array([0, len(x)])
    for strict in (False, True):
        want = oracle.ball_query_csr(0.5, 40, x, x, ptr, ptr, strict=strict)
        out = torch.empty((len(x), 40), dtype=torch.int64, device=DEV)
        d2 = torch.empty((len(x), 40), dtype=torch.float32, device=DEV)
        ops._ball_query_batch(torch.device(DEV), T(x), T(x), T(bx), T(bx), 0.5, 40, int(strict), len(x), out, d2)
        assert np.array_equal(N(out), want[0]) and np.array_equal(N(d2), want[1])
    assert (oracle.ball_query_csr(0.5, 40, x, x, ptr, ptr, strict=False)[0] !=
            oracle.ball_query_csr(0.5, 40, x, x, ptr, ptr, strict=True)[0]).any()
    # clouds far from the origin (keys wrap around the 2^14-cell period; huge offsets collapse into one cell)
    rng = np.random.default_rng(7)
    for offset, r in ((2000.0, 0.05), (-3.0e5, 0.5), (1.0e9, 100.0)):
        x, bx = _ragged(rng, [1200, 800], scale=2.0, offset=offset)
        p = _ptr(bx, 2)
        want = oracle.ball_query_csr(r, 20, x, x, p, p)
        got = ops.ball_query(r, 20, T(x), T(x), mode="partial_dense", batch_x=T(bx), batch_y=T(bx))
        assert np.array_equal(N(got[0]), want[0]) and np.array_equal(N(got[1]), want[1])
    # no batch vectors = one cloud
    x, _ = _ragged(rng, [3000])
    out = torch.empty((3000, 12), dtype=torch.int64, device=DEV)
    ops._ball_query_batch(torch.device(DEV), T(x), T(x), None, None, 0.07, 12, 0, 3000, out, None)
    assert np.array_equal(N(out), oracle.ball_query_csr(0.07, 12, x, x, np.array([0, 3000]), np.array([0, 3000]))[0])


@pytest.mark.parametrize("seed,sizes,k", [
    (0, [3000, 5, 600, 0, 2000], 16),     # a cloud with fewer than k points (-1 fill), a swept cloud, an empty cloud
    (1, [7000], 1),
    (2, [2500, 2500], 32),
    (3, [5000], 3),
    (4, [1000, 4000], 7),
])
def test_knn_grid_bit_exact(seed, sizes, k):
    rng = np.random.default_rng(100 + seed)
    x, bx = _ragged(rng, sizes, dup=0.0)
    # a few far outliers per cloud: their queries need ring growth or the sweep
    for b, n in enumerate(sizes):
        if n > 100:
            own = np.nonzero(bx == b)[0]
            x[own[:5]] += rng.normal(0, 30.0, (5, 3)).astype(np.float32)
    y, by = _queries(rng, x, bx, [max(1, s // 2) for s in sizes], 0.05)
    y = np.concatenate([y, x[bx == 0][:5]])           # queries ON the outliers
    by = np.concatenate([by, np.zeros(5, dtype=np.int64)])
    order = np.argsort(by, kind="stable")
    y, by = y[order], by[order]
    nc = len(sizes)
    want_c, want_d = oracle.knn_csr(x, y, k, _ptr(bx, nc), _ptr(by, nc))
    got_c, got_d = ops.knn_table(T(x), T(y), k, T(bx), T(by))
    assert np.array_equal(N(got_d), want_d)
    assert np.array_equal(N(got_c), want_c)           # distinct random points: no exact ties


def test_knn_grid_ties_shapes_and_edges():
    # duplicated points and a regular lattice: exact distance ties -> lower index first, like the oracle
    g = np.stack(np.meshgrid(np.arange(10), np.arange(10), np.arange(10), indexing="ij"), -1).reshape(-1, 3)
    x = np.concatenate([g, g[:200]]).astype(np.float32) * 0.5
    bx = np.zeros(len(x), dtype=np.int64)
    ptr = np.array([0, len(x)])
    want_c, want_d = oracle.knn_csr(x, x[:300], 9, ptr, np.array([0, 300]))
    got_c, got_d = ops.knn_table(T(x), T(x[:300]), 9, T(bx), T(bx[:300]))
    assert np.array_equal(N(got_d), want_d) and np.array_equal(N(got_c), want_c)
    # flat (planar) and line-shaped clouds, far from the origin
    rng = np.random.default_rng(9)
    flat = (rng.random((4000, 3)) * np.array([10, 10, 0])).astype(np.float32) + 500.0
    line = (rng.random((3000, 3)) * np.array([50, 0, 0])).astype(np.float32) - 1000.0
    for cloud in (flat, line):
        p = np.array([0, len(cloud)])
        q = cloud[::7]
        want_c, want_d = oracle.knn_csr(cloud, q, 8, p, np.array([0, len(q)]))
        got_c, got_d = ops.knn_table(T(cloud), T(q), 8)
        assert np.array_equal(N(got_d), want_d)
        same = N(got_c) == want_c
        assert same.all() or all(len(set(want_d[r].tolist())) < 8 for r in np.nonzero(~same)[0])
    # edge list form and k = 1 interpolation tables (FPModule_PD / KNNInterpolate, interpolate.py:25)
    pos, batch = synthetic.indoor_spheres(num_spheres=2, raw_points=20000, radius=0.6, seed=5)
    coarse, cb = pos[::4], batch[::4]
    p_c, p_f = _ptr(cb, 2), _ptr(batch, 2)
    row, col = oracle.knn_edges(coarse, pos, 1, p_c, p_f)
    e = N(ops.knn_edges(T(coarse), T(pos), 1, T(cb), T(batch)))
    assert np.array_equal(e, np.stack([row, col]))


def test_grid_and_sweep_kernels_agree_at_c3_scale():
    """the C3 level-0 shape (50k points in 8 spheres, r = 0.1, M = 34): the cell-list table equals the sweep kernel's"""
    pos, batch = synthetic.indoor_spheres()
    tp, tb = T(pos), T(batch)
    grid = ops.ball_query(0.1, 34, tp, tp, mode="partial_dense", batch_x=tb, batch_y=tb)
    ops.GRID_SEARCH_MIN_POINTS = 1 << 40
    try:
        sweep = ops.ball_query(0.1, 34, tp, tp, mode="partial_dense", batch_x=tb, batch_y=tb)
        k_sweep = ops.knn_table(tp[::4].contiguous(), tp, 1, tb[::4].contiguous(), tb)
    finally:
        ops.GRID_SEARCH_MIN_POINTS = 0
    assert torch.equal(grid[0], sweep[0]) and torch.equal(grid[1], sweep[1])
    k_grid = ops.knn_table(tp[::4].contiguous(), tp, 1, tb[::4].contiguous(), tb)
    assert torch.equal(k_grid[0], k_sweep[0]) and torch.equal(k_grid[1], k_sweep[1])
