"""GPU parity tests: every operator of libp3d_b200.so, called through the public API (ops.py ->
ctypes -> C ABI), against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): FPS / ball-query / three_nn / knn indices bit-exact (fp32
squared distances in the same operation order); grouped and interpolated features bit-exact
(same fp32 operation order); scatter-type gradients and KPConv within 1e-5 relative (fp32
summation order is unspecified upstream).  Large-size cases use size-independent properties.
"""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from torch_points3d_b200 import _cabi, kpconv, ops, pointnet2, spatial_ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5  # relative to max|reference| (north_star: 1e-5 fp32)
RTOL_TENSOR = 2e-5  # KPConv forward with the contraction on tcgen05 (bf16 hi+lo split operands, fp32 accumulate)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module", autouse=True)
def _lib_loaded():
    import __graft_entry__
    __graft_entry__.build()
    assert torch.cuda.is_available()
    before = _cabi.launch_count()
    yield
    assert _cabi.launch_count() > before, "no kernel of libp3d_b200.so was launched"


# ------------------------------------------------------------------------------ FPS ---------
def test_fps_reference_known_answer():
    # reference test/test_fps.py:35-42
    pos = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.4, 0.2, 0], [2, 2, 2], [-1, -2, -0.01]], dtype=np.float32)
    assert N(ops.furthest_point_sample(T(pos[None]), 3)).tolist() == [[0, 3, 4]]
    idx = ops.fps(T(pos), torch.zeros(5, dtype=torch.long, device=DEV), ratio=3.0 / 5.0, random_start=False)
    assert N(idx).tolist() == [0, 3, 4]


@pytest.mark.parametrize("B,n,S", [(16, 1024, 512), (16, 2048, 512), (3, 257, 64), (2, 5000, 700), (1, 33, 33),
                                   (2, 100, 1), (4, 8191, 300), (2, 12000, 257), (1, 40000, 300),
                                   (1, 120000, 128)])
def test_fps_dense_bit_exact(B, n, S):
    xyz = synthetic.surface_clouds(B, n, seed=100 + n)
    got = N(ops.furthest_point_sample(T(xyz), S))
    assert got.dtype == np.int64
    assert np.array_equal(got, oracle.fps_dense(xyz, S))


def test_fps_fused_coordinates():
    xyz = synthetic.surface_clouds(4, 3000, seed=9)
    for S in (1, 257):
        idx, new_xyz = ops.furthest_point_sample(T(xyz), S, return_xyz=True)
        assert np.array_equal(N(idx), oracle.fps_dense(xyz, S))
        assert np.array_equal(N(new_xyz), np.take_along_axis(xyz, N(idx)[..., None], 1))
    big = synthetic.surface_clouds(1, 70000, seed=10)      # cluster path
    idx, new_xyz = ops.furthest_point_sample(T(big), 40, return_xyz=True)
    assert np.array_equal(N(new_xyz), np.take_along_axis(big, N(idx)[..., None], 1))


def test_fps_ties_and_duplicates():
    dup = np.zeros((1, 6, 3), dtype=np.float32)
    dup[0, 2:4] = [1, 0, 0]
    dup[0, 4:6] = [-1, 0, 0]
    assert N(ops.furthest_point_sample(T(dup), 3)).tolist() == [[0, 2, 4]]
    # all points identical: every round is an all-way tie -> lowest index not yet... (mind = 0 everywhere)
    same = np.ones((2, 300, 3), dtype=np.float32)
    assert np.array_equal(N(ops.furthest_point_sample(T(same), 5)), oracle.fps_dense(same, 5))
    # integer lattice: many exact ties
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8)), -1).reshape(1, -1, 3).astype(np.float32)
    assert np.array_equal(N(ops.furthest_point_sample(T(g), 200)), oracle.fps_dense(g, 200))


def test_fps_csr_ragged_clouds():
    rng = np.random.default_rng(5)
    sizes = [1, 700, 3, 2049, 64, 9000]
    pos = rng.standard_normal((sum(sizes), 3)).astype(np.float32)
    batch = np.repeat(np.arange(len(sizes)), sizes)
    ptr = oracle.batch_to_ptr(batch)
    for ratio in (0.25, 0.5, 1.0):
        got = N(ops.fps(T(pos), T(batch), ratio=ratio, random_start=False))
        counts = oracle.fps_counts(ptr, ratio)
        assert np.array_equal(got, oracle.fps_csr(pos, ptr, counts))
    # random start: still a valid FPS run from the drawn start point
    torch.manual_seed(0)
    got = N(ops.fps(T(pos), T(batch), ratio=0.25, random_start=True))
    counts = oracle.fps_counts(ptr, 0.25)
    out_ptr = np.concatenate([[0], np.cumsum(counts)])
    start = got[out_ptr[:-1]] - ptr[:-1]
    assert ((start >= 0) & (start < np.array(sizes))).all()
    assert np.array_equal(got, oracle.fps_csr(pos, ptr, counts, start=start))


def test_fps_sampler_classes():
    xyz = synthetic.surface_clouds(2, 500, seed=3)
    idx = spatial_ops.DenseFPSSampler(num_to_sample=100)(T(xyz))
    assert np.array_equal(N(idx), oracle.fps_dense(xyz, 100))
    idx = spatial_ops.DenseFPSSampler(ratio=0.25)(T(xyz))
    assert idx.shape == (2, 125)
    pos = xyz.reshape(-1, 3)
    batch = np.repeat(np.arange(2), 500)
    idx = spatial_ops.FPSSampler(ratio=0.5, random_start=False)(T(pos), batch=T(batch))
    assert idx.shape[0] == 500  # reference test/test_sampler.py:16-23 style count check
    idx = spatial_ops.FPSSampler(num_to_sample=5)(T(pos[:10]), batch=torch.zeros(10, dtype=torch.long, device=DEV))
    assert idx.shape[0] == 5


# ----------------------------------------------------------------------- ball query ---------
@pytest.mark.parametrize("B,n,S,r,K", [(16, 1024, 512, 0.2, 32), (16, 2048, 512, 0.1, 32), (16, 2048, 512, 0.4, 128),
                                       (16, 512, 128, 0.8, 128), (3, 333, 77, 0.3, 5), (2, 4100, 130, 0.05, 16),
                                       (1, 1, 1, 0.1, 4), (2, 50, 50, 10.0, 64)])
def test_ball_query_dense_bit_exact(B, n, S, r, K):
    x = synthetic.surface_clouds(B, n, seed=200 + n)
    y = x[:, ::max(n // S, 1)][:, :S].copy()
    y[:, ::7] += 5.0  # some empty balls -> all-zero rows
    idx, d2 = ops.ball_query(r, K, T(x), T(y))
    oidx, od2 = oracle.ball_query_dense(r, K, x, y)
    assert np.array_equal(N(idx), oidx)
    assert np.array_equal(N(d2), od2)
    idx2, none = ops.ball_query(r, K, T(x), T(y), return_dist=False)
    assert none is None and np.array_equal(N(idx2), oidx)


def test_ball_query_dense_spec_cases():
    x = np.array([[[0, 0, 0], [0.1, 0, 0], [0.5, 0, 0], [0.05, 0, 0], [0.2, 0, 0]]], dtype=np.float32)
    y = np.array([[[0, 0, 0], [10, 10, 10], [0.5, 0, 0]]], dtype=np.float32)
    idx, d2 = ops.ball_query(0.2, 4, T(x), T(y))
    assert N(idx)[0].tolist() == [[0, 1, 3, 0], [0, 0, 0, 0], [2, 2, 2, 2]]
    assert N(d2)[0, 1].tolist() == [-1, -1, -1, -1]
    # exact boundary d2 == r2: dense is strict, partial_dense is not
    xb = np.array([[[0.25, 0, 0]]], dtype=np.float32)
    yb = np.zeros((1, 1, 3), dtype=np.float32)
    assert N(ops.ball_query(0.25, 1, T(xb), T(yb))[1])[0, 0, 0] == -1
    z = torch.zeros(1, dtype=torch.long, device=DEV)
    assert N(ops.ball_query(0.25, 1, T(xb[0]), T(yb[0]), mode="partial_dense", batch_x=z, batch_y=z)[0])[0, 0] == 0


@pytest.mark.parametrize("B,n,S,radii,ks", [(16, 2048, 512, [0.1, 0.2, 0.4], [32, 64, 128]), (16, 512, 128, [0.4, 0.8], [64, 128]),
                                            (3, 2500, 77, [0.3, 0.05, 0.6, 0.2], [16, 5, 200, 1]), (2, 40, 9, [0.5], [7])])
def test_ball_query_multi_scale_equals_single_scale(B, n, S, radii, ks):
    """one sweep for all scales of an MSG layer == one single-scale query per radius, bit for bit (candidate tiles
    of 2048 points: n = 2500 crosses a tile boundary; unsorted radii; rows that never fill)"""
    xyz = synthetic.surface_clouds(B, n, seed=33)
    pos = T(xyz)
    ctr = pos[:, torch.linspace(0, n - 1, S, device=DEV).long()].contiguous()
    multi = ops.ball_query_multi(radii, ks, pos, ctr)
    for r, k, got in zip(radii, ks, multi):
        assert torch.equal(got, ops.ball_query(r, k, pos, ctr)[0]), (r, k)
        assert np.array_equal(N(got), oracle.ball_query_dense(r, k, xyz, N(ctr))[0])


def test_ball_query_partial_reference_grid():
    # reference test/test_transform.py:67-113: 10x10 grid, widths 100/150/200, shadow = N
    n = 10
    xv, yv = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n))
    pos = np.stack([xv.flatten(), yv.flatten(), np.zeros(n * n)], 1).astype(np.float32)
    ptr = np.array([0, n * n])
    z = torch.zeros(n * n, dtype=torch.long, device=DEV)
    for r, width in [(0.5, 100), (0.5, 150), (1.0, 200)]:
        idx, d2 = ops.ball_query(r, width, T(pos), T(pos), mode="partial_dense", batch_x=z, batch_y=z)
        oidx, od2 = oracle.ball_query_csr(r, width, pos, pos, ptr, ptr)
        assert np.array_equal(N(idx), oidx) and np.array_equal(N(d2), od2)
        assert (N(idx) == n * n).any()  # padded with the shadow index x.shape[0]


def test_ball_query_partial_ragged_bit_exact():
    rng = np.random.default_rng(11)
    sizes_x = [1, 500, 3, 2500, 40, 4097]
    sizes_y = [1, 100, 3, 2500, 1, 900]
    x = rng.uniform(0, 1, (sum(sizes_x), 3)).astype(np.float32)
    y = rng.uniform(0, 1, (sum(sizes_y), 3)).astype(np.float32)
    bx, by = np.repeat(np.arange(6), sizes_x), np.repeat(np.arange(6), sizes_y)
    px, py = oracle.batch_to_ptr(bx), oracle.batch_to_ptr(by)
    for r, M in [(0.08, 16), (0.2, 34), (0.02, 8)]:
        idx, d2 = ops.ball_query(r, M, T(x), T(y), mode="partial_dense", batch_x=T(bx), batch_y=T(by))
        oidx, od2 = oracle.ball_query_csr(r, M, x, y, px, py)
        assert np.array_equal(N(idx), oidx)
        assert np.array_equal(N(d2), od2)
    # RadiusNeighbourFinder (blocks.py:53) routes here
    nf = spatial_ops.RadiusNeighbourFinder(0.2, 34, conv_type="partial_dense")
    assert np.array_equal(N(nf(T(x), T(y), T(bx), T(by))), oracle.ball_query_csr(0.2, 34, x, y, px, py)[0])


def test_ball_query_c3_spheres_bit_exact():
    pos, batch = synthetic.indoor_spheres()
    ptr = oracle.batch_to_ptr(batch)
    idx, _ = ops.ball_query(0.1, 34, T(pos), T(pos), mode="partial_dense", batch_x=T(batch), batch_y=T(batch))
    assert np.array_equal(N(idx), oracle.ball_query_csr(0.1, 34, pos, pos, ptr, ptr)[0])


def test_radius_edge_list_reference_counts():
    # reference test/test_sampler.py:36-86
    x = np.array([[-1, -1, 0], [-1, 1, 0], [1, -1, 0], [1, 1, 0]], dtype=np.float32)
    y = np.array([[-1, 0, 0], [1, 0, 0]], dtype=np.float32)
    bx, by = torch.zeros(4, dtype=torch.long, device=DEV), torch.zeros(2, dtype=torch.long, device=DEV)
    for r, maxn, edges in [(1.1, 4, 4), (10, 4, 8), (10, 3, 6), (1.1, 3, 4)]:
        e = ops.radius(T(x), T(y), r, bx, by, max_num_neighbors=maxn)
        assert tuple(e.shape) == (2, edges)
        oe = oracle.radius_csr(x, y, r, np.array([0, 4]), np.array([0, 2]), max_num_neighbors=maxn, strict=True)
        assert np.array_equal(N(e), oe)
    f = spatial_ops.MultiscaleRadiusNeighbourFinder([1.1, 10], [4, 3])
    assert f(T(x), T(y), bx, by, 0).shape[1] == 4 and f(T(x), T(y), bx, by, 1).shape[1] == 6


# ------------------------------------------------------------------------------ knn ---------
def _knn_sets_equal_modulo_ties(col, d2, ocol, od2):
    if np.array_equal(col, ocol):
        return True
    bad = np.nonzero((col != ocol).any(1))[0]
    for q in bad:  # differing rows must have identical distance lists (exact ties only)
        if not np.array_equal(d2[q], od2[q]):
            return False
    return True


@pytest.mark.parametrize("k", [1, 3, 16, 40])
def test_knn_batch_bit_exact(k):
    rng = np.random.default_rng(21)
    sizes_x = [50, 3000, 2, 700, 4200]
    sizes_y = [10, 600, 5, 700, 33]
    x = rng.standard_normal((sum(sizes_x), 3)).astype(np.float32)
    y = rng.standard_normal((sum(sizes_y), 3)).astype(np.float32)
    bx, by = np.repeat(np.arange(5), sizes_x), np.repeat(np.arange(5), sizes_y)
    col, d2 = ops.knn_table(T(x), T(y), k, T(bx), T(by))
    ocol, od2 = oracle.knn_csr(x, y, k, oracle.batch_to_ptr(bx), oracle.batch_to_ptr(by))
    assert np.array_equal(N(d2), od2)
    assert _knn_sets_equal_modulo_ties(N(col), N(d2), ocol, od2)
    e = ops.knn_edges(T(x), T(y), k, T(bx), T(by))
    oe = oracle.knn_edges(x, y, k, oracle.batch_to_ptr(bx), oracle.batch_to_ptr(by))
    assert e.shape == oe.shape and np.array_equal(N(e)[0], oe[0])
    assert np.array_equal(N(spatial_ops.KNNNeighbourFinder(k)(T(x), T(y), T(bx), T(by)))[0], oe[0])


def test_knn_reference_cases():
    # reference test/test_interpolateop.py:10-23 (1-NN mapping) + tie / short-cloud cases of the oracle tests
    pos = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 0], [0.1, 0, 0]], dtype=np.float32)
    query = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [0.05, 0, 0]], dtype=np.float32)
    e = N(ops.knn_edges(T(query), T(pos), 1))
    assert e[0].tolist() == [0, 1, 2, 3, 4] and e[1][4] == e[1][3]
    x = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [5, 5, 5]], dtype=np.float32)
    y = np.array([[0, 0, 0], [5, 5, 4]], dtype=np.float32)
    col, d2 = ops.knn_table(T(x), T(y), 3, T(np.array([0, 0, 0, 1])), T(np.array([0, 1])))
    assert N(col).tolist() == [[0, 1, 2], [3, -1, -1]]
    assert N(d2)[1].tolist() == [1.0, -1.0, -1.0]
    e = N(ops.knn_edges(T(x), T(y), 3, T(np.array([0, 0, 0, 1])), T(np.array([0, 1]))))
    assert e.tolist() == [[0, 0, 0, 1], [0, 1, 2, 3]]


def test_knn_c4_shape_properties():
    """C4 size (120k supports): properties instead of the full oracle -- ascending distances,
    the k-th distance bounds a random sample of other candidates, the query's own point is first."""
    scans = synthetic.lidar_scans(B=1)
    x = scans[0]
    idx = oracle.fps_dense(x[None, :20000], 64)[0]
    xs = T(x)
    q_idx = N(ops.furthest_point_sample(xs[None], 4096))[0]
    y = x[q_idx]
    col, d2 = ops.knn_table(xs, T(y), 16)
    col, d2 = N(col), N(d2)
    assert (np.diff(d2, axis=1) >= 0).all()
    assert (col[:, 0] == q_idx).all() or (d2[:, 0] == 0).all()
    sub = slice(0, 256)
    ocol, od2 = oracle.knn_csr(x, y[sub], 16, np.array([0, x.shape[0]]), np.array([0, 256]))
    assert np.array_equal(d2[sub], od2)
    assert _knn_sets_equal_modulo_ties(col[sub], d2[sub], ocol, od2)
    assert idx.shape == (64,)


def test_knn_dense_and_three_nn():
    rng = np.random.default_rng(31)
    for B, n, m in [(16, 512, 128), (16, 2048, 512), (2, 77, 3), (1, 5, 9)]:
        unknown = rng.standard_normal((B, n, 3)).astype(np.float32)
        known = rng.standard_normal((B, m, 3)).astype(np.float32)
        dist, idx = ops.three_nn(T(unknown), T(known))
        odist, oidx = oracle.three_nn(unknown, known)
        assert np.array_equal(N(idx), oidx)
        assert np.array_equal(N(dist), odist)  # correctly rounded sqrt on both sides
    dist, idx, w = ops.three_nn(T(unknown), T(known), return_weight=True)
    recip = 1.0 / (dist + 1e-8)                      # core/base_conv/dense.py:137-139
    assert torch.allclose(w, recip / recip.sum(2, keepdim=True), rtol=1e-6, atol=0)
    known[:, 1] = known[:, 0]  # duplicated known point: tie -> lower index first
    dist, idx = ops.three_nn(T(unknown), T(known))
    assert np.array_equal(N(idx), oracle.three_nn(unknown, known)[1])
    idx, d2 = ops.knn_dense(T(known), T(unknown), 4)
    order = np.argsort(((unknown[:, :, None] - known[:, None]) ** 2).sum(-1), -1, kind="stable")[..., :4]
    assert np.array_equal(N(idx), order)


# ------------------------------------------------------------------------- grouping ---------
@pytest.mark.parametrize("B,C,n,S,K", [(16, 64, 1024, 512, 32), (4, 320, 512, 128, 64), (2, 3, 2048, 512, 128),
                                       (3, 5, 17, 6, 4), (1, 1, 9, 7, 3), (2, 7, 30000, 33, 5)])
def test_grouping_forward_bit_exact_and_backward(B, C, n, S, K):
    rng = np.random.default_rng(41)
    feat = rng.standard_normal((B, C, n)).astype(np.float32)
    idx = rng.integers(0, n, (B, S, K))
    f = T(feat).requires_grad_(True)
    out = ops.grouping_operation(f, T(idx))
    assert np.array_equal(N(out), oracle.group_fwd(feat, idx))
    dout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(T(dout))
    assert relerr(N(f.grad), oracle.group_bwd(dout, idx, n)) <= RTOL


@pytest.mark.parametrize("B,C,n,S,K,r", [(3, 70, 500, 37, 29, 0.3), (16, 320, 512, 128, 128, 0.8), (1, 16, 1024, 4, 3, 0.2),
                                         (5, 33, 700, 64, 48, 0.25), (2, 40, 64, 300, 16, 0.5), (149, 32, 96, 8, 64, 0.4),
                                         (16, 3, 2048, 512, 64, 0.2), (4, 6, 1024, 256, 32, 0.2), (3, 8, 700, 37, 29, 0.3),
                                         (1, 1, 6144, 700, 9, 0.15), (200, 2, 50, 3, 5, 0.6)])
def test_grouping_backward_on_ball_query_tables(B, C, n, S, K, r):
    """backward kernels (sorted lists with run compression for wide tables, per-warp shared-memory tables for the
    narrow first-level ones) on real ball-query tables: rows padded with their first hit give long runs of one
    index; sizes that are not multiples of the chunk or the warp step; the result is deterministic"""
    xyz = synthetic.surface_clouds(B, n, seed=60 + C)
    pos = T(xyz)
    new_pos = pos[:, torch.linspace(0, n - 1, S, device=DEV).long()].contiguous()
    idx = ops.ball_query(r, K, pos, new_pos)[0]
    rng = np.random.default_rng(43)
    feat = T(rng.standard_normal((B, C, n)).astype(np.float32)).requires_grad_(True)
    out = ops.grouping_operation(feat, idx)
    dout = T(rng.standard_normal(tuple(out.shape)).astype(np.float32))
    (g1,) = torch.autograd.grad(out, feat, dout, retain_graph=True)
    (g2,) = torch.autograd.grad(out, feat, dout)
    assert torch.equal(g1, g2)
    assert relerr(N(g1), oracle.group_bwd(N(dout), N(idx), n)) <= RTOL


def test_group_features_matches_reference_formulation():
    """fused _prepare_features == the reference's op-by-op chain (modules/pointnet2/dense.py:36-54)"""
    B, n, S, K, C = 4, 1024, 256, 32, 16
    xyz = synthetic.surface_clouds(B, n, seed=51)
    feat = np.random.default_rng(51).standard_normal((B, C, n)).astype(np.float32)
    pos = T(xyz)
    idx = ops.furthest_point_sample(pos, S)
    new_pos = pos.gather(1, idx.unsqueeze(-1).repeat(1, 1, 3))
    ridx = ops.ball_query(0.2, K, pos, new_pos)[0]
    for normalize in (None, 0.2):
        x1 = T(feat).requires_grad_(True)
        fused = ops.group_features(x1, pos, new_pos, ridx, normalize_by=normalize)
        x2 = T(feat).requires_grad_(True)
        gp = ops.grouping_operation(pos.transpose(1, 2).contiguous(), ridx) - new_pos.transpose(1, 2).unsqueeze(-1)
        if normalize:
            gp = gp / normalize
        ref = torch.cat([gp, ops.grouping_operation(x2, ridx)], dim=1)
        assert torch.equal(fused[:, 3:], ref[:, 3:])
        assert torch.equal(fused[:, :3], ref[:, :3]), normalize
        g = torch.randn_like(ref)
        fused.backward(g)
        ref.backward(g)
        assert relerr(N(x1.grad), N(x2.grad)) <= RTOL
    # numpy restatement of the same chain
    gi = N(ridx)
    want = np.stack([xyz[b][gi[b]] - N(new_pos)[b][:, None] for b in range(B)]).transpose(0, 3, 1, 2)
    plain = ops.group_features(T(feat), pos, new_pos, ridx)
    assert np.array_equal(N(plain)[:, :3], want)
    assert np.array_equal(N(plain)[:, 3:], oracle.group_fwd(feat, gi))


@pytest.mark.parametrize("S,K,C", [(37, 29, 5), (64, 30, 16), (50, 6, 40)])
def test_group_features_fused_rows_crossing_vector_boundaries(S, K, C):
    """fused relative-coordinate + feature grouping when a 4-slot vector straddles two centroids (K % 4 != 0) and
    when the slot count is not a multiple of 4 (scalar tail path)"""
    B, n = 3, 777
    xyz = synthetic.surface_clouds(B, n, seed=52)
    feat = np.random.default_rng(53).standard_normal((B, C, n)).astype(np.float32)
    pos = T(xyz)
    new_pos = pos[:, torch.linspace(0, n - 1, S, device=DEV).long()].contiguous()
    ridx = ops.ball_query(0.3, K, pos, new_pos)[0]
    gi = N(ridx)
    for normalize in (None, 0.3):
        fused = N(ops.group_features(T(feat), pos, new_pos, ridx, normalize_by=normalize))
        want = np.stack([xyz[b][gi[b]] - N(new_pos)[b][:, None] for b in range(B)]).transpose(0, 3, 1, 2)
        if normalize:
            want = want * (np.float32(1.0) / np.float32(normalize))
        assert np.array_equal(fused[:, :3], want.astype(np.float32))
        assert np.array_equal(fused[:, 3:], oracle.group_fwd(feat, gi))


# -------------------------------------------------------------------- interpolation ---------
@pytest.mark.parametrize("B,C,m,n", [(16, 256, 128, 512), (16, 128, 512, 2048), (2, 5, 7, 9), (1, 33, 40000, 100),
                                     (3, 100, 77, 1001), (150, 64, 16, 40), (5, 40, 512, 300)])
def test_three_interpolate_bit_exact_and_backward(B, C, m, n):
    rng = np.random.default_rng(61)
    feat = rng.standard_normal((B, C, m)).astype(np.float32)
    idx = rng.integers(0, m, (B, n, 3))
    w = rng.random((B, n, 3)).astype(np.float32)
    f = T(feat).requires_grad_(True)
    out = ops.three_interpolate(f, T(idx), T(w))
    assert np.array_equal(N(out), oracle.three_interpolate_fwd(feat, idx, w))
    g = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(T(g))
    assert relerr(N(f.grad), oracle.three_interpolate_bwd(g, idx, w, m)) <= RTOL


def test_dense_fp_module_chain():
    """DenseFPModule.conv (core/base_conv/dense.py:132-144): three_nn -> 1/(d+1e-8) -> interpolate"""
    rng = np.random.default_rng(62)
    B, n, m, C = 4, 512, 128, 64
    pos_skip = rng.standard_normal((B, n, 3)).astype(np.float32)
    pos = pos_skip[:, :m].copy()
    x = rng.standard_normal((B, C, m)).astype(np.float32)
    mod = pointnet2.DenseFPModule([C, 32], bn=False).to(DEV)
    got = mod.conv(T(pos), T(pos_skip), T(x))
    dist, idx = oracle.three_nn(pos_skip, pos)
    recip = (1.0 / (torch.from_numpy(dist) + 1e-8))
    weight = (recip / recip.sum(2, keepdim=True)).numpy()
    assert np.allclose(N(got), oracle.three_interpolate_fwd(x, idx, weight), rtol=1e-5, atol=1e-6)
    assert np.allclose(N(got)[:, :, :m], x, rtol=1e-4, atol=1e-4)  # known points interpolate to themselves


def test_knn_interpolate_matches_reference_formula():
    rng = np.random.default_rng(63)
    pos_x = rng.standard_normal((300, 3)).astype(np.float32)
    pos_y = rng.standard_normal((900, 3)).astype(np.float32)
    bx, by = np.repeat([0, 1], [100, 200]), np.repeat([0, 1], [400, 500])
    x = rng.standard_normal((300, 8)).astype(np.float32)
    got = N(ops.knn_interpolate(T(x), T(pos_x), T(pos_y), T(bx), T(by), k=3))
    col, d2 = oracle.knn_csr(pos_x, pos_y, 3, oracle.batch_to_ptr(bx), oracle.batch_to_ptr(by))
    w = 1.0 / np.maximum(d2.astype(np.float64), 1e-16)
    want = (x[col] * w[..., None]).sum(1) / w.sum(1, keepdims=True)
    assert relerr(got, want) <= 1e-5
    # KNNInterpolate precompute path == direct path (reference test/test_interpolateop.py:25-35)
    from types import SimpleNamespace as NS
    q = NS(pos=T(pos_x), x=T(x), batch=T(bx))
    s = NS(pos=T(pos_y), batch=T(by))
    up = spatial_ops.KNNInterpolate(3)
    pre = up.precompute(q, s)
    assert relerr(N(up(q, s, precomputed=pre)), N(up(q, s))) <= 1e-5


# --------------------------------------------------------------------------- KPConv ---------
def _golden(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "kpconv_*.npz")))


@pytest.mark.parametrize("name", ["kpconv_linear_sum", "kpconv_gaussian_sum", "kpconv_constant_sum",
                                  "kpconv_linear_closest", "kpconv_linear_sum_c32"])
def test_kpconv_against_reference_golden(golden_dir, name):
    """outputs and autograd gradients of the REFERENCE's KPConv_ops (tests/golden/make_golden.py)"""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    x = T(z["x"]).requires_grad_(True)
    W = T(z["W"]).requires_grad_(True)
    out = ops.KPConv_ops(T(z["query"]), T(z["support"]), T(z["neighbors"]), x, T(z["K_points"]), W,
                         float(z["extent"]), str(z["influence"]), str(z["aggregation"]), precision="fp32")
    assert relerr(N(out), z["out"]) <= RTOL
    out_t = ops.KPConv_ops(T(z["query"]), T(z["support"]), T(z["neighbors"]), x, T(z["K_points"]), W,
                           float(z["extent"]), str(z["influence"]), str(z["aggregation"]), precision="tensor")
    assert relerr(N(out_t), z["out"]) <= RTOL_TENSOR
    if str(z["influence"]) != "constant":
        assert (N(out)[0] == 0).all()  # all-shadow row
    out.backward(T(z["dout"]))
    assert relerr(N(x.grad), z["dx"]) <= RTOL
    assert relerr(N(W.grad), z["dW"]) <= RTOL


@pytest.mark.parametrize("cin,cout,M,influence,agg", [(5, 64, 25, "linear", "sum"), (32, 32, 34, "linear", "sum"),
                                                      (64, 64, 30, "linear", "sum"), (128, 128, 34, "linear", "sum"),
                                                      (33, 47, 19, "gaussian", "sum"), (16, 8, 40, "linear", "closest"),
                                                      (256, 256, 34, "linear", "sum"), (8, 8, 128, "constant", "sum"),
                                                      (32, 64, 40, "linear", "closest"), (32, 32, 25, "gaussian", "sum"),
                                                      (64, 32, 38, "constant", "sum"), (32, 96, 7, "linear", "sum")])
def test_kpconv_vs_oracle(cin, cout, M, influence, agg):
    rng = np.random.default_rng(cin * 1000 + cout)
    pos, batch = synthetic.indoor_spheres(num_spheres=2, raw_points=4000, radius=0.35, seed=71)
    n0 = pos.shape[0]
    q = pos[::2] + rng.normal(0, 0.002, (pos[::2].shape)).astype(np.float32)
    qb = batch[::2]
    nb = N(ops.ball_query(0.1, M, T(pos), T(q), mode="partial_dense", batch_x=T(batch), batch_y=T(qb))[0])
    nb[5, :] = -1          # gathering.py:10 convention
    nb[6, ::2] = n0        # shadow
    x = rng.standard_normal((n0, cin)).astype(np.float32)
    W = (rng.standard_normal((15, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    kp = kpconv.load_kernels(1.5 * 0.04)
    tx, tW = T(x).requires_grad_(True), T(W).requires_grad_(True)
    out = ops.KPConv_ops(T(q), T(pos), T(nb), tx, T(kp), tW, 0.04, influence, agg, precision="fp32")
    ref = oracle.kpconv_fwd(q, pos, nb, x, kp, W, 0.04, influence, agg)
    assert relerr(N(out), ref) <= RTOL
    with torch.no_grad():
        out_t = ops.KPConv_ops(T(q), T(pos), T(nb), tx, T(kp), tW, 0.04, influence, agg, precision="tensor")
    assert relerr(N(out_t), ref) <= RTOL_TENSOR
    dout = rng.standard_normal(ref.shape).astype(np.float32)
    out.backward(T(dout))
    dx, dW = oracle.kpconv_bwd(q, pos, nb, x, kp, W, dout, 0.04, influence, agg)
    assert relerr(N(tx.grad), dx) <= RTOL
    assert relerr(N(tW.grad), dW) <= RTOL


@pytest.mark.parametrize("n,N0,M,cin,cout,influence,agg", [(114, 464, 38, 256, 256, "linear", "sum"), (464, 2580, 34, 128, 128, "linear", "sum"),
                                                           (300, 900, 20, 64, 96, "gaussian", "closest"), (64, 64, 45, 70, 64, "linear", "sum")])
def test_kpconv_split_path_small_point_sets(n, N0, M, cin, cout, influence, agg):
    """deep-level shapes: aggregation kernel + library GEMM (exact fp32) against the oracle, forward and gradients"""
    rng = np.random.default_rng(n + cin)
    sup = (rng.random((N0, 3)) * 0.6).astype(np.float32)
    qry = sup[rng.permutation(N0)[:n]]
    d = ((qry[:, None] - sup[None]) ** 2).sum(-1)
    nb = np.argsort(d, 1)[:, :M].astype(np.int64)
    nb[rng.random((n, M)) < 0.2] = N0
    Kp = kpconv.load_kernels(0.09, random_rotation=False).astype(np.float32)
    x = rng.standard_normal((N0, cin)).astype(np.float32)
    W = (rng.standard_normal((15, cin, cout)) * 0.1).astype(np.float32)
    dout = rng.standard_normal((n, cout)).astype(np.float32)
    tx, tW = T(x).requires_grad_(True), T(W).requires_grad_(True)
    ref = oracle.kpconv_fwd(qry, sup, nb, x, Kp, W, 0.06, influence, agg)
    dx, dW = oracle.kpconv_bwd(qry, sup, nb, x, Kp, W, dout, 0.06, influence, agg)
    # default path: the fused kernel, slabs of the contraction / output channels spread over the SMs
    assert ops.KPCONV_SPLIT_MAX_POINTS == 0
    out = ops.KPConv_ops(T(qry), T(sup), T(nb), tx, T(Kp), tW, 0.06, influence, agg)
    assert out.grad_fn.name().startswith("_KPConv") and not out.grad_fn.name().startswith("_KPConvSplit")
    assert relerr(N(out), ref) <= (RTOL_TENSOR if cin % 32 == 0 and cout % 32 == 0 and M <= 40 else RTOL)
    out.backward(T(dout))
    assert relerr(N(tx.grad), dx) <= RTOL and relerr(N(tW.grad), dW) <= RTOL
    again = ops.KPConv_ops(T(qry), T(sup), T(nb), tx, T(Kp), tW, 0.06, influence, agg)
    assert torch.equal(out, again)          # split contractions add their partial sums in a fixed order
    # opt-in: aggregation kernel + library GEMM (exact fp32)
    ops.KPCONV_SPLIT_MAX_POINTS = 4096
    try:
        tx.grad = tW.grad = None
        assert W.size >= ops.KPCONV_SPLIT_MIN_WEIGHTS
        out = ops.KPConv_ops(T(qry), T(sup), T(nb), tx, T(Kp), tW, 0.06, influence, agg)
        assert out.grad_fn.name().startswith("_KPConvSplit")
        assert relerr(N(out), ref) <= RTOL
        out.backward(T(dout))
        assert relerr(N(tx.grad), dx) <= RTOL and relerr(N(tW.grad), dW) <= RTOL
    finally:
        ops.KPCONV_SPLIT_MAX_POINTS = 0


@pytest.mark.parametrize("K,cin,cout,n", [(16, 32, 32, 3000), (16, 64, 64, 900), (16, 128, 128, 400), (7, 32, 64, 1500),
                                          (16, 24, 40, 700)])
def test_kpconv_kernel_point_counts(K, cin, cout, n):
    """kernel-point counts other than the usual 15 -- in particular the full 16 the padded layout holds -- on every
    path: exact fp32, tcgen05 forward, training forward that keeps wf + its backward, and the small-point-set path"""
    rng = np.random.default_rng(K * 100 + cin)
    N0, M = 2 * n, 30
    sup = (rng.random((N0, 3)) * 0.5).astype(np.float32)
    qry = sup[rng.permutation(N0)[:n]]
    bz = np.zeros(N0, dtype=np.int64)
    nb = N(ops.ball_query(0.09, M, T(sup), T(qry), mode="partial_dense", batch_x=T(bz), batch_y=T(bz[:n]))[0])
    Kp = (rng.standard_normal((K, 3)) * 0.035).astype(np.float32)
    Kp[0] = 0
    x = rng.standard_normal((N0, cin)).astype(np.float32)
    W = (rng.standard_normal((K, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    dout = rng.standard_normal((n, cout)).astype(np.float32)
    ref = oracle.kpconv_fwd(qry, sup, nb, x, Kp, W, 0.05, "linear", "sum")
    dx, dW = oracle.kpconv_bwd(qry, sup, nb, x, Kp, W, dout, 0.05, "linear", "sum")
    for prec, tol in (("fp32", RTOL), ("tensor", RTOL_TENSOR), (None, RTOL_TENSOR)):
        tx, tW = T(x).requires_grad_(True), T(W).requires_grad_(True)
        out = ops.KPConv_ops(T(qry), T(sup), T(nb), tx, T(Kp), tW, 0.05, "linear", "sum", precision=prec)
        assert relerr(N(out), ref) <= tol, prec
        out.backward(T(dout))
        assert relerr(N(tx.grad), dx) <= RTOL, prec
        assert relerr(N(tW.grad), dW) <= RTOL, prec


def test_kpconv_layer_module_and_linearity():
    """KPConvLayer (kernels.py:20-104) + size-independent properties at the C3 size: linear in the
    features, all-shadow rows give exactly 0, shard-by-cloud == whole batch."""
    pos, batch = synthetic.indoor_spheres()
    tp, tb = T(pos), T(batch)
    layer = kpconv.KPConvLayer(32, 32, point_influence=0.04).to(DEV)
    nb = spatial_ops.RadiusNeighbourFinder(0.1, 34, conv_type="partial_dense")(tp, tp, tb, tb)
    g = torch.Generator(device=DEV).manual_seed(1)
    a = torch.randn(pos.shape[0], 32, device=DEV, generator=g)
    b = torch.randn(pos.shape[0], 32, device=DEV, generator=g)
    ya, yb, yab = layer(tp, tp, nb, a), layer(tp, tp, nb, b), layer(tp, tp, nb, 2 * a - 3 * b)
    assert relerr(N(yab), N(2 * ya - 3 * yb)) <= RTOL_TENSOR
    nb2 = nb.clone()
    nb2[::100] = pos.shape[0]
    assert (N(layer(tp, tp, nb2, a))[::100] == 0).all()
    # shard equivalence: clouds 0-3 alone == rows of the full batch
    cut = int((batch < 4).sum())
    nb_half = spatial_ops.RadiusNeighbourFinder(0.1, 34, conv_type="partial_dense")(tp[:cut], tp[:cut], tb[:cut], tb[:cut])
    nb_full = nb[:cut].clone()
    nb_full[nb_full == pos.shape[0]] = cut
    assert torch.equal(nb_half, nb_full)
    assert torch.equal(layer(tp[:cut], tp[:cut], nb_half, a[:cut]), ya[:cut])
    # oracle on a slice of query rows
    sub = slice(1000, 1256)
    ref = oracle.kpconv_fwd(pos[sub], pos, N(nb)[sub], N(a), N(layer.K_points), N(layer.weight), 0.04)
    assert relerr(N(ya)[sub], ref) <= RTOL_TENSOR


def test_resnet_block_against_reference_composition():
    """ResnetBBlock (modules/KPConv/blocks.py:100-214), strided + bottleneck, eval-mode BN: our block ==
    unary_1 -> KPConv_ops (CPU oracle) -> BN -> LeakyReLU -> unary_2 + shortcut_op(max over neighbour rows)"""
    torch.manual_seed(3)
    pos_h, batch_h = synthetic.indoor_spheres(num_spheres=2, raw_points=5000, radius=0.4, seed=91)
    pos, batch = T(pos_h), T(batch_h)
    block = kpconv.ResnetBBlock([64, 128], grid_size=0.08, prev_grid_size=0.04, max_num_neighbors=30,
                                random_rotation=False).to(DEV).eval()
    x = torch.randn(pos.shape[0], 64, device=DEV)
    with torch.no_grad():
        out = block(kpconv.Cloud(pos, x, batch))
        q_pos, q_batch = kpconv.grid_subsample(pos, batch, 0.08)
        assert torch.equal(out.pos, q_pos) and torch.equal(out.batch, q_batch)
        assert 0 < q_pos.shape[0] < pos.shape[0]
        nb = N(out.idx_neighboors)
        ptr_x, ptr_y = oracle.batch_to_ptr(batch_h), oracle.batch_to_ptr(N(q_batch))
        assert np.array_equal(nb, oracle.ball_query_csr(0.1, 30, pos_h, N(q_pos), ptr_x, ptr_y)[0])
        h = block.unary_1(x)
        layer = block.kp_conv.kp_conv
        conv = oracle.kpconv_fwd(N(q_pos), pos_h, nb, N(h), N(layer.K_points), N(layer.weight), 0.04)
        h2 = block.kp_conv.activation(block.kp_conv.bn(T(conv)))
        main = block.unary_2(h2)
        xs = torch.cat([x, torch.zeros_like(x[:1])])
        want = main + block.shortcut_op(xs[out.idx_neighboors].max(1)[0])
    assert relerr(N(out.x), N(want)) <= 5e-5


def test_kpconv_encoder_stack_forward_backward():
    """the KPConvPaper down-conv stack (conf/models/segmentation/kpconv.yaml:14-57) end to end on the device:
    grid subsampling, partial-dense neighbour search, KPConv (tcgen05 where the shape allows), BN, shortcuts"""
    torch.manual_seed(4)
    pos_h, batch_h = synthetic.indoor_spheres(num_spheres=2, raw_points=24000, radius=0.62, seed=92)
    pos, batch = T(pos_h), T(batch_h)
    enc = kpconv.KPConvEncoder(5, max_nn_cap=34).to(DEV)
    feats = torch.cat([torch.ones(pos.shape[0], 1, device=DEV), torch.randn(pos.shape[0], 4, device=DEV)], 1)
    out = enc(pos, feats, batch)
    assert out.x.shape[1] == enc.out_channels == 2048 and out.x.shape[0] == out.pos.shape[0] < pos.shape[0] // 50
    assert torch.isfinite(out.x).all()
    out.x.square().mean().backward()
    grads = [p.grad for n_, p in enc.named_parameters() if p.requires_grad]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert sum(float(g.abs().sum()) > 0 for g in grads) >= len(grads) - 2


def test_multiscale_transform_precomputed_path_equals_on_the_fly():
    """MultiScaleTransform (core/data_transform/transforms.py:579-651) on the device: the `precomputed=` path of the
    blocks (modules/KPConv/blocks.py:70-79) reproduces the on-the-fly path bit for bit, and the pyramid obeys the
    reference test's consistency checks (test/test_transform.py:67-113: level positions chain through the samplers,
    neighbour tables have the finder's width, shadow index = support count)"""
    torch.manual_seed(6)
    pos_h, batch_h = synthetic.indoor_spheres(num_spheres=2, raw_points=20000, radius=0.6, seed=93)
    pos, batch = T(pos_h), T(batch_h)
    enc = kpconv.KPConvEncoder(5, max_nn_cap=34, random_rotation=False).to(DEV).eval()
    feats = torch.randn(pos.shape[0], 5, device=DEV)
    ms = kpconv.MultiScaleTransform(enc.strategies())(pos, batch)
    assert len(ms) == len(enc.blocks) == 10
    support, s_batch = pos, batch
    for level, block in zip(ms, enc.blocks):
        if block.sampler is None:
            assert torch.equal(level.pos, support)
        else:
            q_pos, q_batch = block.sampler(support, s_batch)
            assert torch.equal(level.pos, q_pos) and torch.equal(level.batch, q_batch)
        nf = block.neighbour_finder
        assert level.idx_neighboors.shape == (level.pos.shape[0], nf._max_num_neighbors)
        assert int(level.idx_neighboors.max()) <= support.shape[0] and int(level.idx_neighboors.min()) >= 0
        support, s_batch = level.pos, level.batch
    with torch.no_grad():
        a = enc(pos, feats, batch)
        b = enc(pos, feats, batch, precomputed=ms)
    assert torch.equal(a.x, b.x) and torch.equal(a.pos, b.pos) and b.block_idx == 10
    # upsample half (transforms.py:616-629): one precompute per strided level, innermost first, and
    # "precomputed == direct" (test/test_interpolateop.py:10-35)
    st = dict(enc.strategies(), upsample_op=[spatial_ops.KNNInterpolate(3) for _ in range(4)])
    tr = kpconv.MultiScaleTransform(st)
    ups = tr.upsample(pos, batch, ms)
    assert len(ups) == 4
    assert [u.num_nodes for u in ups] == [ms[7].pos.shape[0], ms[5].pos.shape[0], ms[3].pos.shape[0], pos.shape[0]]
    query = kpconv.Cloud(ms[2].pos, torch.randn(ms[2].pos.shape[0], 8, device=DEV), ms[2].batch)
    support = kpconv.Cloud(pos, None, batch)
    direct = spatial_ops.KNNInterpolate(3)(query, support)
    pre = spatial_ops.KNNInterpolate(3)(query, support, precomputed=ups[-1])
    assert relerr(N(pre), N(direct)) <= 1e-5
    with pytest.raises(ValueError):
        kpconv.MultiScaleTransform(dict(enc.strategies(), upsample_op=[spatial_ops.KNNInterpolate(1)])).upsample(pos, batch, ms)


@pytest.mark.parametrize("C,n,M", [(70, 300, 20), (128, 300, 38), (1024, 57, 38), (36, 1000, 5)])
def test_neighbor_max(C, n, M):
    rng = np.random.default_rng(81)
    x = rng.standard_normal((500, C)).astype(np.float32)
    nb = rng.integers(0, 501, (n, M))
    tx = T(x).requires_grad_(True)
    out = ops.neighbor_max(tx, T(nb))
    assert np.array_equal(N(out), oracle.neighbor_max(x, nb))
    out.sum().backward()
    x1 = torch.cat([T(x), torch.zeros(1, C, device=DEV)]).requires_grad_(True)
    x1[T(nb)].max(1)[0].sum().backward()
    assert torch.equal(tx.grad, x1.grad[:-1])


# ------------------------------------------------------------------- layer op chains ---------
def test_pointnet_msg_down_layer_c1_chain():
    """C1: FPS 512 of 1024 + ball_query r=0.2 k=32 + grouping, batch 16 -- the whole chain through
    PointNetMSGDown against the oracle chain."""
    xyz = synthetic.surface_clouds(16, 1024, seed=1001)
    feat = np.random.default_rng(1001).standard_normal((16, 64, 1024)).astype(np.float32)
    layer = pointnet2.PointNetMSGDown(npoint=512, radii=[0.2], nsample=[32], down_conv_nn=[[67, 32]]).to(DEV)
    from types import SimpleNamespace as NS
    pos, x = T(xyz), T(feat)
    idx = layer.sampler(pos)
    oidx = oracle.fps_dense(xyz, 512)
    assert np.array_equal(N(idx), oidx)
    new_pos = np.take_along_axis(xyz, oidx[..., None], 1)
    ridx = layer.neighbour_finder(pos, T(new_pos), scale_idx=0)
    oridx = oracle.ball_query_dense(0.2, 32, xyz, new_pos)[0]
    assert np.array_equal(N(ridx), oridx)
    feats = layer._prepare_features(x, pos, T(new_pos), ridx, 0)
    want_f = oracle.group_fwd(feat, oridx)
    want_p = oracle.group_fwd(np.ascontiguousarray(xyz.transpose(0, 2, 1)), oridx) - new_pos.transpose(0, 2, 1)[..., None]
    assert np.array_equal(N(feats)[:, 3:], want_f) and np.array_equal(N(feats)[:, :3], want_p)
    out = layer(NS(x=x, pos=pos))
    assert out.x.shape == (16, 32, 512) and out.pos.shape == (16, 512, 3)
    out.x.sum().backward()
    assert all(p.grad is not None for p in layer.parameters())



def test_compat_tp_namespace_on_the_rsconv_call_sites():
    """`import torch_points3d_b200.compat.tp as tp` serves the reference's call sites unchanged: the RSConv feature
    preparation (modules/RSConv/dense.py:144-163: two `tp.grouping_operation` calls on transposed positions and on
    features, concatenated with the centred positions) and the DENSE sampler / finder / FP calls
    (sampling.py:100, neighbour_finder.py:164, core/base_conv/dense.py:136-140), against the oracle."""
    import torch_points3d_b200.compat.tp as tp

    B, n, S, K, C = 3, 600, 64, 24, 10
    xyz = synthetic.surface_clouds(B, n, seed=91)
    feat = np.random.default_rng(91).standard_normal((B, C, n)).astype(np.float32)
    pos, x = T(xyz), T(feat).requires_grad_(True)
    idx = tp.furthest_point_sample(pos, S)
    oidx = oracle.fps_dense(xyz, S)
    assert np.array_equal(N(idx), oidx)
    new_pos = pos.gather(1, idx.unsqueeze(-1).repeat(1, 1, 3))
    radius_idx = tp.ball_query(0.25, K, pos, new_pos, mode="dense")[0]
    onew = np.take_along_axis(xyz, oidx[..., None], 1)
    oball = oracle.ball_query_dense(0.25, K, xyz, onew)[0]
    assert np.array_equal(N(radius_idx), oball)
    # RSConv._prepare_features, line for line
    new_pos_trans = pos.transpose(1, 2).contiguous()
    grouped_pos_absolute = tp.grouping_operation(new_pos_trans, radius_idx)
    centroids = new_pos.transpose(1, 2).unsqueeze(-1)
    grouped_pos_normalized = grouped_pos_absolute - centroids
    grouped_features = tp.grouping_operation(x, radius_idx)
    new_features = torch.cat([grouped_pos_absolute, grouped_pos_normalized, grouped_features], dim=1)
    want_abs = oracle.group_fwd(np.ascontiguousarray(xyz.transpose(0, 2, 1)), oball)
    want = np.concatenate([want_abs, want_abs - onew.transpose(0, 2, 1)[..., None], oracle.group_fwd(feat, oball)], 1)
    assert np.array_equal(N(new_features), want)
    g = torch.randn_like(new_features)
    new_features.backward(g)
    assert relerr(N(x.grad), oracle.group_bwd(N(g)[:, 6:], oball, n)) <= RTOL
    # DenseFPModule.conv's two calls
    dist, i3 = tp.three_nn(pos, new_pos)
    odist, oi3 = oracle.three_nn(xyz, onew)
    assert np.array_equal(N(i3), oi3)
    recip = 1.0 / (dist + 1e-8)
    weight = recip / torch.sum(recip, dim=2, keepdim=True)
    coarse = np.random.default_rng(92).standard_normal((B, C, S)).astype(np.float32)
    got = tp.three_interpolate(T(coarse), i3, weight)
    assert np.array_equal(N(got), oracle.three_interpolate_fwd(coarse, oi3, N(weight)))


# ---------------------------------------------------------------- tall tcgen05 GEMM / unary layers ---------
@pytest.mark.parametrize("n,Cin,Cout", [(50000, 64, 128), (15000, 128, 32), (777, 4, 32), (500, 1536, 512),
                                        (300, 512, 1024), (100, 64, 13), (5000, 100, 96), (2000, 40, 48), (1, 256, 256),
                                        (129, 16, 16), (4000, 480, 7680 // 16), (0, 32, 32)])
def test_linear_forward_backward(n, Cin, Cout, monkeypatch):
    """ops.linear / ops.Linear (the unary layers of the KPConv blocks, blocks.py:176-194) against float64 matmuls:
    tcgen05 row GEMM for output widths that are multiples of 16 (forward: K-contiguous weights, input gradient:
    column-contiguous weights; partial last row tile, reductions that are not multiples of 64, several column groups),
    library fallback otherwise.  Tolerance 2e-5 of the largest magnitude."""
    rng = np.random.default_rng(n + Cin)
    x = rng.standard_normal((n, Cin)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin)) / np.sqrt(Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    dy = rng.standard_normal((n, Cout)).astype(np.float32)
    monkeypatch.setattr(ops, "LINEAR_MIN_ROWS", 1)   # every size through the tensor-core kernels
    lin = ops.Linear(Cin, Cout, bias=True).to(DEV)
    with torch.no_grad():
        lin.weight.copy_(T(w))
        lin.bias.copy_(T(b))
    assert set(lin.state_dict()) == {"weight", "bias"}
    tx = T(x).requires_grad_(True)
    y = lin(tx)
    assert y.shape == (n, Cout)
    y.backward(T(dy))
    x64, w64, dy64 = x.astype(np.float64), w.astype(np.float64), dy.astype(np.float64)
    if n:
        assert relerr(N(y), x64 @ w64.T + b) <= 2e-5
        assert relerr(N(tx.grad), dy64 @ w64) <= 2e-5
        assert relerr(N(lin.weight.grad), dy64.T @ x64) <= 2e-5
        g1 = lin.weight.grad.clone()
        lin.weight.grad = None
        lin(T(x)).backward(T(dy))
        assert torch.equal(g1, lin.weight.grad)          # the split reduction is deterministic
    else:
        assert float(lin.weight.grad.abs().sum()) == 0.0
