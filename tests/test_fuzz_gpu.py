"""Seeded random-shape sweep of every operator against the oracle.  compute-sanitizer is not available on the GPU
pool, so memory safety rests on shapes nobody would pick by hand: odd sizes around the tile / warp / vector widths,
ragged and empty clouds, duplicate points (exact ties), K larger than the cloud, guard rows around every output."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = 12


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def _clouds(rng, B, n, dup=0.1):
    xyz = rng.random((B, n, 3)).astype(np.float32)
    if n > 1 and dup > 0:  # duplicates: exact distance ties
        src = rng.integers(0, n, max(1, int(dup * n)))
        dst = rng.integers(0, n, src.shape[0])
        xyz[:, dst] = xyz[:, src]
    return xyz


def _ragged(rng, clouds, max_n):
    counts = rng.integers(0, max_n + 1, clouds)
    counts[rng.integers(0, clouds)] = 0  # an empty cloud
    if counts.sum() == 0:
        counts[0] = 3
    batch = np.repeat(np.arange(clouds), counts).astype(np.int64)
    return rng.random((batch.shape[0], 3)).astype(np.float32), batch


@pytest.mark.parametrize("seed", range(CASES))
def test_fuzz_dense_chain(seed):
    from torch_points3d_b200 import ops
    rng = np.random.default_rng(1000 + seed)
    B, n = int(rng.integers(1, 6)), int(rng.choice([1, 2, 31, 33, 127, 257, 1000, 2049, 4100]))
    S = int(rng.integers(1, n + 1))
    K = int(rng.choice([1, 3, 16, 32, 50, 129]))
    r = float(rng.choice([0.05, 0.2, 0.7, 5.0]))
    xyz = _clouds(rng, B, n)
    pos = T(xyz)
    idx = ops.furthest_point_sample(pos, S)
    assert np.array_equal(N(idx), oracle.fps_dense(xyz, S))
    ctr = np.take_along_axis(xyz, N(idx)[..., None], 1)
    bidx, bd2 = ops.ball_query(r, K, pos, T(ctr))
    oidx, od2 = oracle.ball_query_dense(r, K, xyz, ctr)
    assert np.array_equal(N(bidx), oidx) and np.array_equal(N(bd2), od2)
    radii, ks = [r, r * 0.5, r * 2][: int(rng.integers(1, 4))], [K, max(1, K // 2), K + 1]
    for rr, kk, got in zip(radii, ks, ops.ball_query_multi(radii, ks[: len(radii)], pos, T(ctr))):
        assert np.array_equal(N(got), oracle.ball_query_dense(rr, kk, xyz, ctr)[0])
    C = int(rng.choice([1, 3, 8, 17, 64]))
    feat = rng.standard_normal((B, C, n)).astype(np.float32)
    tf = T(feat).requires_grad_(True)
    fused = ops.group_features(tf, pos, T(ctr), bidx)
    assert np.array_equal(N(fused)[:, 3:], oracle.group_fwd(feat, oidx))
    g = rng.standard_normal(tuple(fused.shape)).astype(np.float32)
    fused.backward(T(g))
    want = oracle.group_bwd(g[:, 3:], oidx, n)
    assert np.abs(N(tf.grad) - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)
    if n >= 3:
        dist, i3 = ops.three_nn(T(ctr), pos)
        od, oi = oracle.three_nn(ctr, xyz)
        assert np.array_equal(N(i3), oi) and np.array_equal(N(dist), od)
        w = rng.random(oi.shape).astype(np.float32)
        tf2 = T(feat).requires_grad_(True)
        out = ops.three_interpolate(tf2, i3, T(w))
        assert np.array_equal(N(out), oracle.three_interpolate_fwd(feat, oi, w))
        go = rng.standard_normal(tuple(out.shape)).astype(np.float32)
        out.backward(T(go))
        wb = oracle.three_interpolate_bwd(go, oi, w, n)
        assert np.abs(N(tf2.grad) - wb).max() <= 1e-5 * max(np.abs(wb).max(), 1e-30)


@pytest.mark.parametrize("seed", range(CASES))
def test_fuzz_batched_searches_and_clustering(seed):
    from torch_points3d_b200 import ops
    rng = np.random.default_rng(2000 + seed)
    clouds = int(rng.integers(1, 7))
    x, bx = _ragged(rng, clouds, int(rng.choice([5, 40, 300, 2500])))
    y, by = _ragged(rng, clouds, int(rng.choice([3, 20, 200])))
    px, py = oracle.batch_to_ptr(bx, clouds), oracle.batch_to_ptr(by, clouds)
    M, r = int(rng.choice([1, 7, 34, 65])), float(rng.choice([0.05, 0.3, 2.0]))
    got = ops.ball_query(r, M, T(x), T(y), mode="partial_dense", batch_x=T(bx), batch_y=T(by))
    want = oracle.ball_query_csr(r, M, x, y, px, py)
    assert np.array_equal(N(got[0]), want[0]) and np.array_equal(N(got[1]), want[1])
    k = int(rng.choice([1, 3, 16, 33]))
    col, d2 = ops.knn_table(T(x), T(y), k, T(bx), T(by))
    wc, wd = oracle.knn_csr(x, y, k, px, py)
    assert np.array_equal(N(d2), wd)
    assert np.array_equal(N(col), wc) or np.array_equal(np.sort(N(d2), 1), np.sort(wd, 1))  # exact-distance ties
    size = float(rng.choice([0.03, 0.2, 1.5]))
    vc = ops.grid_cluster(T(x), size, T(bx), want_mean=True, want_coords=True)
    cl, perm, coords = oracle.grid_cluster(x, size, bx)
    assert vc.n == len(perm) and np.array_equal(N(vc.cluster), cl) and np.array_equal(N(vc.unique_pos_indices), perm)
    assert np.array_equal(N(vc.mean_pos), oracle.scatter_mean(x, cl, vc.n))
    feats = rng.standard_normal((x.shape[0], int(rng.integers(1, 9)))).astype(np.float32)
    assert np.array_equal(N(ops.segment_mean(T(feats), vc)), oracle.scatter_mean(feats, cl, vc.n))


@pytest.mark.parametrize("seed", range(CASES))
def test_fuzz_kpconv(seed):
    from torch_points3d_b200 import kpconv, ops
    rng = np.random.default_rng(3000 + seed)
    n, N0 = int(rng.choice([1, 33, 200, 1500])), int(rng.choice([1, 50, 400, 2000]))
    M = int(rng.choice([1, 8, 34, 40, 61]))
    cin, cout = int(rng.choice([1, 5, 32, 64, 96])), int(rng.choice([1, 7, 32, 64]))
    infl, agg = str(rng.choice(["linear", "gaussian", "constant"])), str(rng.choice(["sum", "closest"]))
    sup = (rng.random((N0, 3)) * 0.3).astype(np.float32)
    qry = (rng.random((n, 3)) * 0.3).astype(np.float32)
    nb = rng.integers(-1, N0 + 1, (n, M)).astype(np.int64)  # -1 and N0 are both shadow neighbours
    Kp = kpconv.load_kernels(0.09, random_rotation=False).astype(np.float32)
    x = rng.standard_normal((N0, cin)).astype(np.float32)
    W = (rng.standard_normal((15, cin, cout)) * 0.1).astype(np.float32)
    dout = rng.standard_normal((n, cout)).astype(np.float32)
    ref = oracle.kpconv_fwd(qry, sup, nb, x, Kp, W, 0.06, infl, agg)
    dx, dW = oracle.kpconv_bwd(qry, sup, nb, x, Kp, W, dout, 0.06, infl, agg)
    scale = max(np.abs(ref).max(), 1e-6)
    for precision, tol in (("fp32", 1e-5), (None, 2e-5)):
        tx, tW = T(x).requires_grad_(True), T(W).requires_grad_(True)
        out = ops.KPConv_ops(T(qry), T(sup), T(nb), tx, T(Kp), tW, 0.06, infl, agg, precision=precision)
        assert np.abs(N(out) - ref).max() <= tol * scale, (precision, infl, agg)
        out.backward(T(dout))
        assert np.abs(N(tx.grad) - dx).max() <= 2e-5 * max(np.abs(dx).max(), 1e-6)
        assert np.abs(N(tW.grad) - dW).max() <= 2e-5 * max(np.abs(dW).max(), 1e-6)
    if agg == "sum":
        off = (rng.standard_normal((n, 15, 3)) * 0.01).astype(np.float32)
        got = ops.KPConv_deform_ops(T(qry), T(sup), T(nb), T(x), T(Kp), T(off), None, T(W), 0.06, infl, "sum")
        want = oracle.kpconv_deform_fwd(qry, sup, nb, x, Kp, off, None, W, 0.06, infl)
        assert np.abs(N(got[0]) - want[0]).max() <= 2e-5 * max(np.abs(want[0]).max(), 1e-6)
        assert np.array_equal(N(got[1]), want[1])
