"""CPU tests: the oracle against every vector the reference's own tests hold for this
path, plus the KPConv golden vectors produced by running the
reference's KPConv_ops (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

import oracle


# ---- reference test/test_fps.py:35-42 : known answer [0, 3, 4] --------------------------
def test_fps_reference_known_answer():
    pos = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.4, 0.2, 0], [2, 2, 2], [-1, -2, -0.01]], dtype=np.float32)
    ptr = np.array([0, 5], dtype=np.int64)
    counts = oracle.fps_counts(ptr, 3.0 / 5.0)
    assert counts.tolist() == [3]
    idx = oracle.fps_csr(pos, ptr, counts)
    assert idx.tolist() == [0, 3, 4]
    assert oracle.fps_dense(pos[None], 3).tolist() == [[0, 3, 4]]


# ---- reference test/test_sampler.py:16-23 : FPSSampler counts ---------------------------
def test_fps_count_rule():
    ptr = np.array([0, 10], dtype=np.int64)
    assert oracle.fps_counts(ptr, 0.5).tolist() == [5]
    assert oracle.fps_counts(ptr, 5 / float(10)).tolist() == [5]
    assert oracle.fps_counts(np.array([0, 3, 10]), 0.5).tolist() == [2, 4]  # ceil per cloud


# ---- reference test/test_sampler.py:36-86 : radius edge counts on a square -------------
def _square():
    x = np.array([[-1, -1, 0], [-1, 1, 0], [1, -1, 0], [1, 1, 0]], dtype=np.float32)
    y = np.array([[-1, 0, 0], [1, 0, 0]], dtype=np.float32)
    return x, y, np.array([0, 4]), np.array([0, 2])


@pytest.mark.parametrize("r,maxn,edges", [(1.1, 4, 4), (10, 4, 8), (10, 3, 6), (1.1, 3, 4)])
def test_radius_reference_edge_counts(r, maxn, edges):
    x, y, px, py = _square()
    e = oracle.radius_csr(x, y, r, px, py, max_num_neighbors=maxn)
    assert e.shape == (2, edges)
    assert (np.diff(e[0]) >= 0).all()


# ---- reference test/test_transform.py:67-113 : 10x10 grid, widths, shadow = N ----------
def test_ball_query_partial_reference_grid():
    N = 10
    xv, yv = np.meshgrid(np.linspace(0, 1, N), np.linspace(0, 1, N))
    pos = np.stack([xv.flatten(), yv.flatten(), np.zeros(N * N)], 1).astype(np.float32)
    ptr = np.array([0, N * N])
    for r, width in [(0.5, 100), (0.5, 150), (1.0, 200)]:
        idx, d2 = oracle.ball_query_csr(r, width, pos, pos, ptr, ptr)
        assert idx.shape == (N * N, width)
        d = ((pos[:, None, :].astype(np.float64) - pos[None].astype(np.float64)) ** 2).sum(-1)
        for i in range(N * N):
            want = set(np.nonzero(d[i] <= r * r + 1e-9)[0].tolist())
            got = set(idx[i].tolist())
            assert got - {N * N} == want  # set equality (test_transform.py:102-103)
            if len(want) < width:
                assert N * N in got  # padded with the shadow index x.shape[0]
        assert ((idx == N * N) == (d2 < 0)).all()


# ---- reference test/test_interpolateop.py:10-23 : 1-NN mapping --------------------------
def test_knn_reference_precompute_mapping():
    pos = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 0], [0.1, 0, 0]], dtype=np.float32)
    # GridSampling3D(1) merges points 3 and 4 (same voxel) -> query = 4 cell means
    query = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [0.05, 0, 0]], dtype=np.float32)
    e = oracle.knn_edges(query, pos, 1, np.array([0, 4]), np.array([0, 5]))
    assert e.shape == (2, 5)
    assert e[0].tolist() == [0, 1, 2, 3, 4]
    assert e[1][4] == e[1][3]  # up.x_idx[4] == up.x_idx[3]


def test_knn_sorted_ties_and_short_clouds():
    x = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [5, 5, 5]], dtype=np.float32)
    y = np.array([[0, 0, 0], [5, 5, 4]], dtype=np.float32)
    col, d2 = oracle.knn_csr(x, y, 3, np.array([0, 3, 4]), np.array([0, 1, 2]))
    assert col[0].tolist() == [0, 1, 2]  # tie 1 vs 2 -> lower index first
    assert col[1].tolist() == [3, -1, -1]
    assert d2[1].tolist() == [1.0, -1.0, -1.0]


# ---- dense ball query: semantic spec A.2 (parity unpinned upstream) --------------------
def test_ball_query_dense_spec():
    x = np.array([[[0, 0, 0], [0.1, 0, 0], [0.5, 0, 0], [0.05, 0, 0], [0.2, 0, 0]]], dtype=np.float32)
    y = np.array([[[0, 0, 0], [10, 10, 10], [0.5, 0, 0]]], dtype=np.float32)
    idx, d2 = oracle.ball_query_dense(0.2, 4, x, y)
    assert idx[0, 0].tolist() == [0, 1, 3, 0]       # 0.2 is NOT < 0.2*0.2 fp32 -> strict; pad with first
    assert idx[0, 1].tolist() == [0, 0, 0, 0]       # empty ball -> zeros
    assert idx[0, 2].tolist() == [2, 2, 2, 2]
    assert d2[0, 1].tolist() == [-1, -1, -1, -1]
    assert d2[0, 0, 3] == -1
    idx2, _ = oracle.ball_query_dense(0.2, 2, x, y)
    assert idx2[0, 0].tolist() == [0, 1]            # truncation keeps index order


def test_boundary_strictness_switch():
    x = np.array([[[0.25, 0, 0]]], dtype=np.float32)
    y = np.zeros((1, 1, 3), dtype=np.float32)
    r = 0.25  # d2 == r2 exactly in fp32
    assert oracle.ball_query_dense(r, 1, x, y, strict=True)[1][0, 0, 0] == -1
    assert oracle.ball_query_dense(r, 1, x, y, strict=False)[1][0, 0, 0] == np.float32(0.0625)
    ptr = np.array([0, 1])
    assert oracle.ball_query_csr(r, 1, x[0], y[0], ptr, ptr)[0][0, 0] == 0            # non-strict default
    assert oracle.ball_query_csr(r, 1, x[0], y[0], ptr, ptr, strict=True)[0][0, 0] == 1  # shadow


# ---- grouping / three_nn / three_interpolate vs straightforward numpy -------------------
def test_grouping_and_interpolate_against_numpy():
    rng = np.random.default_rng(0)
    B, C, N, S, K = 2, 5, 17, 6, 4
    feat = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, S, K))
    out = oracle.group_fwd(feat, idx)
    ref = np.stack([feat[b][:, idx[b]] for b in range(B)])
    assert np.array_equal(out, ref)
    dout = rng.standard_normal(out.shape).astype(np.float32)
    d = oracle.group_bwd(dout, idx, N)
    ref = np.zeros((B, C, N))
    for b in range(B):
        for s in range(S):
            for k in range(K):
                ref[b, :, idx[b, s, k]] += dout[b, :, s, k]
    np.testing.assert_allclose(d, ref, rtol=1e-6, atol=1e-6)

    unknown = rng.standard_normal((B, 9, 3)).astype(np.float32)
    known = rng.standard_normal((B, 7, 3)).astype(np.float32)
    dist, i3 = oracle.three_nn(unknown, known)
    dd = np.sqrt(((unknown[:, :, None] - known[:, None]) ** 2).sum(-1))
    np.testing.assert_array_equal(i3, np.argsort(dd, -1, kind="stable")[..., :3])
    np.testing.assert_allclose(dist, np.sort(dd, -1)[..., :3], rtol=1e-6)
    w = rng.random((B, 9, 3)).astype(np.float32)
    f = rng.standard_normal((B, C, 7)).astype(np.float32)
    o = oracle.three_interpolate_fwd(f, i3, w)
    ref = np.stack([(f[b][:, i3[b]] * w[b][None]).sum(-1) for b in range(B)])
    np.testing.assert_allclose(o, ref, rtol=1e-5, atol=1e-6)
    g = rng.standard_normal(o.shape).astype(np.float32)
    df = oracle.three_interpolate_bwd(g, i3, w, 7)
    ref = np.zeros((B, C, 7))
    for b in range(B):
        for j in range(9):
            for t in range(3):
                ref[b, :, i3[b, j, t]] += g[b, :, j] * w[b, j, t]
    np.testing.assert_allclose(df, ref, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        oracle.three_nn(unknown, known[:, :2])


# ---- KPConv: oracle vs the reference's own KPConv_ops (golden .npz) --------------------
def _golden_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "kpconv_*.npz")))


def test_kpconv_golden_present(golden_dir):
    assert len(_golden_cases(golden_dir)) >= 5


@pytest.mark.parametrize("name", ["kpconv_linear_sum", "kpconv_gaussian_sum", "kpconv_constant_sum",
                                  "kpconv_linear_closest", "kpconv_linear_sum_c32"])
def test_kpconv_oracle_matches_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    args = (z["query"], z["support"], z["neighbors"], z["x"], z["K_points"], z["W"])
    kw = dict(extent=float(z["extent"]), influence=str(z["influence"]), aggregation=str(z["aggregation"]))
    out = oracle.kpconv_fwd(*args, **kw)
    scale = np.abs(z["out"]).max()
    assert np.abs(out - z["out"]).max() <= 1e-5 * scale
    if kw["influence"] != "constant":
        assert np.all(out[0] == 0)  # all-shadow row -> exactly zero
    dx, dW = oracle.kpconv_bwd(*args, z["dout"], **kw)
    assert np.abs(dx - z["dx"]).max() <= 1e-5 * np.abs(z["dx"]).max()
    assert np.abs(dW - z["dW"]).max() <= 1e-5 * np.abs(z["dW"]).max()


def test_fps_properties():
    rng = np.random.default_rng(3)
    xyz = rng.standard_normal((3, 257, 3)).astype(np.float32)
    idx = oracle.fps_dense(xyz, 64)
    assert (idx[:, 0] == 0).all()
    for b in range(3):
        assert len(set(idx[b].tolist())) == 64
    # duplicated points: tie -> lowest index
    dup = np.zeros((1, 6, 3), dtype=np.float32)
    dup[0, 2:4] = [1, 0, 0]
    dup[0, 4:6] = [-1, 0, 0]
    assert oracle.fps_dense(dup, 3).tolist() == [[0, 2, 4]]
    # S == N selects everything once distinct points remain, S == 1 -> [0]
    assert oracle.fps_dense(xyz[:, :5], 1).tolist() == [[0]] * 3
    assert sorted(oracle.fps_dense(xyz[:1, :9], 9)[0].tolist()) == list(range(9))
    with pytest.raises(ValueError):
        oracle.fps_dense(xyz[:, :4], 5)


# ------------------------------------------------------------ deformable KPConv ----
@pytest.mark.parametrize("name", ["kpdeform_linear_mod", "kpdeform_linear", "kpdeform_gaussian_mod", "kpdeform_constant"])
def test_kpconv_deform_oracle_against_reference_golden(name):
    """forward and all four gradients of the numpy restatement against the reference's own KPConv_deform_ops +
    torch autograd (tests/golden/make_golden.py); the loss there feeds all three outputs, including the
    fitting-loss min over neighbours"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    mods = g["modulations"] if "modulations" in g.files else None
    infl, ext = str(g["influence"]), float(g["extent"])
    args = (g["query"], g["support"], g["neighbors"], g["x"], g["K_points"], g["offsets"], mods, g["W"], ext, infl)
    out, sqd, kpd = oracle.kpconv_deform_fwd(*args)
    assert np.abs(out - g["out"]).max() <= 1e-5 * np.abs(g["out"]).max()
    assert np.array_equal(sqd, g["sqd"]) and np.array_equal(kpd, g["deformed"])
    n, M, K = sqd.shape
    kpmin = sqd.min(1)
    assert abs(float(np.mean(kpmin / np.float32((1.5 * ext) ** 2))) - float(g["fitting"])) <= 1e-5 * abs(float(g["fitting"]))
    # cotangent of sq_distances = the golden's dense one + the fitting term through the first argmin
    dsq = g["dsq"].copy()
    am = sqd.argmin(1)
    np.add.at(dsq, (np.arange(n)[:, None], am, np.arange(K)[None, :]), np.float32(3.0 / ((1.5 * ext) ** 2) / (n * K)))
    dx, dW, doff, dm = oracle.kpconv_deform_bwd(*args, g["dout"], dsq, g["ddef"])
    for got, key in ((dx, "dx"), (dW, "dW"), (doff, "doffsets")):
        assert np.abs(got - g[key]).max() <= 2e-5 * np.abs(g[key]).max(), key
    if mods is not None:
        assert np.abs(dm - g["dmodulations"]).max() <= 2e-5 * np.abs(g["dmodulations"]).max()


def test_kpconv_regularisers_against_reference_golden():
    """fitting / repulsion / permissive losses of the deformable layer (mirrors in torch_points3d_b200/kpconv.py)
    against values and gradients produced by the reference's modules/KPConv/losses.py (make_golden.py)"""
    import torch
    from torch_points3d_b200 import kpconv
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kpconv_losses.npz"))
    deformed = torch.from_numpy(g["deformed"]).requires_grad_(True)
    sqd = torch.from_numpy(g["sqd"]).requires_grad_(True)
    for name, fn, arg, rad in (("fitting", kpconv.fitting_loss, sqd, 0.06), ("repulsion", kpconv.repulsion_loss, deformed, 0.04),
                               ("permissive", kpconv.permissive_loss, deformed, 1.2)):
        val = fn(arg, rad)
        assert abs(float(val.detach()) - float(g[name])) <= 1e-6 * abs(float(g[name])), name
        (grad,) = torch.autograd.grad(val, arg)
        assert np.abs(grad.numpy() - g["d" + name]).max() <= 1e-6 * np.abs(g["d" + name]).max(), name
