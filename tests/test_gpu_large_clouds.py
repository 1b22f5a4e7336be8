"""GPU parity of the large-cloud paths the C4 numbers run on (BASELINE.json configs[3]: FPS 4096 of 120 000 points
+ knn k = 16 on a batch of 8 LiDAR-shaped scans): every FPS dispatch variant for clouds that do not fit one CTA
(thread-block clusters of 2 / 4 / 8 / 16 CTAs, the many-clouds variant of the 65 537..131 072 range) against the
oracle's FPS, bit for bit, at the full 4096 rounds; and the whole C4 step -- FPS indices bit-exact, knn indices and
squared distances equal on every one of the 8 x 4096 queries.  Reference semantics: test/test_fps.py:35-42,
SURVEY.md Appendix A.1 / A.6.
"""
import numpy as np
import pytest
import torch

import oracle
from torch_points3d_b200 import _cabi, ops, spatial_ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module", autouse=True)
def _lib_loaded():
    import __graft_entry__
    __graft_entry__.build()
    assert torch.cuda.is_available()
    before = _cabi.launch_count()
    yield
    assert _cabi.launch_count() > before, "no kernel of libp3d_b200.so was launched"


@pytest.fixture(scope="module")
def scans():
    return synthetic.lidar_scans(B=8)


# (clouds, points, samples): which variant fps_dispatch picks (csrc/fps.cu)
@pytest.mark.parametrize("B,n,S", [
    (3, 12000, 500),     # 8 193..16 384 points: 2-CTA cluster
    (3, 20000, 300),     # 16 385..32 768: 4-CTA cluster
    (3, 30000, 300),
    (2, 50000, 400),     # 32 769..65 536: 8-CTA cluster
    (5, 70000, 200),     # 65 537..131 072 with more than 4 clouds: the many-clouds variant
    (2, 131072, 64),     # the upper edge of the on-chip range
    (1, 131073, 40),     # first size of the global-memory fallback (workspace path)
])
def test_fps_cluster_variants_bit_exact(B, n, S):
    xyz = synthetic.surface_clouds(B, n, seed=300 + n % 997)
    got = N(ops.furthest_point_sample(T(xyz), S))
    assert np.array_equal(got, oracle.fps_dense(xyz, S))


def test_fps_pruned_kernel_ties_and_orders():
    """the bounding-box-pruned cluster kernel (32 769..131 072 points): exact ties (integer lattice, duplicated and
    identical points) go to the lowest index; the result does not depend on how compact consecutive points are
    (lattice order = compact blocks, shuffled order = every block spans the cloud, pruning never fires)"""
    g = np.stack(np.meshgrid(np.arange(36), np.arange(36), np.arange(36), indexing="ij"), -1).reshape(1, -1, 3)
    lattice = g.astype(np.float32) * 0.25
    assert np.array_equal(N(ops.furthest_point_sample(T(lattice), 300)), oracle.fps_dense(lattice, 300))
    rng = np.random.default_rng(11)
    shuffled = lattice[:, rng.permutation(lattice.shape[1])]
    assert np.array_equal(N(ops.furthest_point_sample(T(shuffled), 300)), oracle.fps_dense(shuffled, 300))
    same = np.ones((2, 40000, 3), dtype=np.float32)
    assert np.array_equal(N(ops.furthest_point_sample(T(same), 6)), oracle.fps_dense(same, 6))
    dup = synthetic.surface_clouds(1, 50000, seed=77)
    dup[0, 25000:] = dup[0, :25000]          # every point twice: each round is a two-way tie
    assert np.array_equal(N(ops.furthest_point_sample(T(dup), 500)), oracle.fps_dense(dup, 500))
    # CSR form with a random start inside a large cloud and a small cloud beside it
    pos = np.concatenate([synthetic.surface_clouds(1, 70000, seed=78)[0], synthetic.surface_clouds(1, 900, seed=79)[0]])
    batch = np.repeat(np.arange(2), [70000, 900])
    ptr = oracle.batch_to_ptr(batch)
    got = N(ops.fps(T(pos), T(batch), ratio=0.004, random_start=False))
    assert np.array_equal(got, oracle.fps_csr(pos, ptr, oracle.fps_counts(ptr, 0.004)))


def test_fps_c4_shape_all_rounds_bit_exact(scans):
    """B = 8 x 120 000 -> 4096 (the kernel bench_c4 times at N = 1) and B = 1 (one scan per GPU at N = 8)"""
    want = oracle.fps_dense(scans, 4096)
    got, new_xyz = ops.furthest_point_sample(T(scans), 4096, return_xyz=True)
    assert np.array_equal(N(got), want)
    assert np.array_equal(N(new_xyz), np.take_along_axis(scans, want[..., None], 1))
    one = N(ops.furthest_point_sample(T(scans[3:4]), 4096))
    assert np.array_equal(one, want[3:4])
    two = N(ops.furthest_point_sample(T(scans[4:6]), 4096))       # N = 4 GPUs: two scans per rank
    assert np.array_equal(two, want[4:6])


def test_c4_full_step_against_oracle(scans):
    """the whole C4 step as bench_c4 runs it: FPS -> gather the samples -> knn k = 16 of every sample among its scan"""
    B, n, S, k = scans.shape[0], scans.shape[1], 4096, 16
    t_scans = T(scans)
    qi = ops.furthest_point_sample(t_scans, S)
    q = t_scans.gather(1, qi.unsqueeze(-1).expand(-1, -1, 3)).reshape(-1, 3)
    x = t_scans.reshape(-1, 3)
    bx = torch.arange(B, device=DEV).repeat_interleave(n)
    by = torch.arange(B, device=DEV).repeat_interleave(S)
    col, d2 = ops.knn_table(x, q, k, bx, by, return_dist=True)
    want_i = oracle.fps_dense(scans, S)
    assert np.array_equal(N(qi), want_i)
    oq = np.take_along_axis(scans, want_i[..., None], 1).reshape(-1, 3)
    ptr_x, ptr_y = np.arange(B + 1) * n, np.arange(B + 1) * S
    ocol, od2 = oracle.knn_csr(scans.reshape(-1, 3), oq, k, ptr_x, ptr_y)
    col, d2 = N(col), N(d2)
    assert np.array_equal(d2, od2), "knn squared distances differ from the oracle"
    # indices may differ only inside groups of exactly equal distance (north_star); on this data there are none
    diff = col != ocol
    if diff.any():
        rows = np.unique(np.nonzero(diff)[0])
        for r in rows:
            assert sorted(col[r].tolist()) == sorted(ocol[r].tolist()) or len(set(od2[r].tolist())) < k
    # the edge-list form the reference's KNNNeighbourFinder returns (neighbour_finder.py:46-47)
    row_e, col_e = spatial_ops.KNNNeighbourFinder(k)(x, q, bx, by)
    assert np.array_equal(N(col_e), ocol.reshape(-1))
    assert np.array_equal(N(row_e), np.repeat(np.arange(B * S), k))


def test_ball_query_partial_on_lidar_scan_bit_exact(scans):
    """partial-dense ball query at 120 k points per cloud (2 scans): the size where a brute-force sweep costs the most
    and a cell list pays; bit-exact tables and distances, saturated and sparse rows both present"""
    pts = scans[:2]
    n = pts.shape[1]
    x = pts.reshape(-1, 3)
    idx = oracle.fps_dense(pts, 1500)
    y = np.take_along_axis(pts, idx[..., None], 1).reshape(-1, 3)
    bx, by = np.repeat(np.arange(2), n), np.repeat(np.arange(2), 1500)
    for r, M in ((0.5, 16), (2.0, 40)):
        got_i, got_d = ops.ball_query(r, M, T(x), T(y), mode="partial_dense", batch_x=T(bx), batch_y=T(by))
        want_i, want_d = oracle.ball_query_csr(r, M, x, y, np.array([0, n, 2 * n]), np.array([0, 1500, 3000]))
        assert np.array_equal(N(got_i), want_i)
        assert np.array_equal(N(got_d), want_d)
        full = (want_i[:, -1] != x.shape[0]).mean()
        assert 0.0 < full < 1.0 or M == 16
