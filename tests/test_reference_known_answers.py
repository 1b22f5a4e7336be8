"""The reference's own unit tests for this path that need no third-party kernel, restated on the mirrors with the
reference's inputs and expected values: test/test_msdata.py (MultiScaleData.apply, MultiScaleBatch.from_data_list)
and test/test_kpconv.py (the three regularisers of the deformable KPConv).  The other reference tests of the path
(test_fps, test_sampler, test_interpolateop, test_grid_sampling, test_transform's multi-scale case) drive kernels
and are reproduced in tests/test_oracle.py, tests/test_grid_sampling.py and the -m gpu suites.  No device needed."""
import numpy as np
import torch

from torch_points3d_b200 import kpconv, multiscale
from torch_points3d_b200.partial_dense import _Data


def test_msdata_apply():
    """test/test_msdata.py:11-22"""
    x = torch.tensor([1])
    pos = torch.tensor([1])
    d1 = _Data(x=2 * x, pos=2 * pos)
    d2 = _Data(x=3 * x, pos=3 * pos)
    data = multiscale.MultiScaleData(x=x, pos=pos, multiscale=[d1, d2])
    data.apply(lambda x: 2 * x)
    assert data.x[0] == 2 and data.pos[0] == 2
    assert data.multiscale[0].pos[0] == 4 and data.multiscale[0].x[0] == 4
    assert data.multiscale[1].pos[0] == 6 and data.multiscale[1].x[0] == 6


def test_msdata_batch():
    """test/test_msdata.py:24-45"""
    samples = []
    for v in (1, 2):
        x = torch.tensor([v])
        samples.append(multiscale.MultiScaleData(x=x, pos=x, multiscale=[_Data(x=x, pos=x), _Data(x=4 * x, pos=4 * x)]))
    batch = multiscale.MultiScaleBatch.from_data_list(samples)
    assert batch.x.tolist() == [1, 2] and batch.batch.tolist() == [0, 1]
    assert batch.multiscale[0].batch.tolist() == [0, 1] and batch.multiscale[1].batch.tolist() == [0, 1]
    assert batch.multiscale[1].x.tolist() == [4, 8]


def test_kpconv_losses_known_answers():
    """test/test_kpconv.py:15-42"""
    square = torch.from_numpy(np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float64))
    assert kpconv.permissive_loss(square, 1).item() == np.sqrt(2)
    target = torch.from_numpy(np.asarray([[0.5, 0.5, 0]])).repeat([4, 1])
    sq_distances = torch.sum((target - square) ** 2, dim=-1).unsqueeze(0)
    assert kpconv.fitting_loss(sq_distances, 1).item() == 0.5
    loss = kpconv.repulsion_loss(square.unsqueeze(0), 1).item()
    np.testing.assert_almost_equal(loss, 4 * np.sum(np.asarray([0.25, 0.25, 0.0074])), decimal=3)


def test_kernel_points_are_the_reference_disposition():
    """tests/golden/kernel_points.npz (make_golden_kernel_points.py): the table the reference's own read_ply returns
    for kernels/dispositions/k_015_center.ply, and one output of its load_kernels(0.06, 15, 1, 3, "center").  The shipped
    table is that table; a loaded kernel is the table scaled by the radius, rotated and jittered by ~1 % of the radius --
    compared through what a rotation keeps (distance of every point to the centre point and to its neighbours)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kernel_points.npz"))
    table = kpconv.load_kernels(1.0, random_rotation=False)
    assert table.shape == (15, 3) and np.array_equal(table, g["disposition"].astype(np.float32))
    r = float(g["radius"])

    def invariants(k):
        d = np.linalg.norm(k[:, None, :] - k[None, :, :], axis=-1)
        return d / r                                           # row i = distances of point i to all points, in order

    ref = invariants(g["loaded"].astype(np.float64))
    unrotated = invariants(kpconv.load_kernels(r, random_rotation=False).astype(np.float64))
    ours = invariants(kpconv.load_kernels(r, rng=np.random.default_rng(5)).astype(np.float64))
    assert np.abs(ref - unrotated).max() < 0.08 and np.abs(ours - unrotated).max() < 0.08   # 1 % jitter, two points
    assert np.abs(ours - unrotated).max() > 1e-4               # the jitter is there
    assert np.linalg.norm(g["loaded"][0]) / r < 0.05           # the centre point stays at the centre
    assert list(g["loaded_shape"]) == [1, 15, 3]               # the reference returns [num_kernels, K, 3]
