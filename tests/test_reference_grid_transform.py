"""torch_points3d_b200.grid_transform against the REFERENCE's own core/data_transform/grid_transform.py (SURVEY 8f
row 1, host side).

tests/golden/grid_transform.npz holds outputs of the reference's GridSampling3D / group_data / SaveOriginalPosId run by
tests/golden/make_golden_grid.py.  The same cases run here on the mirror with `ops.grid_cluster`, `ops.segment_mean`
and `ops.segment_vote` replaced by CPU stand-ins built from the same operator restatements (grid_cases.py): which
attribute gets the mean, the vote or the representative point, bool / integer attributes, attributes of another length,
quantised coordinates, `last` mode after the seeded shuffle, list inputs, error types and messages.  The device
kernels themselves are compared with the oracle in tests/test_grid_sampling.py (-m gpu).  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import grid_cases as cases  # noqa: E402

from torch_points3d_b200 import grid_transform as gt, ops  # noqa: E402


def _grid_cluster(pos, size, batch=None, want_mean=True, want_coords=False):
    coords = torch.round(pos / size)
    cluster, unique = cases.consecutive_cluster(cases.voxel_key(coords, batch))
    n = unique.shape[0]
    return ops.VoxelClusters(cluster, unique, None, None, n, cases.scatter_mean(pos, cluster) if want_mean else None,
                             None if batch is None else batch[unique], coords[unique].int() if want_coords else None)


def _voxels_from_cluster(cluster, unique_pos_indices=None):
    return ops.VoxelClusters(cluster, unique_pos_indices, None, None, int(cluster.max()) + 1, None, None, None)


def _segment_vote(labels, vc):
    lo = labels.min()
    hist = cases.scatter_add(torch.nn.functional.one_hot(labels.long() - lo), vc.cluster, dim_size=vc.n)
    return hist.argmax(dim=-1) + lo


def test_grid_transform_equals_the_reference(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "grid_transform.npz"))
    monkeypatch.setattr(ops, "grid_cluster", _grid_cluster)
    monkeypatch.setattr(ops, "voxels_from_cluster", _voxels_from_cluster)
    monkeypatch.setattr(ops, "segment_mean", lambda x, vc: cases.scatter_mean(x.float(), vc.cluster, dim_size=vc.n))
    monkeypatch.setattr(ops, "segment_vote", _segment_vote)
    ns = SimpleNamespace(Data=gt.Data, GridSampling3D=gt.GridSampling3D, group_data=gt.group_data,
                         SaveOriginalPosId=gt.SaveOriginalPosId)
    got = cases.run_cases(ns)
    assert sorted(got.keys()) == sorted(want.files)
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "USbi":
            assert value.shape == ref.shape and np.array_equal(value, ref), (key, value, ref)
        else:
            assert value.shape == ref.shape, key
            assert np.allclose(value, ref, rtol=1e-6, atol=1e-7), (key, float(np.abs(value - ref).max()))


def test_voxels_from_cluster_segment_list():
    """the segment list rebuilt from a caller's cluster ids: points by (cluster, point index), first slot per cluster"""
    cluster = torch.tensor([2, 0, 1, 0, 2, 2, 1])
    vc = ops.voxels_from_cluster(cluster)
    assert vc.n == 3 and vc.order.tolist() == [1, 3, 2, 6, 0, 4, 5] and vc.seg_start.tolist() == [0, 2, 4, 7]
    assert vc.order.dtype == torch.int32 and vc.seg_start.dtype == torch.int32
    empty = ops.voxels_from_cluster(torch.zeros(0, dtype=torch.long))
    assert empty.n == 0 and empty.seg_start.tolist() == [0]
