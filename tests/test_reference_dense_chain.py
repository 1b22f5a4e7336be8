"""The dense PointNet++ host chain of torch_points3d_b200.pointnet2 against the REFERENCE's own modules.

tests/golden/dense_chain.npz holds outputs of the reference's PointNetMSGDown / DenseFPModule / GlobalDenseBaseModule
(modules/pointnet2/dense.py, core/base_conv/dense.py; file-path import with `torch_points_kernels` replaced by the
oracle-backed stand-ins of tests/golden/dense_chain_cases.py: tests/golden/make_golden_dense_chain.py).  Here the same
cases run on the mirrors with `ops.*` replaced by the same stand-ins, once through the op-by-op path and once through
the fused path (`ops.group_features` / `ops.ball_query_multi` / FPS returning the centres, stood in for by their
documented contracts), so the only thing under test is the host chain: call order, centring, normalisation, channel
order, shared MLPs, pooling, interpolation weights, skip concatenation, parameter names.  No device needed; the device
kernels behind `ops.*` are compared with the same oracle in the -m gpu tests.  Tolerance 1e-5 (fp32, same operations;
the fused stand-in divides by the radius where the reference multiplies by nothing else)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import dense_chain_cases  # noqa: E402

from torch_points3d_b200 import ops, pointnet2  # noqa: E402


def _patch(monkeypatch, k):
    def fps(xyz, npoint, return_xyz=False):
        idx = k.furthest_point_sample(xyz, npoint)
        if return_xyz:
            return idx, xyz.gather(1, idx.unsqueeze(-1).expand(-1, -1, 3))
        return idx

    def ball_query(radius, nsample, x, y, mode="dense", batch_x=None, batch_y=None, sort=False, return_dist=True):
        return k.ball_query(radius, nsample, x, y, mode=mode, batch_x=batch_x, batch_y=batch_y)

    def ball_query_multi(radii, nsamples, x, y, strict=True):
        return [k.ball_query(r, n, x, y)[0] for r, n in zip(radii, nsamples)]

    def group_features(x, pos, new_pos, idx, normalize_by=None, use_xyz=True, prepared=None):
        """contract of the fused kernel: cat([(pos[idx] - new_pos) / r, x[idx]], dim=1)"""
        if x is None and not use_xyz:
            raise AssertionError("Cannot have not features and not use xyz as a feature!")
        parts = []
        if use_xyz:
            rel = k.grouping_operation(pos.transpose(1, 2).contiguous(), idx) - new_pos.transpose(1, 2).unsqueeze(-1)
            parts.append(rel if normalize_by is None else rel / float(normalize_by))
        if x is not None:
            parts.append(k.grouping_operation(x, idx))
        return torch.cat(parts, dim=1)

    def three_nn(unknown, known, return_weight=False):
        dist, idx = k.three_nn(unknown, known)
        if not return_weight:
            return dist, idx
        recip = 1.0 / (dist + 1e-8)                     # core/base_conv/dense.py:137-139
        return dist, idx, recip / recip.sum(dim=2, keepdim=True)

    monkeypatch.setattr(ops, "furthest_point_sample", fps)
    monkeypatch.setattr(ops, "ball_query", ball_query)
    monkeypatch.setattr(ops, "ball_query_multi", ball_query_multi)
    monkeypatch.setattr(ops, "grouping_operation", k.grouping_operation)
    monkeypatch.setattr(ops, "group_features", group_features)
    monkeypatch.setattr(ops, "three_nn", three_nn)
    monkeypatch.setattr(ops, "three_interpolate", k.three_interpolate)


@pytest.mark.parametrize("fused", [False, True])
def test_dense_chain_equals_the_reference_modules(monkeypatch, fused):
    want = np.load(os.path.join(GOLDEN, "dense_chain.npz"))
    k = dense_chain_cases.Kernels()
    _patch(monkeypatch, k)
    ns = SimpleNamespace(PointNetMSGDown=pointnet2.PointNetMSGDown, DenseFPModule=pointnet2.DenseFPModule,
                         GlobalDenseBaseModule=pointnet2.GlobalDenseBaseModule, kwargs={"fused_grouping": fused})
    got = dense_chain_cases.run_cases(ns, SimpleNamespace)
    assert sorted(got.keys()) == sorted(key for key in want.files if key != "kernel_calls")
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "US" or value.dtype == bool:
            assert np.array_equal(value, ref), key              # parameter names, flags, error message
        else:
            assert value.shape == ref.shape, key
            assert np.allclose(value, ref, rtol=1e-5, atol=1e-6), (key, float(np.abs(value - ref).max()))
    if not fused:                                              # op by op: the same kernels the same number of times
        calls = sorted("%s=%d" % kv for kv in k.calls.items())
        assert calls == [str(c) for c in want["kernel_calls"]]
