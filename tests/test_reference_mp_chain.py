"""The MESSAGE_PASSING PointNet++ chain of torch_points3d_b200.message_passing against the REFERENCE's own modules.

tests/golden/mp_chain.npz holds outputs of the reference's SAModule (modules/pointnet2/message_passing.py:9-37 over
BaseMSConvolutionDown.forward, core/base_conv/message_passing.py:60-93), GlobalBaseModule (:129-150) and FPModule
(:156-176 over BaseConvolutionUp.forward :96-127), run by tests/golden/make_golden_mp_chain.py with torch_geometric's
functions stood in for (see that file).  The same cases run here on the mirrors with `ops.*` replaced by stand-ins of
the same definitions.  Multi-scale loop, edge_index = [support, query], concatenation of the scales, carried
attributes, pooled outputs with / without positions, interpolation + skip concatenation (the reference's FPModule does
not forward its keywords, so `skip=False` has no effect there: reproduced), parameter names.  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import mp_chain_cases as cases  # noqa: E402

from torch_points3d_b200 import message_passing as mp, ops, partial_dense as pd  # noqa: E402


def test_mp_chain_equals_the_reference_modules(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "mp_chain.npz"))

    def edge_message(x, pos_src, pos_dst, src, dst):
        rel = pos_src[src] - pos_dst[dst]
        return rel if x is None else torch.cat([x[src], rel], dim=1)

    def segment_sum(msg, dst, T):
        return torch.zeros((T, msg.shape[1]), dtype=msg.dtype).index_add_(0, dst, msg)

    monkeypatch.setattr(ops, "radius", cases.radius)
    monkeypatch.setattr(ops, "knn_interpolate", cases.knn_interpolate)
    monkeypatch.setattr(ops, "edge_message", edge_message)
    monkeypatch.setattr(ops, "segment_max", cases.segment_max)
    monkeypatch.setattr(ops, "segment_sum", segment_sum)
    ns = SimpleNamespace(SAModule=mp.SAModule, GlobalBaseModule=mp.GlobalBaseModule, FPModule=mp.FPModule)
    got = cases.run_cases(ns, pd._Data)
    assert sorted(got.keys()) == sorted(want.files)
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "USbi":
            assert np.array_equal(value, ref), key
        else:
            assert value.shape == ref.shape, key
            assert np.allclose(value, ref, rtol=1e-5, atol=1e-6), (key, float(np.abs(value - ref).max()))
