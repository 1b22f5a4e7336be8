"""The MESSAGE_PASSING RSConv modules of torch_points3d_b200.rsconv against the REFERENCE's own
modules/RSConv/message_passing.py.

tests/golden/rsconv_mp_chain.npz holds outputs of the reference's Convolution and RSConvDown (file-path import over
core/base_conv/message_passing.py, tests/golden/make_golden_rsconv_mp.py; torch_geometric's `radius` / `knn` /
`MessagePassing` stood in for).  The same cases run here on the mirrors with `ops.radius`, `ops.edge_encoding` and
`ops.segment_max` replaced by stand-ins of the same definitions: order of the relation vector, positions as features,
targets without edges, channel raising, the chain's idx / edge_index / carried attributes (reference's literal
position-pair order at ratio 1), parameter names.  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import mp_chain_cases  # noqa: E402
import rsconv_mp_cases as cases  # noqa: E402

from torch_points3d_b200 import ops, partial_dense as pd, rsconv  # noqa: E402


def test_rsconv_message_passing_equals_the_reference_modules(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "rsconv_mp_chain.npz"))

    def edge_encoding(pos_src, pos_dst, src, dst):
        pos_i, pos_j = pos_dst[dst], pos_src[src]
        vij = pos_i - pos_j
        return torch.cat([pos_i, pos_j, vij, torch.norm(vij, dim=1).unsqueeze(1)], dim=1)

    monkeypatch.setattr(ops, "radius", cases.radius)
    monkeypatch.setattr(ops, "edge_encoding", edge_encoding)
    monkeypatch.setattr(ops, "segment_max", mp_chain_cases.segment_max)
    ns = SimpleNamespace(Convolution=rsconv.Convolution, RSConvDown=rsconv.RSConvDown,
                         extra={"reference_pos_order": True})
    got = cases.run_cases(ns, pd._Data)
    assert sorted(got.keys()) == sorted(want.files)
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "USbi":
            assert np.array_equal(value, ref), key
        else:
            assert value.shape == ref.shape, key
            assert np.allclose(value, ref, rtol=1e-5, atol=1e-6), (key, float(np.abs(value - ref).max()))
