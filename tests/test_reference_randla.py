"""The RandLA-Net chain of torch_points3d_b200.message_passing against the REFERENCE's own modules.

tests/golden/randla_chain.npz holds outputs of the reference's RandlaKernel / RandlaConv / RandLANetRes
(modules/RandLANet/modules.py:9-127 over core/base_conv/message_passing.py:35-58,212-257; file-path import with
torch_geometric's `knn` and `MessagePassing` replaced by the stand-ins described in tests/golden/make_golden_randla.py).
The same cases run here on the mirrors with `ops.*` replaced by torch stand-ins of the same definitions: the message
arithmetic (relative position encoding, attention softmax, per-target sum, global MLP), the chain's index bookkeeping
(idx, edge_index = [support, query], carried attributes, k = 16 in the residual block), the first-layer size rewrite
and the parameter names must agree.  The reference's literal position-pair order is used on both sides
(`reference_pos_order=True`, ratio-1 samplers: see randla_cases.py).  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import randla_cases as cases  # noqa: E402

from torch_points3d_b200 import message_passing as mp, ops, partial_dense as pd  # noqa: E402


def test_randla_chain_equals_the_reference_modules(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "randla_chain.npz"))

    def edge_encoding(pos_src, pos_dst, src, dst):
        pos_i, pos_j = pos_dst[dst], pos_src[src]
        vij = pos_i - pos_j
        return torch.cat([pos_i, pos_j, vij, torch.norm(vij, dim=1).unsqueeze(1)], dim=1)   # modules.py:36-40

    def segment_sum(msg, dst, T):
        return torch.zeros((T, msg.shape[1]), dtype=msg.dtype).index_add_(0, dst, msg)

    monkeypatch.setattr(ops, "knn_edges", cases.knn)
    monkeypatch.setattr(ops, "edge_encoding", edge_encoding)
    monkeypatch.setattr(ops, "segment_sum", segment_sum)
    ns = SimpleNamespace(RandlaKernel=mp.RandlaKernel, RandlaConv=mp.RandlaConv, RandLANetRes=mp.RandLANetRes,
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
