"""The KPConv block orchestration of torch_points3d_b200.kpconv against the REFERENCE's own blocks.

tests/golden/kpconv_blocks.npz holds outputs of the reference's SimpleBlock / ResnetBBlock / KPDualBlock
(modules/KPConv/blocks.py over kernels.py and the reference's own torch KPConv_ops, run on CPU by
tests/golden/make_golden_kpconv_blocks.py; the partial-dense ball query and the voxel means come from the oracle-backed
stand-ins of tests/golden/kpconv_blocks_cases.py).  The same cases run here on the mirrors with `ops.*` replaced:
ball query and voxel means by the same stand-ins, the KPConv aggregation and the neighbour max by the oracle (itself
pinned to reference-made goldens).  Compared: neighbour tables and strided query points (exactly), block outputs in
eval and training mode (rigid and deformable blocks, with modulations, both loss modes and the regularisers they
leave on the layer), running statistics (2e-5 of max|.|: the oracle's kernel-point sum is ordered differently from
the reference's batched matmul), search and kernel radii, counters, parameter names.  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import kpconv_blocks_cases as cases  # noqa: E402

import oracle  # noqa: E402
from torch_points3d_b200 import kpconv, ops  # noqa: E402


def test_kpconv_blocks_equal_the_reference_blocks(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "kpconv_blocks.npz"))

    def grid_cluster(pos, size, batch=None, want_mean=False, want_coords=False):
        mean, qb = cases.grid_means(pos, batch, size)
        return SimpleNamespace(mean_pos=mean, batch=qb)

    def kpconv_ops(query, support, neighbors, feats, K_points, K_values, extent, influence, aggregation,
                   precision=None):
        out = oracle.kpconv_fwd(query.detach().numpy(), support.detach().numpy(), neighbors.numpy(),
                                feats.detach().numpy(), K_points.detach().numpy(), K_values.detach().numpy(),
                                float(extent), influence, aggregation)
        return torch.from_numpy(out)

    def kpconv_deform_fitting(query, support, neighbors, feats, K_points, offsets, modulations, K_values, extent,
                              influence, aggregation):
        out, sqd, kpd = oracle.kpconv_deform_fwd(
            query.detach().numpy(), support.detach().numpy(), neighbors.numpy(), feats.detach().numpy(),
            K_points.detach().numpy(), offsets.detach().numpy(),
            None if modulations is None else modulations.detach().numpy(), K_values.detach().numpy(), float(extent),
            influence)
        return torch.from_numpy(out), torch.from_numpy(np.asarray(sqd)).min(dim=1)[0], torch.from_numpy(np.asarray(kpd))

    def neighbor_max(x, neighbors):
        return torch.from_numpy(oracle.neighbor_max(x.detach().numpy(), neighbors.numpy()))

    monkeypatch.setattr(ops, "ball_query", cases.ball_query)
    monkeypatch.setattr(ops, "grid_cluster", grid_cluster)
    monkeypatch.setattr(ops, "KPConv_ops", kpconv_ops)
    monkeypatch.setattr(ops, "KPConv_deform_fitting", kpconv_deform_fitting)
    monkeypatch.setattr(ops, "neighbor_max", neighbor_max)
    ns = SimpleNamespace(SimpleBlock=kpconv.SimpleBlock, ResnetBBlock=kpconv.ResnetBBlock,
                         KPDualBlock=kpconv.KPDualBlock)
    with torch.no_grad():
        got = cases.run_cases(ns, kpconv.Cloud)
    assert sorted(got.keys()) == sorted(want.files)
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "USbi":
            assert np.array_equal(value, ref), key          # names, flags, counters, neighbour tables, batch vectors
        elif key.endswith("_loss"):
            assert np.allclose(value, ref, rtol=1e-4, atol=1e-7), (key, value, ref)   # means over thousands of terms
        elif key.endswith("/pos") or key.endswith("radius"):
            assert np.allclose(value, ref, rtol=1e-6, atol=1e-7), key
        else:
            scale = float(np.abs(ref).max())
            assert value.shape == ref.shape, key
            assert float(np.abs(value - ref).max()) <= 2e-5 * scale, (key, float(np.abs(value - ref).max()), scale)
