"""kpconv.KPConvPaper -- the whole C5 network -- against the REFERENCE's own network code over its own modules.

tests/golden/kpconv_net.npz was made by tests/golden/make_golden_kpconv_net.py: the reference's `set_input`, `forward`
and `compute_loss` (models/segmentation/kpconv.py:82-153) run over five KPDualBlocks and four FPModule_PD built from
the reference's architecture file with the reference's classes and its own torch KPConv_ops, on the two-cloud input of
tests/golden/kpconv_net_cases.py.  Here the mirror runs on the same input and the same seeded parameters with `ops.*`
replaced by the oracle-backed stand-ins of the block test: the ones column, five levels of strided / unstrided
blocks, the skip stack, four nearest-neighbour upsamplings with skip concatenation, the head, log-probabilities,
predictions and the NLL loss with the ignore label.  No device needed.  Tolerance 5e-5 of max|reference| per tensor
(ten KPConv layers deep; the oracle sums kernel points in a different order than the reference's batched matmul)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import kpconv_blocks_cases as block_cases  # noqa: E402
import kpconv_net_cases as cases  # noqa: E402
import mp_chain_cases  # noqa: E402

import oracle  # noqa: E402
from torch_points3d_b200 import kpconv, multiscale, ops  # noqa: E402


def test_kpconv_paper_forward_and_loss_equal_the_reference(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "kpconv_net.npz"))

    def grid_cluster(pos, size, batch=None, want_mean=False, want_coords=False):
        mean, qb = block_cases.grid_means(pos, batch, size)
        return SimpleNamespace(mean_pos=mean, batch=qb)

    def kpconv_ops(query, support, neighbors, feats, K_points, K_values, extent, influence, aggregation,
                   precision=None):
        return torch.from_numpy(oracle.kpconv_fwd(
            query.detach().numpy(), support.detach().numpy(), neighbors.numpy(), feats.detach().numpy(),
            K_points.detach().numpy(), K_values.detach().numpy(), float(extent), influence, aggregation))

    monkeypatch.setattr(ops, "ball_query", block_cases.ball_query)
    monkeypatch.setattr(ops, "grid_cluster", grid_cluster)
    monkeypatch.setattr(ops, "KPConv_ops", kpconv_ops)
    monkeypatch.setattr(ops, "neighbor_max", lambda x, nb: torch.from_numpy(oracle.neighbor_max(x.detach().numpy(), nb.numpy())))
    monkeypatch.setattr(ops, "knn_interpolate", mp_chain_cases.knn_interpolate)
    monkeypatch.setattr(ops, "knn_edges", mp_chain_cases.knn)

    model = kpconv.KPConvPaper(cases.FEAT, cases.NUM_CLASSES, first_subsampling=cases.FIRST_SUBSAMPLING,
                               random_rotation=False)
    assert list(model.state_dict().keys()) == [str(k) for k in want["state_dict_keys"]]
    cases.seeded_parameters(model, 700)
    pos, x, batch, y = cases.make_input()
    got = {}
    for tag, training in (("eval", False), ("train", True)):
        model.train(training)
        with torch.no_grad():
            log_probs = model(pos, x, batch)
            loss = torch.nn.functional.nll_loss(log_probs, y, ignore_index=-1)      # kpconv.py:150, IGNORE_LABEL
        cases.record(got, tag, log_probs, loss)
    got["train/head_running_mean"] = model.FC_layer[0][1].batch_norm.running_mean.numpy().copy()
    # ---- the network fed with a pyramid made the reference's way: one MultiScaleTransform per sample with the
    # network's own samplers / finders / upsample operators, collated by MultiScaleBatch.from_data_list (the
    # reference side ran its own MultiScaleTransform + multiscale_data.py + the blocks' `precomputed=` branches)
    transform = multiscale.MultiScaleTransform(model.strategies())
    for tag, training in (("eval_precomputed", False), ("train_precomputed", True)):
        samples = [transform(multiscale.MultiScaleData(pos=pos[batch == b], x=x[batch == b], y=y[batch == b]))
                   for b in range(int(batch.max()) + 1)]
        mb = multiscale.MultiScaleBatch.from_data_list(samples)
        assert torch.equal(mb.pos, pos) and torch.equal(mb.batch, batch)
        got["pyramid/levels"] = np.array([int(l.pos.shape[0]) for l in mb.multiscale])
        got["pyramid/last_table"] = mb.multiscale[-1].idx_neighboors.numpy().copy()
        model.train(training)
        with torch.no_grad():
            log_probs = model(mb.pos, mb.x, mb.batch, multiscale=mb.multiscale, upsample=mb.upsample)
            loss = torch.nn.functional.nll_loss(log_probs, mb.y, ignore_index=-1)
        cases.record(got, tag, log_probs, loss)
    assert sorted(got) == sorted(k for k in want.files if k != "state_dict_keys")
    for key, value in got.items():
        ref = want[key]
        assert value.shape == ref.shape, key
        if key.startswith("pyramid/"):
            assert np.array_equal(value, ref), key                                  # level sizes, the collated table
        elif key.endswith("/pred"):
            assert float((value == ref).mean()) >= 0.995, (key, float((value == ref).mean()))   # near-ties may flip
        else:
            scale = max(float(np.abs(ref).max()), 1e-30)
            assert float(np.abs(value - ref).max()) <= 5e-5 * scale, (key, float(np.abs(value - ref).max()), scale)
