"""pointnet2.PointNet2_D -- the whole `pointnet2_charlesmsg` network of BASELINE config 2 -- against the REFERENCE's own
network code over its own nested modules.

tests/golden/pointnet2_net.npz was made by tests/golden/make_golden_pointnet2_net.py: the reference's `set_input` /
`forward` (models/segmentation/pointnet2.py:66-113) over the U-Net its own `_init_from_compact_format` /
`UnetSkipConnectionBlock` (models/base_architectures/unet.py) nest from its architecture file with its own
PointNetMSGDown / GlobalDenseBaseModule / DenseFPModule.  Here the mirror (flat module lists, skips wired by hand)
runs on the same input with the same seeded parameters and `ops.*` replaced by the oracle-backed stand-ins of the
dense-chain test: which level's output each feature-propagation module receives as its skip, the global module, the
head, logits and the cross-entropy with the ignore label.  No device needed."""
import os
import sys

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
sys.path.insert(0, os.path.dirname(__file__))
import dense_chain_cases  # noqa: E402
import pointnet2_net_cases as cases  # noqa: E402
from test_reference_dense_chain import _patch  # noqa: E402

from torch_points3d_b200 import pointnet2  # noqa: E402


@pytest.mark.parametrize("fused", [False, True])
def test_pointnet2_d_forward_and_loss_equal_the_reference(monkeypatch, fused):
    want = np.load(os.path.join(GOLDEN, "pointnet2_net.npz"))
    _patch(monkeypatch, dense_chain_cases.Kernels())
    model = pointnet2.PointNet2_D(feat_dim=cases.FEAT, num_classes=cases.NUM_CLASSES)
    for m in model.down:
        m.fused_grouping = fused
    # the reference nests its modules (model.down, model.submodule.down, ...): other names, same order and shapes
    assert ["%s" % (tuple(v.shape),) for v in model.state_dict().values()] == [str(s) for s in want["state_dict_shapes"]]
    cases.seeded_parameters(model, 800)
    pos, x, y = cases.make_input()
    for tag, training in (("eval", False), ("train_mode_no_dropout", True)):
        model.train(training)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.eval()
        with torch.no_grad():
            logits = model(pos, x)
            loss = model.loss(logits, y)
        ref = want[tag + "/logits"]
        assert logits.shape == ref.shape
        scale = float(np.abs(ref).max())
        assert float(np.abs(logits.numpy() - ref).max()) <= 2e-5 * scale, (tag, float(np.abs(logits.numpy() - ref).max()), scale)
        assert abs(float(loss) - float(want[tag + "/loss"])) <= 2e-5 * abs(float(want[tag + "/loss"])), tag
