"""rsconv.RSConv_MSN -- the whole multi-scale RSConv segmentation network the reference benchmarks -- against the
REFERENCE's own network code over its own modules.

tests/golden/rsconv_net.npz was made by tests/golden/make_golden_rsconv_net.py: `set_input` / `forward` of
models/segmentation/rsconv.py:51-124 over four RSConvOriginalMSGDown, two GlobalDenseBaseModule and the DenseFPModules
built from conf/models/segmentation/rsconv.yaml (`RSConv_MSN`) with the reference's resolver, argument extraction and
classes.  Here the mirror (architecture written out in code) runs on the same two-cloud input with the same seeded
parameters and `ops.*` replaced by the oracle-backed stand-ins of the dense-chain test: parameter names, which level
feeds which feature-propagation module, the two global vectors, the head, logits and the cross-entropy with the
ignore label, in eval and training mode.  No device needed."""
import os
import sys

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
sys.path.insert(0, os.path.dirname(__file__))
import dense_chain_cases  # noqa: E402
import rsconv_net_cases as cases  # noqa: E402
from test_reference_dense_chain import _patch  # noqa: E402

from torch_points3d_b200 import rsconv  # noqa: E402


@pytest.mark.parametrize("fused", [False, True])
def test_rsconv_msn_forward_and_loss_equal_the_reference(monkeypatch, fused):
    want = np.load(os.path.join(GOLDEN, "rsconv_net.npz"))
    _patch(monkeypatch, dense_chain_cases.Kernels())
    model = rsconv.RSConv_MSN(cases.FEAT, cases.NUM_CLASSES, fused_grouping=fused)
    assert list(model.state_dict().keys()) == [str(k) for k in want["state_dict_keys"]]
    cases.seeded_parameters(model, 950)
    pos, x, y = cases.make_input()
    for tag, training in (("eval", False), ("train", True)):
        model.train(training)
        with torch.no_grad():
            logits = model(pos, x)
            loss = model.loss(logits, y)
        ref = want[tag + "/logits"]
        assert logits.shape == ref.shape
        scale = float(np.abs(ref).max())
        assert float(np.abs(logits.numpy() - ref).max()) <= 2e-5 * scale, (tag, float(np.abs(logits.numpy() - ref).max()), scale)
        assert abs(float(loss) - float(want[tag + "/loss"])) <= 2e-5 * abs(float(want[tag + "/loss"])), tag
