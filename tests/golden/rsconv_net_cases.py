"""Input of the whole-network comparison of `RSConv_MSN` (tests/golden/make_golden_rsconv_net.py on the REFERENCE's own
RSConvLogicModel forward over its own modules; tests/test_reference_rsconv_net.py on rsconv.RSConv_MSN): two clouds of
1100 points (the first level samples 1024 centres)."""
import numpy as np
import torch

from dense_chain_cases import seeded_parameters  # noqa: F401

FEAT, NUM_CLASSES = 3, 50


def make_input(seed=37):
    g = torch.Generator().manual_seed(seed)
    B, N = 2, 1100
    pos = torch.rand((B, N, 3), generator=g)
    x = torch.randn((B, N, FEAT), generator=g)
    y = torch.randint(0, NUM_CLASSES, (B, N), generator=g)
    y[:, ::23] = -1
    return pos, x, y


def record(out, tag, logits, loss):
    out[tag + "/logits"] = logits.detach().numpy()
    out[tag + "/loss"] = np.asarray(float(loss))
