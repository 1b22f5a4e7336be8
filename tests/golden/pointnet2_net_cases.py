"""Inputs of the whole-network comparison of `pointnet2_charlesmsg` (tests/golden/make_golden_pointnet2_net.py on the
REFERENCE's own PointNet2_D forward over its own nested modules; tests/test_reference_pointnet2_net.py on
pointnet2.PointNet2_D): two clouds of 640 points (the architecture samples 512 then 128 of them)."""
import numpy as np
import torch

from dense_chain_cases import seeded_parameters  # noqa: F401

FEAT, NUM_CLASSES = 3, 50


def make_input(seed=23):
    g = torch.Generator().manual_seed(seed)
    B, N = 2, 640
    pos = torch.rand((B, N, 3), generator=g)
    x = torch.randn((B, N, FEAT), generator=g)
    y = torch.randint(0, NUM_CLASSES, (B, N), generator=g)
    y[:, ::19] = -1                                          # IGNORE_LABEL rows do not count in the loss
    return pos, x, y


def record(out, tag, logits, loss):
    out[tag + "/logits"] = logits.detach().numpy()
    out[tag + "/loss"] = np.asarray(float(loss))
