"""Inputs of the whole-network comparison (tests/golden/make_golden_kpconv_net.py on the REFERENCE's own KPConvPaper
forward over its own modules; tests/test_reference_kpconv_net.py on kpconv.KPConvPaper): two slab-shaped clouds, wide
enough that the coarsest of the five levels still holds several points per cloud (the reference's FPModule_PD treats
"as many rows as clouds" as an innermost module), dense enough that the first-level balls hold ~10 neighbours."""
import numpy as np
import torch

from kpconv_blocks_cases import seeded_parameters  # noqa: F401

FEAT, NUM_CLASSES, FIRST_SUBSAMPLING = 4, 13, 0.25


def make_input(seed=17):
    g = torch.Generator().manual_seed(seed)
    sizes = [420, 380]
    pos = torch.rand((sum(sizes), 3), generator=g) * torch.tensor([8.0, 8.0, 0.8])
    batch = torch.repeat_interleave(torch.arange(2), torch.tensor(sizes))
    x = torch.randn((sum(sizes), FEAT), generator=g)
    y = torch.randint(0, NUM_CLASSES, (sum(sizes),), generator=g)
    y[::17] = -1                                            # IGNORE_LABEL rows do not count in the loss
    return pos, x, batch, y


def record(out, tag, log_probs, loss):
    out[tag + "/log_probs"] = log_probs.detach().numpy()
    out[tag + "/loss"] = np.asarray(float(loss))
    out[tag + "/pred"] = torch.max(log_probs, -1)[1].numpy()
