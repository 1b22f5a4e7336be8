"""Input and stand-ins of the MESSAGE_PASSING network comparison (tests/golden/make_golden_mp_net.py: the REFERENCE's
own nested U-Net of its `pointnet2` architecture entry; tests/test_reference_mp_net.py: the mirrored SAModule /
GlobalBaseModule / FPModule composed the same way).  One cloud: the reference hands the batch vector to its FPS sampler
in the wrong slot (see message_passing.BaseConvolutionDown), which only a single cloud is indifferent to."""
import numpy as np
import torch

import oracle
from mp_chain_cases import global_max_pool, global_mean_pool, knn, knn_interpolate, radius, segment_max  # noqa: F401
from randla_cases import seeded_parameters  # noqa: F401

FEAT = 3


def fps(pos, batch=None, ratio=0.5, random_start=True):
    """torch_geometric.nn.fps restated on the oracle: ceil(ratio * n) samples per cloud, first sample = first point"""
    b = np.zeros(pos.shape[0], np.int64) if batch is None else batch.numpy()
    ptr = oracle.batch_to_ptr(b)
    return torch.from_numpy(np.asarray(oracle.fps_csr(pos.detach().numpy(), ptr, oracle.fps_counts(ptr, float(ratio))))).long()


def make_input(seed=29):
    g = torch.Generator().manual_seed(seed)
    n = 300
    pos = torch.rand((n, 3), generator=g)
    x = torch.randn((n, FEAT), generator=g)
    return pos, x, torch.zeros(n, dtype=torch.long)
