"""Case list shared by tests/golden/make_golden_rsconv_mp.py (the REFERENCE's own modules/RSConv/message_passing.py,
loaded by file path) and tests/test_reference_rsconv_mp.py (torch_points3d_b200.rsconv): the relation-shape
convolution on a bipartite graph (with features, with positions as features, without the channel-raising MLP) and the
RSConvDown chain at ratio 1 in the reference's literal position-pair order (the order breaks for ratio < 1, see
message_passing.BaseConvolutionDown)."""
import numpy as np
import torch

from mp_chain_cases import radius  # noqa: F401  (oracle-backed torch_cluster.radius)
from randla_cases import FixedSampler, knn, seeded_parameters


def run_cases(ns, data_cls):
    out = {}
    g = torch.Generator().manual_seed(31)
    sizes = [22, 27]
    n = sum(sizes)
    batch = torch.repeat_interleave(torch.arange(2), torch.tensor(sizes))
    pos = torch.rand((n, 3), generator=g)
    C = 5
    x = torch.randn((n, C), generator=g)
    extra = getattr(ns, "extra", {})

    targets = torch.sort(torch.randperm(n, generator=g)[:13]).values
    row, col = knn(pos, pos[targets], 4, batch, batch[targets])
    edge_index = torch.stack([col, row])[:, 3:]                # target 0 keeps one edge only
    conv = ns.Convolution(local_nn=[10, 8, C], global_nn=[C, 12])
    seeded_parameters(conv, 500)
    conv.eval()
    out["conv/state_dict_keys"] = np.array(list(conv.state_dict().keys()))
    xin = x.clone().requires_grad_(True)
    res = conv(xin, (pos, pos[targets]), edge_index)
    out["conv/x"] = res.detach().numpy()
    (res * torch.linspace(0.5, 1.5, res.numel()).reshape(res.shape)).sum().backward()
    out["conv/dx"] = xin.grad.numpy()
    bare = ns.Convolution(local_nn=[10, 6, 3])
    seeded_parameters(bare, 501)
    bare.eval()
    row2, col2 = knn(pos, pos, 3, batch, batch)
    out["conv_positions_as_features/x"] = bare(None, pos, torch.stack([col2, row2])).detach().numpy()
    lonely = torch.stack([col, row])[:, 8:]                    # targets 0 and 1 have no edge at all: they get 0
    out["conv_targets_without_edges/x"] = conv(x, (pos, pos[targets]), lonely).detach().numpy()

    perm = torch.randperm(n, generator=g)
    perm = torch.cat([perm[batch[perm] == b] for b in range(2)])
    down = ns.RSConvDown(ratio=1.0, radius=0.35, local_nn=[10, 8, C], down_conv_nn=[C, 9, 14], index=2, **extra)
    down.sampler = FixedSampler(perm)
    seeded_parameters(down, 502)
    down.eval()
    res = down(data_cls(x=x, pos=pos, batch=batch, y=torch.arange(n)))
    out["down/state_dict_keys"] = np.array(list(down.state_dict().keys()))
    for key in ("x", "pos", "batch", "idx", "edge_index", "y"):
        t = getattr(res, key)
        out["down/" + key] = t.detach().numpy()
    out["down/finder"] = np.array([float(down.neighbour_finder._radius), float(down.neighbour_finder._max_num_neighbors)])
    return out
