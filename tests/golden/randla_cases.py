"""Case list and stand-ins shared by tests/golden/make_golden_randla.py (the REFERENCE's own RandLA-Net modules,
modules/RandLANet/modules.py over core/base_conv/message_passing.py, loaded by file path) and
tests/test_reference_randla.py (torch_points3d_b200.message_passing with `ops.*` replaced by torch stand-ins):
RandlaKernel's message (relative position encoding [pos_i, pos_j, pos_i - pos_j, |.|], attention softmax, sum over the
edges of a target, global MLP), the single-scale down-convolution chain BaseConvolutionDown.forward as RandlaConv drives
it, and the residual block around two of them.

The reference chain hands `conv` the position pair as (pos[idx], pos) (message_passing.py:52), which is only
index-safe when the sampler keeps every point; the cases therefore use ratio-1 samplers (fixed permutations), and the
mirror is run with `reference_pos_order=True` (its literal mode).
"""
import numpy as np
import torch


def knn(x, y, k, batch_x=None, batch_y=None, **kw):
    """brute force: the k nearest x of every y inside its cloud, ascending distance; [2, E] = (y index, x index)"""
    bx = batch_x if batch_x is not None else torch.zeros(x.shape[0], dtype=torch.long)
    by = batch_y if batch_y is not None else torch.zeros(y.shape[0], dtype=torch.long)
    d = torch.cdist(y.double(), x.double())
    d[by.unsqueeze(1) != bx.unsqueeze(0)] = float("inf")
    col = d.topk(k, dim=1, largest=False).indices
    row = torch.arange(y.shape[0]).unsqueeze(1).expand(-1, k)
    return torch.stack([row.reshape(-1), col.reshape(-1)])


def seeded_parameters(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if name.endswith("num_batches_tracked"):
                continue
            if name.endswith("running_var"):
                t.copy_(torch.rand(t.shape, generator=g) + 0.5)
            else:
                t.copy_(torch.randn(t.shape, generator=g) * 0.3)


class FixedSampler:
    def __init__(self, idx):
        self.idx = idx

    def __call__(self, pos, batch=None, **kw):
        return self.idx


def run_cases(ns, data_cls):
    """ns: RandlaKernel, RandlaConv, RandLANetRes, extra (kwargs for the chain classes)"""
    out = {}
    g = torch.Generator().manual_seed(21)
    sizes = [24, 26]
    n = sum(sizes)
    batch = torch.repeat_interleave(torch.arange(2), torch.tensor(sizes))
    pos = torch.rand((n, 3), generator=g)
    C = 6
    x = torch.randn((n, C), generator=g)
    extra = getattr(ns, "extra", {})

    # ---- the kernel on a bipartite graph (14 targets drawn from the cloud, 4 nearest sources each)
    targets = torch.sort(torch.randperm(n, generator=g)[:14]).values
    row, col = knn(pos, pos[targets], 4, batch, batch[targets])
    edge_index = torch.stack([col, row])                      # [source index, target index]
    kernel = ns.RandlaKernel(point_pos_nn=[10, 8], attention_nn=[C + 8, C + 8], global_nn=[C + 8, 12])
    seeded_parameters(kernel, 400)
    kernel.eval()
    out["kernel/state_dict_keys"] = np.array(list(kernel.state_dict().keys()))
    out["kernel/x"] = kernel(x, (pos, pos[targets]), edge_index).detach().numpy()
    same = ns.RandlaKernel(point_pos_nn=[10, 8], attention_nn=[3 + 8, 3 + 8], global_nn=[3 + 8, 5])
    seeded_parameters(same, 401)
    same.eval()
    row2, col2 = knn(pos, pos, 3, batch, batch)
    out["kernel_positions_as_features/x"] = same(None, pos, torch.stack([col2, row2])).detach().numpy()

    # ---- the down-convolution chain at ratio 1 (fixed permutation instead of the random draw)
    perm = torch.randperm(n, generator=g)
    perm = torch.cat([perm[batch[perm] == b] for b in range(2)])          # keeps batch[perm] sorted, as knn requires
    conv = ns.RandlaConv(1.0, 5, point_pos_nn=[10, 8], attention_nn=[C + 8, C + 8], down_conv_nn=[C + 8, 12],
                         index=1, **extra)
    conv.sampler = FixedSampler(perm)
    seeded_parameters(conv, 402)
    conv.eval()
    res = conv(data_cls(x=x, pos=pos, batch=batch, y=torch.arange(n)))
    out["conv/state_dict_keys"] = np.array(list(conv.state_dict().keys()))
    out["conv/x"] = res.x.detach().numpy()
    out["conv/pos"] = res.pos.numpy()
    out["conv/batch"] = res.batch.numpy()
    out["conv/idx"] = res.idx.numpy()
    out["conv/edge_index"] = res.edge_index.numpy()
    out["conv/copied_keys"] = np.array(sorted(k for k in res.keys if k not in ("x", "pos", "batch", "idx", "edge_index")))
    out["conv/k"] = np.array(int(conv.neighbour_finder.k))

    # ---- nb_feature rewrite of the first layer's sizes (modules.py:59-63)
    sizes_in = {"point_pos_nn": [10, 8], "attention_nn": [99, 99], "down_conv_nn": [99, 12]}
    first = ns.RandlaConv(1.0, 5, index=0, nb_feature=7, **sizes_in, **extra)
    out["conv_first_layer/sizes"] = np.array([sizes_in["point_pos_nn"][-1], sizes_in["attention_nn"][0],
                                              sizes_in["attention_nn"][-1], sizes_in["down_conv_nn"][0]])
    out["conv_first_layer/state_dict_shapes"] = np.array(
        ["%s:%s" % (k, tuple(v.shape)) for k, v in first.state_dict().items() if k.endswith("weight")])

    # ---- residual block around two convs (k = 16 hard-coded)
    block = ns.RandLANetRes(C, 16, [1.0, 1.0], [[10, 8], [10, 8]], [[C + 8, C + 8], [12 + 8, 12 + 8]],
                            [[C + 8, 12], [12 + 8, 16]], **extra)
    perm2 = torch.randperm(n, generator=g)
    perm2 = torch.cat([perm2[batch[perm][perm2] == b] for b in range(2)])
    block._conv.conv1.sampler = FixedSampler(perm)
    block._conv.conv2.sampler = FixedSampler(perm2)
    seeded_parameters(block, 403)
    block.eval()
    res = block(data_cls(x=x, pos=pos, batch=batch))
    out["block/state_dict_keys"] = np.array(list(block.state_dict().keys()))
    out["block/x"] = res.x.detach().numpy()
    out["block/pos"] = res.pos.numpy()
    out["block/batch"] = res.batch.numpy()
    out["block/idx"] = res.idx.numpy()
    out["block/k"] = np.array([int(block._conv.conv1.neighbour_finder.k), int(block._conv.conv2.neighbour_finder.k)])
    return out
