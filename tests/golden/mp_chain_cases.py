"""Case list and stand-ins shared by tests/golden/make_golden_mp_chain.py (the REFERENCE's own MESSAGE_PASSING
PointNet++ modules: modules/pointnet2/message_passing.py SAModule over core/base_conv/message_passing.py
BaseMSConvolutionDown / GlobalBaseModule / FPModule, loaded by file path) and tests/test_reference_mp_chain.py
(torch_points3d_b200.message_passing with `ops.*` replaced by stand-ins of the same definitions).
"""
import numpy as np
import torch

import oracle


def _ptr(batch, n):
    counts = torch.bincount(batch, minlength=n)
    return np.concatenate([[0], np.cumsum(counts.numpy())]).astype(np.int64)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, **kw):
    """torch_cluster.radius: [2, E] = (y index, x index), at most max_num_neighbors per y, inside its cloud"""
    n = int(max(batch_x.max(), batch_y.max())) + 1
    row, col = oracle.radius_csr(x.numpy(), y.numpy(), float(r), _ptr(batch_x, n), _ptr(batch_y, n),
                                 max_num_neighbors=int(max_num_neighbors))[:2]
    return torch.stack([torch.from_numpy(np.asarray(row)).long(), torch.from_numpy(np.asarray(col)).long()])


def knn(x, y, k, batch_x=None, batch_y=None, **kw):
    bx = batch_x if batch_x is not None else torch.zeros(x.shape[0], dtype=torch.long)
    by = batch_y if batch_y is not None else torch.zeros(y.shape[0], dtype=torch.long)
    d = torch.cdist(y.double(), x.double())
    d[by.unsqueeze(1) != bx.unsqueeze(0)] = float("inf")
    col = d.topk(k, dim=1, largest=False).indices
    row = torch.arange(y.shape[0]).unsqueeze(1).expand(-1, k)
    return torch.stack([row.reshape(-1), col.reshape(-1)])


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3, **kw):
    """torch_geometric 1.7.2 knn_interpolate"""
    with torch.no_grad():
        y_idx, x_idx = knn(pos_x, pos_y, k, batch_x, batch_y)
        diff = pos_x[x_idx] - pos_y[y_idx]
        w = 1.0 / torch.clamp((diff * diff).sum(dim=-1, keepdim=True), min=1e-16)
    num = torch.zeros((pos_y.shape[0], x.shape[1]), dtype=x.dtype).index_add_(0, y_idx, x[x_idx] * w)
    den = torch.zeros((pos_y.shape[0], 1), dtype=x.dtype).index_add_(0, y_idx, w)
    return num / den


def global_max_pool(x, batch, size=None):
    n = int(batch.max()) + 1 if size is None else size
    return torch.stack([x[batch == b].max(dim=0).values for b in range(n)])


def global_mean_pool(x, batch, size=None):
    n = int(batch.max()) + 1 if size is None else size
    return torch.stack([x[batch == b].mean(dim=0) for b in range(n)])


def segment_max(msg, dst, T):
    """per-target max of the edge messages, 0 for a target without edges (torch_scatter's `max` fill)"""
    out = torch.zeros((T, msg.shape[1]), dtype=msg.dtype)
    for t in torch.unique(dst).tolist():
        out[t] = msg[dst == t].max(dim=0).values
    return out


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
    """ns: SAModule, GlobalBaseModule, FPModule"""
    out = {}
    g = torch.Generator().manual_seed(31)
    sizes = [30, 26]
    n = sum(sizes)
    batch = torch.repeat_interleave(torch.arange(2), torch.tensor(sizes))
    pos = torch.rand((n, 3), generator=g)
    C = 4
    x = torch.randn((n, C), generator=g)
    keep = torch.sort(torch.cat([torch.randperm(30, generator=g)[:14], 30 + torch.randperm(26, generator=g)[:12]])).values

    sa = ns.SAModule(ratio=0.5, radius=[0.3, 0.5], radius_num_point=[5, 9], down_conv_nn=[C + 3, 8, 10], index=1)
    sa.sampler = FixedSampler(keep)
    seeded_parameters(sa, 500)
    sa.eval()
    level1 = sa(data_cls(x=x, pos=pos, batch=batch, y=torch.arange(n)))
    out["sa/state_dict_keys"] = np.array(list(sa.state_dict().keys()))
    out["sa/x"] = level1.x.detach().numpy()
    out["sa/pos"] = level1.pos.numpy()
    out["sa/batch"] = level1.batch.numpy()
    out["sa/idx"] = level1.idx.numpy()
    out["sa/copied_keys"] = np.array(sorted(k for k in level1.keys if k not in ("x", "pos", "batch", "idx")))
    out["sa/scales"] = np.array([sa.neighbour_finder.num_scales] + list(sa.neighbour_finder._max_num_neighbors))
    level1 = data_cls(x=level1.x.detach(), pos=level1.pos, batch=level1.batch)

    for aggr in ("max", "mean"):
        glob = ns.GlobalBaseModule([20 + 3, 9], aggr=aggr)
        seeded_parameters(glob, 501)
        glob.eval()
        pooled = glob(level1)
        out["global_%s/x" % aggr] = pooled.x.detach().numpy()
        out["global_%s/pos" % aggr] = pooled.pos.numpy()
        out["global_%s/batch" % aggr] = pooled.batch.numpy()
    glob = ns.GlobalBaseModule([20, 9])
    seeded_parameters(glob, 502)
    glob.eval()
    no_pos = glob(data_cls(x=level1.x, pos=None, batch=level1.batch))
    out["global_no_pos/x"] = no_pos.x.detach().numpy()
    out["global_no_pos/has_pos"] = np.array(getattr(no_pos, "pos", None) is not None)

    level0 = data_cls(x=x, pos=pos, batch=batch)
    fp = ns.FPModule(3, [20 + C, 12, 6], bn_momentum=0.05)
    seeded_parameters(fp, 503)
    fp.eval()
    up = fp((level1, level0))
    out["fp/state_dict_keys"] = np.array(list(fp.state_dict().keys()))
    out["fp/x"] = up.x.detach().numpy()
    out["fp/pos_is_skip_pos"] = np.array(bool(up.pos is level0.pos))
    fp1 = ns.FPModule(1, [20, 6])
    seeded_parameters(fp1, 504)
    fp1.eval()
    out["fp_no_skip_features/x"] = fp1((level1, data_cls(x=None, pos=pos, batch=batch))).x.detach().numpy()
    fp2 = ns.FPModule(2, [20 + C, 6], skip=False)     # what the `skip` keyword does in the reference: see the test
    seeded_parameters(fp2, 505)
    fp2.eval()
    out["fp_skip_keyword/x"] = fp2((level1, level0)).x.detach().numpy()
    return out
