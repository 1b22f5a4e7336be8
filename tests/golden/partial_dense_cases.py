"""Case list and stand-ins shared by tests/golden/make_golden_partial_dense.py (the REFERENCE's own
core/base_conv/partial_dense.py classes, loaded by file path) and tests/test_reference_partial_dense.py
(torch_points3d_b200.partial_dense with `ops.*` replaced by the same stand-ins): the PARTIAL_DENSE down-convolution
chain (sampler -> self ball query -> shadow-padded gathers -> conv), the global pooling module and the KPConv decoder's
FPModule_PD (interpolation with and without precomputed tables, the innermost broadcast, the `up_idx` counter).
"""
import numpy as np
import torch

import oracle


def knn(x, y, k, batch_x=None, batch_y=None, **kw):
    """brute force: the k nearest x of every y inside its cloud, ascending distance; [2, E] = (y index, x index)"""
    bx = batch_x if batch_x is not None else torch.zeros(x.shape[0], dtype=torch.long)
    by = batch_y if batch_y is not None else torch.zeros(y.shape[0], dtype=torch.long)
    d = torch.cdist(y.double(), x.double())
    d[by.unsqueeze(1) != bx.unsqueeze(0)] = float("inf")
    col = d.topk(k, dim=1, largest=False).indices
    row = torch.arange(y.shape[0]).unsqueeze(1).expand(-1, k)
    return torch.stack([row.reshape(-1), col.reshape(-1)])


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3, **kw):
    """torch_geometric 1.7.2 knn_interpolate: y = sum_k w x / sum_k w, w = 1 / clamp(d^2, 1e-16)"""
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


def scatter_add(src, index, dim=0, dim_size=None):
    return torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, src)


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


def run_cases(ns, data_cls):
    """ns: BasePartialDenseConvolutionDown, GlobalPartialDenseBaseModule, FPModule_PD; data_cls(**attrs) makes the
    data objects (needs `.keys`, `.clone()`)"""
    out = {}
    g = torch.Generator().manual_seed(9)
    sizes = [13, 9, 11]
    n = sum(sizes)
    batch = torch.repeat_interleave(torch.arange(3), torch.tensor(sizes))
    pos = torch.rand((n, 3), generator=g) * 0.5
    x = torch.randn((n, 4), generator=g)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

    # ---- the down-convolution chain
    keep = torch.tensor([0, 3, 4, 12, 13, 20, 22, 30, 32])

    def sampler(pos=None, x=None, batch=None):
        return keep

    def finder(px, py, batch_x=None, batch_y=None):
        idx = oracle.ball_query_csr(0.15, 5, px.numpy(), py.numpy(), ptr, ptr)[0]     # shadow index = n
        return torch.from_numpy(np.ascontiguousarray(idx)).long()

    class Down(ns.BasePartialDenseConvolutionDown):
        def __init__(self):
            super().__init__(sampler, finder, index=2)
            self.shadow_features_fill = 0.0
            self.shadow_points_fill_ = 1e6

        def conv(self, x, pos, x_neighbour, pos_centered_neighbour, idx_neighbour, idx_sampler):
            near = (pos_centered_neighbour.norm(dim=-1) < 1.0).to(x.dtype)             # shadow rows are far away
            pooled = (x_neighbour * near.unsqueeze(-1)).sum(1)
            spread = (pos_centered_neighbour * near.unsqueeze(-1)).abs().sum(1)
            return torch.cat([pooled, spread, near.sum(1, keepdim=True)], dim=1)[idx_sampler]

    down = Down()
    res = down(data_cls(x=x, pos=pos, batch=batch, y=torch.arange(n), grid_size=torch.tensor([0.04])))
    out["down/x"] = res.x.numpy()
    out["down/pos"] = res.pos.numpy()
    out["down/batch"] = res.batch.numpy()
    out["down/copied_keys"] = np.array(sorted(k for k in res.keys if k not in ("x", "pos", "batch")))
    out["down/y_is_the_input_y"] = np.array(bool(res.y.shape[0] == n))
    out["down/index"] = np.array(down._index)

    # ---- global module (max and mean)
    level = data_cls(x=x, pos=pos, batch=batch)
    for aggr in ("max", "mean"):
        glob = ns.GlobalPartialDenseBaseModule([4 + 3, 6, 5], aggr=aggr)
        seeded_parameters(glob, 200)
        glob.eval()
        pooled = glob(level)
        out["global_%s/x" % aggr] = pooled.x.detach().numpy()
        out["global_%s/pos" % aggr] = pooled.pos.numpy()
        out["global_%s/batch" % aggr] = pooled.batch.numpy()
        out["global_%s/state_dict_keys" % aggr] = np.array(list(glob.state_dict().keys()))

    # ---- FPModule_PD: coarse level (the kept points) onto the full level
    coarse = data_cls(x=torch.randn((keep.shape[0], 6), generator=g), pos=pos[keep], batch=batch[keep])
    skip = data_cls(x=x, pos=pos, batch=batch)
    fp = ns.FPModule_PD(1, [6 + 4, 8, 5], bn_momentum=0.02)
    seeded_parameters(fp, 201)
    fp.eval()
    out["fp/state_dict_keys"] = np.array(list(fp.state_dict().keys()))
    direct = fp((coarse, skip))
    out["fp_direct/x"] = direct.x.detach().numpy()
    out["fp_direct/has_up_idx"] = np.array(hasattr(direct, "up_idx"))
    tables = [fp.upsample_op.precompute(coarse, skip)]
    pre = fp((coarse, skip), precomputed=tables)
    out["fp_precomputed/x"] = pre.x.detach().numpy()
    out["fp_precomputed/up_idx"] = np.array(int(pre.up_idx))
    fp3 = ns.FPModule_PD(3, [6 + 4, 7])
    seeded_parameters(fp3, 202)
    fp3.eval()
    out["fp_k3/x"] = fp3((coarse, skip)).x.detach().numpy()
    coarse.up_idx = 1
    two = [None, fp3.upsample_op.precompute(coarse, skip)]
    later = fp3((coarse, skip), precomputed=two)
    out["fp_k3_second_table/x"] = later.x.detach().numpy()
    out["fp_k3_second_table/up_idx"] = np.array(int(later.up_idx))
    # innermost: one pooled row per cloud
    inner = data_cls(x=torch.randn((3, 6), generator=g), pos=torch.zeros((3, 3)), batch=torch.arange(3))
    out["fp_innermost/x"] = fp((inner, skip), precomputed=tables).x.detach().numpy()
    no_skip = fp3.__class__(1, [6, 5])
    seeded_parameters(no_skip, 203)
    no_skip.eval()
    out["fp_no_skip_features/x"] = no_skip((coarse, data_cls(x=None, pos=pos, batch=batch))).x.detach().numpy()
    return out
