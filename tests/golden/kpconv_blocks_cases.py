"""Case list and stand-ins shared by tests/golden/make_golden_kpconv_blocks.py (the REFERENCE's own
modules/KPConv/blocks.py + kernels.py + convolution_ops.py, loaded by file path: KPConv_ops itself is the reference's
torch code, run on CPU) and tests/test_reference_kpconv_blocks.py (torch_points3d_b200.kpconv with `ops.*` replaced by
oracle-backed stand-ins).  What is compared is the block orchestration of SURVEY row a13: search radius
(2.5 * sigma * prev_grid_size), kernel extent, strided queries, unary layers, batch norm, activation, the
max-over-neighbours shortcut, block counters, parameter names.
"""
import numpy as np
import torch

import oracle


def _ptr(batch, n=None):
    counts = torch.bincount(batch, minlength=int(batch.max()) + 1 if n is None else n)
    return np.concatenate([[0], np.cumsum(counts.numpy())]).astype(np.int64)


def ball_query(radius, nsample, x, y, mode="dense", batch_x=None, batch_y=None, sort=False, return_dist=True):
    """torch_points_kernels.ball_query(mode="partial_dense"): table [n, nsample], shadow index = x.shape[0]"""
    assert mode == "partial_dense"
    nclouds = int(max(batch_x.max(), batch_y.max())) + 1
    idx = oracle.ball_query_csr(float(radius), int(nsample), x.detach().numpy(), y.detach().numpy(),
                                _ptr(batch_x, nclouds), _ptr(batch_y, nclouds))[0]
    return torch.from_numpy(np.ascontiguousarray(idx)).long(), None


def grid_means(pos, batch, size):
    """voxel means of GridSampling3D(size, mode="mean"): (pos [n, 3], batch [n]) in the reference's voxel order"""
    cluster, representative, _ = oracle.grid_cluster(pos.detach().numpy(), float(size), batch.numpy())
    cluster = torch.from_numpy(np.asarray(cluster)).long()
    n = int(representative.shape[0])
    cnt = torch.zeros(n).index_add_(0, cluster, torch.ones(pos.shape[0]))
    mean = torch.zeros((n, 3)).index_add_(0, cluster, pos) / cnt.unsqueeze(1)
    qb = torch.zeros(n, dtype=torch.long).index_copy_(0, cluster, batch)
    return mean, qb


class GridSampling3D:
    """stand-in for core/data_transform GridSampling3D(size) on the reference side: the strided blocks read `pos` and
    `batch` of what it returns (blocks.py:74-87)"""

    def __init__(self, size, **kw):
        self._grid_size = size

    def __call__(self, data):
        batch = getattr(data, "batch", None)
        if batch is None:                      # one sample on its own (MultiScaleTransform in a DataLoader worker)
            data.pos = grid_means(data.pos, torch.zeros(data.pos.shape[0], dtype=torch.long), self._grid_size)[0]
        else:
            data.pos, data.batch = grid_means(data.pos, batch, self._grid_size)
        data.x = None
        return data


def seeded_parameters(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if name.endswith("num_batches_tracked"):
                continue
            if name.endswith("running_var"):
                t.copy_(torch.rand(t.shape, generator=g) + 0.5)
            elif name.endswith("K_points"):
                continue
            else:
                t.copy_(torch.randn(t.shape, generator=g) * 0.3)
        # kernel points: seeded positions inside the layer's kernel radius (the reference draws a random rotation of
        # the stored disposition at construction; the mirror does not have to reproduce that draw)
        g2 = torch.Generator().manual_seed(seed + 1000)
        for m in module.modules():
            if hasattr(m, "K_points") and hasattr(m, "kernel_radius"):
                m.K_points.copy_((torch.rand(m.K_points.shape, generator=g2) * 2 - 1) * 0.6 * float(m.kernel_radius))


def make_cloud(seed=4):
    g = torch.Generator().manual_seed(seed)
    sizes = [70, 55]
    pos = torch.rand((sum(sizes), 3), generator=g) * torch.tensor([0.3, 0.3, 0.12])
    batch = torch.repeat_interleave(torch.arange(2), torch.tensor(sizes))
    x = torch.randn((sum(sizes), 5), generator=g)
    return pos, x, batch


def _record(out, tag, block, res):
    out[tag + "/state_dict_keys"] = np.array(list(block.state_dict().keys()))
    out[tag + "/x"] = res.x.detach().numpy()
    out[tag + "/pos"] = res.pos.detach().numpy()
    out[tag + "/batch"] = res.batch.numpy()
    out[tag + "/idx_neighboors"] = res.idx_neighboors.numpy()
    out[tag + "/block_idx"] = np.array(int(res.block_idx))


def run_cases(ns, data_cls):
    """ns: SimpleBlock, ResnetBBlock, KPDualBlock; data_cls(pos=, x=, batch=) makes the data objects"""
    out = {}
    pos, x, batch = make_cloud()
    g0 = 0.04

    simple = ns.SimpleBlock(down_conv_nn=[5, 8], grid_size=g0, prev_grid_size=g0, sigma=1.0, max_num_neighbors=12)
    seeded_parameters(simple, 300)
    simple.eval()
    s_out = simple(data_cls(pos=pos, x=x, batch=batch))
    _record(out, "simple", simple, s_out)
    out["simple/search_radius"] = np.array(float(simple.neighbour_finder._radius))
    out["simple/kernel_radius"] = np.array(float(simple.kp_conv.kernel_radius))
    out["simple/has_sampler"] = np.array(simple.sampler is not None)

    simple.train()                                    # batch statistics + the running averages they leave
    t_out = simple(data_cls(pos=pos, x=x, batch=batch))
    out["simple_train/x"] = t_out.x.detach().numpy()
    out["simple_train/running_mean"] = simple.bn.batch_norm.running_mean.numpy().copy()
    out["simple_train/running_var"] = simple.bn.batch_norm.running_var.numpy().copy()

    res = ns.ResnetBBlock(down_conv_nn=[8, 16], grid_size=g0, prev_grid_size=g0, sigma=1.0, max_num_neighbors=12)
    seeded_parameters(res, 301)
    res.eval()
    r_in = data_cls(pos=pos, x=s_out.x.detach(), batch=batch)
    r_out = res(r_in)
    _record(out, "resnet", res, r_out)

    strided = ns.ResnetBBlock(down_conv_nn=[16, 32], grid_size=2 * g0, prev_grid_size=g0, sigma=1.0,
                              max_num_neighbors=10)
    seeded_parameters(strided, 302)
    strided.eval()
    st_out = strided(data_cls(pos=pos, x=r_out.x.detach(), batch=batch))
    _record(out, "strided", strided, st_out)
    out["strided/is_strided"] = np.array(bool(strided.is_strided))
    out["strided/search_radius"] = np.array(float(strided.neighbour_finder._radius))

    wide = ns.ResnetBBlock(down_conv_nn=[5, 6, 5], grid_size=g0, prev_grid_size=g0, sigma=1.0, max_num_neighbors=12,
                           has_bottleneck=True)   # explicit bottleneck width, identity shortcut
    seeded_parameters(wide, 303)
    wide.eval()
    _record(out, "resnet_identity_shortcut", wide, wide(data_cls(pos=pos, x=x, batch=batch)))

    plain = ns.ResnetBBlock(down_conv_nn=[5, 12], grid_size=g0, prev_grid_size=g0, sigma=1.0, max_num_neighbors=12,
                            has_bottleneck=False)
    seeded_parameters(plain, 304)
    plain.eval()
    _record(out, "resnet_no_bottleneck", plain, plain(data_cls(pos=pos, x=x, batch=batch)))

    dual = ns.KPDualBlock(block_names=["SimpleBlock", "ResnetBBlock"], down_conv_nn=[[5, 8], [8, 16]],
                          grid_size=[g0, g0], prev_grid_size=[g0, g0], has_bottleneck=[False, True],
                          max_num_neighbors=[12, 12], deformable=[False, False], sigma=1.0)
    seeded_parameters(dual, 305)
    dual.eval()
    _record(out, "dual", dual, dual(data_cls(pos=pos, x=x, batch=batch)))

    dual2 = ns.KPDualBlock(block_names=["ResnetBBlock", "ResnetBBlock"], down_conv_nn=[[16, 16], [16, 32]],
                           grid_size=[2 * g0, 2 * g0], prev_grid_size=[g0, 2 * g0], has_bottleneck=[True, True],
                           max_num_neighbors=[10, 14], deformable=[False, False], sigma=1.0)
    seeded_parameters(dual2, 306)
    dual2.eval()
    _record(out, "dual_strided", dual2, dual2(data_cls(pos=pos, x=r_out.x.detach(), batch=batch)))

    # deformable blocks (kernels.py:107-256): offset convolution -> offsets (* point_influence) [-> modulations
    # 2 * sigmoid] -> deformable aggregation -> the regularisers kept on the layer
    for tag, modulated, loss_mode in (("deform", False, "fitting"), ("deform_modulated", True, "fitting"),
                                      ("deform_permissive", False, "permissive")):
        blk = ns.SimpleBlock(down_conv_nn=[5, 8], grid_size=g0, prev_grid_size=g0, sigma=1.0, max_num_neighbors=12,
                             deformable=True, modulated=modulated, loss_mode=loss_mode)
        seeded_parameters(blk, 310)
        blk.eval()
        _record(out, tag, blk, blk(data_cls(pos=pos, x=x, batch=batch)))
        out[tag + "/search_radius"] = np.array(float(blk.neighbour_finder._radius))
        for name, value in blk.kp_conv.get_internal_losses().items():
            out[tag + "/" + name] = np.array(float(value), dtype=np.float64)
    return out
