"""Case list and kernel stand-ins shared by tests/golden/make_golden_dense_chain.py (the REFERENCE's own
PointNetMSGDown / DenseFPModule / GlobalDenseBaseModule, loaded by file path, with `torch_points_kernels` replaced by
the CPU stand-ins below) and tests/test_reference_dense_chain.py (torch_points3d_b200.pointnet2 with `ops.*` replaced by
the same stand-ins).  The stand-ins are the oracle's restatements, so what the comparison isolates is the host chain:
sampler -> gather of the centres -> one table per scale -> grouping, centring, optional radius normalisation, channel
order of the concatenation -> shared-per-scale MLP -> max over the neighbours -> concatenation of the scales, and the
three-nearest interpolation weights + skip concatenation of the feature-propagation module.
"""
import numpy as np
import torch

import oracle


class Kernels:
    """torch_points_kernels-shaped CPU stand-ins (oracle underneath); `calls` counts them by name"""

    def __init__(self):
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def furthest_point_sample(self, xyz, npoint):
        self._count("furthest_point_sample")
        return torch.from_numpy(oracle.fps_dense(xyz.detach().numpy(), int(npoint))).long()

    def ball_query(self, radius, nsample, x, y, mode="dense", batch_x=None, batch_y=None, sort=False):
        self._count("ball_query")
        assert mode == "dense" and batch_x is None and batch_y is None
        idx = oracle.ball_query_dense(float(radius), int(nsample), x.detach().numpy(), y.detach().numpy())[0]
        return torch.from_numpy(np.ascontiguousarray(idx)).long(), None

    def grouping_operation(self, features, idx):
        self._count("grouping_operation")
        B, C, N = features.shape
        _, S, K = idx.shape
        return features.gather(2, idx.reshape(B, 1, S * K).expand(B, C, S * K)).reshape(B, C, S, K)

    def three_nn(self, unknown, known):
        self._count("three_nn")
        d2 = ((unknown.unsqueeze(2) - known.unsqueeze(1)) ** 2).sum(-1)
        best = d2.topk(3, dim=2, largest=False)
        return best.values.sqrt(), best.indices

    def three_interpolate(self, features, idx, weight):
        self._count("three_interpolate")
        B, C, m = features.shape
        n = idx.shape[1]
        taps = features.gather(2, idx.reshape(B, 1, n * 3).expand(B, C, n * 3)).reshape(B, C, n, 3)
        return (taps * weight.unsqueeze(1)).sum(-1)


def seeded_parameters(module, seed):
    """fill every parameter and buffer IN state_dict ORDER from one generator (both sides build the same values)"""
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
    """ns: PointNetMSGDown, DenseFPModule, GlobalDenseBaseModule (+ `kwargs` passed to every PointNetMSGDown);
    returns {name: numpy array}"""
    out = {}
    g = torch.Generator().manual_seed(5)
    B, N, C = 2, 96, 3
    pos = torch.rand((B, N, 3), generator=g)
    x = torch.randn((B, C, N), generator=g)
    extra = getattr(ns, "kwargs", {})

    for tag, normalize, use_xyz in (("plain", False, True), ("normalized", True, True), ("no_xyz", False, False)):
        cin = C + 3 if use_xyz else C
        sa = ns.PointNetMSGDown(npoint=24, radii=[0.2, 0.45], nsample=[4, 9],
                                down_conv_nn=[[cin, 8, 10], [cin, 6, 12]], use_xyz=use_xyz, normalize_xyz=normalize,
                                **extra)
        seeded_parameters(sa, 100)
        sa.eval()
        xin = x.clone().requires_grad_(True)
        d1 = sa(data_cls(pos=pos, x=xin))
        out[tag + "/state_dict_keys"] = np.array(list(sa.state_dict().keys()))
        out[tag + "/pos"] = d1.pos.detach().numpy()
        out[tag + "/x"] = d1.x.detach().numpy()
        (gx,) = torch.autograd.grad(d1.x.square().sum(), xin)
        out[tag + "/grad_x"] = gx.numpy()
        if tag == "plain":
            level0, level1 = data_cls(pos=pos, x=x), data_cls(pos=d1.pos.detach(), x=d1.x.detach())

    # positions only (x = None): the grouped relative coordinates are the features
    sa = ns.PointNetMSGDown(npoint=16, radii=[0.3], nsample=[5], down_conv_nn=[[3, 7]], **extra)
    seeded_parameters(sa, 101)
    sa.eval()
    out["xyz_only/x"] = sa(data_cls(pos=pos, x=None)).x.detach().numpy()

    # training mode: batch statistics over (B, S, K) and the running averages they leave behind
    sa = ns.PointNetMSGDown(npoint=24, radii=[0.25], nsample=[6], down_conv_nn=[[C + 3, 9]], **extra)
    seeded_parameters(sa, 102)
    sa.train()
    out["train/x"] = sa(data_cls(pos=pos, x=x)).x.detach().numpy()
    for k, v in sa.state_dict().items():
        if "running" in k:
            out["train/" + k] = v.numpy().copy()

    # feature propagation: level 1 (24 points, 22 channels) onto level 0 (96 points, 3 channels)
    fp = ns.DenseFPModule(up_conv_nn=[22 + C, 16, 8])
    seeded_parameters(fp, 103)
    fp.eval()
    up = fp((level1, level0))
    out["fp/state_dict_keys"] = np.array(list(fp.state_dict().keys()))
    out["fp/x"] = up.x.detach().numpy()
    out["fp/pos_is_skip_pos"] = np.array(bool(up.pos is level0.pos))
    fp2 = ns.DenseFPModule(up_conv_nn=[22, 8])
    seeded_parameters(fp2, 104)
    fp2.eval()
    out["fp_no_skip/x"] = fp2((level1, data_cls(pos=pos, x=None))).x.detach().numpy()

    # global module + propagation from a single pooled column (pos = None branch of the FP module)
    glob = ns.GlobalDenseBaseModule(nn=[22 + 3, 12], aggr="max")
    seeded_parameters(glob, 105)
    glob.eval()
    pooled = glob(level1)
    out["global/x"] = pooled.x.detach().numpy()
    out["global/pos_is_none"] = np.array(pooled.pos is None)
    fp3 = ns.DenseFPModule(up_conv_nn=[12 + 22, 8])
    seeded_parameters(fp3, 106)
    fp3.eval()
    out["fp_from_global/x"] = fp3((pooled, level1)).x.detach().numpy()
    glob_mean = ns.GlobalDenseBaseModule(nn=[22 + 3, 12], aggr="MEAN")
    seeded_parameters(glob_mean, 105)
    glob_mean.eval()
    out["global_mean/x"] = glob_mean(level1).x.detach().numpy()
    try:
        ns.GlobalDenseBaseModule(nn=[4, 4], aggr="sum")
        out["global_bad_aggr"] = np.array("no error")
    except Exception as e:  # noqa: BLE001
        out["global_bad_aggr"] = np.array(str(e))
    return out
