"""Host-side mirror of torch_points3d/core/base_conv/partial_dense.py (the PARTIAL_DENSE convolution API):

  BasePartialDenseConvolutionDown.forward   partial_dense.py:60-81   sampler -> self ball query -> shadow-padded gathers -> conv
  GlobalPartialDenseBaseModule              partial_dense.py:84-100
  FPModule_PD                               partial_dense.py:103-146 KNNInterpolate(k) upsampling + skip + MLP (the KPConv decoder)

plus the two small modules of core/common_modules/base_modules.py they are built from (FastBatchNorm1d :128-153,
MLP :35-45) with the reference's parameter names, so a reference state_dict loads key for key
(`bn.batch_norm.weight`, `nn.0.0.weight`, `nn.0.1.batch_norm.running_mean` ...).

Searches and interpolation tables come from libp3d_b200.so (ops.py); MLPs / BatchNorm are stock torch.
"""
import torch

from . import ops
from .spatial_ops import KNNInterpolate


class FastBatchNorm1d(torch.nn.Module):
    """base_modules.py:128-153.  Same parameters under the same names (`batch_norm.*`).  The reference reshapes
    [N, C] to [1, C, N] around nn.BatchNorm1d because the 2-D path of old cuDNN builds was slow; the statistics are
    the same per-channel mean / biased variance over the N rows either way, so the 2-D call is made directly."""

    def __init__(self, num_features, momentum=0.1, **kwargs):
        super().__init__()
        self.batch_norm = torch.nn.BatchNorm1d(num_features, momentum=momentum, **kwargs)

    def forward(self, x):
        if x.dim() == 2:
            return self.batch_norm(x)
        if x.dim() == 3:
            return self.batch_norm(x.permute(0, 2, 1)).permute(0, 2, 1)
        raise ValueError("Non supported number of dimensions {}".format(x.dim()))


def MLP(channels, activation=None, bn_momentum=0.1, bias=True):
    """base_modules.py:35-45"""
    activation = activation if activation is not None else torch.nn.LeakyReLU(0.2)
    return torch.nn.Sequential(*[
        torch.nn.Sequential(ops.Linear(channels[i - 1], channels[i], bias=bias),
                            FastBatchNorm1d(channels[i], momentum=bn_momentum), activation)
        for i in range(1, len(channels))])


class _Data:
    """attribute bag standing in for torch_geometric's Data / Batch; `clone` copies the attribute table like
    Data.clone() does for the code paths mirrored here (tensors are shared until reassigned)"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def clone(self):
        return _Data(**self.__dict__)


def copy_from_to(data, batch):
    """partial_dense.py:29-32"""
    for key in data.keys:
        if key not in batch.keys:
            setattr(batch, key, getattr(data, key, None))


class BasePartialDenseConvolutionDown(torch.nn.Module):
    """partial_dense.py:35-81.  forward: idx_sampler = sampler(pos, x, batch); idx_neighbour = neighbour_finder(pos,
    pos, batch, batch) (a partial-dense ball query of the cloud on itself, padded with the shadow index N); the
    features and positions get one shadow row appended; `conv` receives the gathered neighbour features [N, M, C] and
    the neighbour positions centred on their query point [N, M, 3].

    The reference reads `self.shadow_features_fill` / `self.shadow_points_fill_` without ever defining them
    (base_conv.py:6-10 sets only sampler and neighbour_finder), so its forward raises AttributeError unless a subclass
    provides them; here they are constructor arguments with the shadow convention of the KPConv path as defaults
    (zero feature row, far-away point: convolution_ops.py:49-50,92)."""

    CONV_TYPE = "partial_dense"

    def __init__(self, sampler, neighbour_finder, *args, shadow_features_fill=0.0, shadow_points_fill=1e6, **kwargs):
        super().__init__()
        self.sampler = sampler
        self.neighbour_finder = neighbour_finder
        self._index = kwargs.get("index", None)
        self.shadow_features_fill = shadow_features_fill
        self.shadow_points_fill_ = shadow_points_fill

    def conv(self, x, pos, x_neighbour, pos_centered_neighbour, idx_neighbour, idx_sampler):
        raise NotImplementedError

    def forward(self, data, **kwargs):
        feats, points, cloud_of = data.x, data.pos, data.batch
        n = points.shape[0]
        idx_sampler = self.sampler(pos=points, x=feats, batch=cloud_of)
        idx_neighbour = self.neighbour_finder(points, points, batch_x=cloud_of, batch_y=cloud_of)  # [n, M], shadow = n
        # one shadow row behind the table, so that the shadow index gathers the fill values
        feats = torch.cat([feats, feats.new_full((1,) + tuple(feats.shape[1:]), self.shadow_features_fill)], dim=0)
        points = torch.cat([points, torch.full((1,) + tuple(points.shape[1:]), self.shadow_points_fill_,
                                               dtype=points.dtype, device=feats.device)], dim=0)
        centred = points[idx_neighbour] - points[:n].unsqueeze(1)  # neighbour positions relative to their query
        out = _Data(x=self.conv(feats, points, feats[idx_neighbour], centred, idx_neighbour, idx_sampler),
                    pos=points[idx_sampler], batch=cloud_of[idx_sampler])
        copy_from_to(data, out)
        return out


def _global_pool(x, batch, num_clouds, aggr):
    """torch_geometric global_max_pool / global_mean_pool over a sorted batch vector"""
    if aggr == "max":
        return ops.segment_max(x, batch, num_clouds)
    s = ops.segment_sum(x, batch, num_clouds)
    cnt = torch.bincount(batch, minlength=num_clouds).clamp(min=1).to(x.dtype).unsqueeze(1)
    return s / cnt


class GlobalPartialDenseBaseModule(torch.nn.Module):
    """partial_dense.py:84-100: MLP over cat(x, pos), then a per-cloud max / mean pool; the pooled clouds become one
    point each at the origin."""

    def __init__(self, nn, aggr="max", *args, **kwargs):
        super().__init__()
        self.nn = MLP(nn)
        self.aggr = aggr

    def forward(self, data, **kwargs):
        cloud_of = data.batch
        num_clouds = kwargs.get("num_clouds")
        if num_clouds is None:
            num_clouds = int(cloud_of[-1].item()) + 1  # sizes the output (global_max_pool does the same read)
        pooled = _global_pool(self.nn(torch.cat([data.x, data.pos], dim=1)), cloud_of, num_clouds, self.aggr)
        out = _Data(x=pooled, pos=data.pos.new_zeros((pooled.shape[0], 3)),
                    batch=torch.arange(pooled.shape[0], device=pooled.device))
        copy_from_to(data, out)
        return out


class FPModule_PD(torch.nn.Module):
    """partial_dense.py:103-146: upsample the coarse level's features onto the skip level with KNNInterpolate(up_k)
    (interpolate.py:7-70 -> libp3d_b200's knn), concatenate the skip features, MLP.

    `has_innermost` (the coarse level is one pooled row per cloud) is decided like the reference does
    (`len(data.x) == data.batch.max() + 1`, one device-to-host read) unless the caller passes it."""

    def __init__(self, up_k, up_conv_nn, *args, **kwargs):
        super().__init__()
        self.upsample_op = KNNInterpolate(up_k)
        bn_momentum = kwargs.get("bn_momentum", 0.1)
        self.nn = MLP(up_conv_nn, bn_momentum=bn_momentum, bias=False)

    def forward(self, data, precomputed=None, has_innermost=None, **kwargs):
        coarse, skip = data
        out = skip.clone()
        if has_innermost is None:
            has_innermost = len(coarse.x) == int(coarse.batch.max().item()) + 1
        table = None
        if precomputed and not has_innermost:
            # position in the list of precomputed upsampling tables: carried from module to module on the data
            out.up_idx = coarse.up_idx if hasattr(coarse, "up_idx") else 0
            table = precomputed[out.up_idx]
            out.up_idx = out.up_idx + 1
        if has_innermost:  # one pooled row per cloud: every point takes the row of its cloud
            x = coarse.x.gather(0, skip.batch.unsqueeze(-1).repeat((1, coarse.x.shape[-1])))
        else:
            x = self.upsample_op(coarse, skip, precomputed=table)
        if skip.x is not None:
            x = torch.cat([x, skip.x], dim=1)
        out.x = self.nn(x) if hasattr(self, "nn") else x
        return out
