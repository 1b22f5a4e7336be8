"""Host-side mirror of the dense PointNet++ layers that drive the operators:
  BaseDenseConvolutionDown.forward   torch_points3d/core/base_conv/dense.py:60-87
  PointNetMSGDown                    torch_points3d/modules/pointnet2/dense.py:11-75
  DenseFPModule                      torch_points3d/core/base_conv/dense.py:126-147
Same constructor arguments and tensor layouts (pos [B,N,3], x [B,C,N]); the MLPs are stock
PyTorch (1x1 Conv2d + BatchNorm2d + LeakyReLU, common_modules/dense_modules.py:5-29), the
sampling / grouping / interpolation steps run on libp3d_b200.so.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .spatial_ops import DenseFPSSampler, DenseRadiusNeighbourFinder


def MLP2D(channels, bias=False, bn=True, activation=None):
    """dense_modules.py:5-29: one (1x1 Conv2d, BatchNorm2d, activation) group per layer, nested like the reference's
    Seq-of-Seq so that the parameter names are the reference's (`0.0.weight`, `0.1.running_mean`, `1.0.weight` ...)"""
    layers = []
    for i in range(len(channels) - 1):
        group = [nn.Conv2d(channels[i], channels[i + 1], kernel_size=(1, 1), stride=(1, 1), bias=bias)]
        if bn:
            group.append(nn.BatchNorm2d(channels[i + 1]))
        group.append(activation if activation is not None else nn.LeakyReLU(negative_slope=0.01))
        layers.append(nn.Sequential(*group))
    return nn.Sequential(*layers)


class PointNetMSGDown(nn.Module):
    """modules/pointnet2/dense.py:11-75 on top of core/base_conv/dense.py:60-87."""

    def __init__(self, npoint=None, radii=None, nsample=None, down_conv_nn=None, bn=True,
                 activation=None, use_xyz=True, normalize_xyz=False, fused_grouping=True, **kwargs):
        super().__init__()
        assert len(radii) == len(nsample) == len(down_conv_nn)
        self.sampler = DenseFPSSampler(num_to_sample=npoint)
        self.neighbour_finder = DenseRadiusNeighbourFinder(radii, nsample)
        self.use_xyz = use_xyz
        self.npoint = npoint
        self.mlps = nn.ModuleList([MLP2D(nn_i, bn=bn, activation=activation, bias=False) for nn_i in down_conv_nn])
        self.radii = radii
        self.normalize_xyz = normalize_xyz
        self.fused_grouping = fused_grouping

    def _prepare_features(self, x, pos, new_pos, idx, scale_idx):
        if self.fused_grouping:
            return ops.group_features(x, pos, new_pos, idx,
                                      normalize_by=self.radii[scale_idx] if self.normalize_xyz else None,
                                      use_xyz=self.use_xyz or x is None)
        # the reference's op-by-op formulation (dense.py:36-54)
        new_pos_trans = pos.transpose(1, 2).contiguous()
        grouped_pos = ops.grouping_operation(new_pos_trans, idx)
        grouped_pos = grouped_pos - new_pos.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_pos = grouped_pos / self.radii[scale_idx]
        if x is not None:
            grouped_features = ops.grouping_operation(x, idx)
            return torch.cat([grouped_pos, grouped_features], dim=1) if self.use_xyz else grouped_features
        assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        return grouped_pos

    def conv(self, x, pos, new_pos, radius_idx, scale_idx):
        assert scale_idx < len(self.mlps)
        new_features = self._prepare_features(x, pos, new_pos, radius_idx, scale_idx)
        new_features = self.mlps[scale_idx](new_features)
        new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        return new_features.squeeze(-1)

    def forward(self, data, sample_idx=None, **kwargs):
        x, pos = data.x, data.pos
        if sample_idx is None and self.fused_grouping and hasattr(self.sampler, "_num_to_sample"):
            # FPS writes the sampled coordinates itself (== pos.gather(1, idx) of dense.py:75-76)
            idx1, new_pos = ops.furthest_point_sample(pos, self.sampler._num_to_sample, return_xyz=True)
            idx = idx1.unsqueeze(-1)
        else:
            idx = sample_idx if sample_idx is not None else self.sampler(pos)
            idx = idx.unsqueeze(-1).repeat(1, 1, pos.shape[-1]).long()
            new_pos = pos.gather(1, idx)
        ms_x = []
        nf = self.neighbour_finder
        if self.fused_grouping and 1 < nf.num_scales <= 4:
            # every scale queries the same (pos, new_pos) pair: one sweep, distances computed once
            tables = ops.ball_query_multi(nf._radius, nf._max_num_neighbors, pos, new_pos)
        else:
            tables = [nf(pos, new_pos, scale_idx=s) for s in range(nf.num_scales)]
        for scale_idx in range(nf.num_scales):
            ms_x.append(self.conv(x, pos, new_pos, tables[scale_idx], scale_idx))
        return SimpleNamespace(pos=new_pos, x=torch.cat(ms_x, 1), sampling_id=idx[:, :, 0])


class DenseFPModule(nn.Module):
    """core/base_conv/dense.py:102-147: three_nn + inverse-distance weights + three_interpolate,
    concatenate the skip features, MLP."""

    def __init__(self, up_conv_nn, bn=True, bias=False, activation=None, **kwargs):
        super().__init__()
        self.nn = MLP2D(up_conv_nn, bn=bn, activation=activation, bias=False)

    def conv(self, pos, pos_skip, x):
        assert pos_skip.shape[2] == 3
        if pos is not None:
            # three_nn also emits the weights 1/(d+1e-8) / sum(1/(d+1e-8)) of dense.py:137-139
            _, idx, weight = ops.three_nn(pos_skip, pos, return_weight=True)
            return ops.three_interpolate(x, idx, weight)
        return x.expand(*(x.size()[0:2] + (pos_skip.size(1),)))

    def forward(self, data, **kwargs):
        data, data_skip = data
        new_features = self.conv(data.pos, data_skip.pos, data.x)
        if data_skip.x is not None:
            new_features = torch.cat([new_features, data_skip.x], dim=1)
        new_features = self.nn(new_features.unsqueeze(-1))
        return SimpleNamespace(x=new_features.squeeze(-1), pos=data_skip.pos)


def Conv1D(in_channels, out_channels, bias=True, bn=True, activation="default"):
    """common_modules/dense_modules.py:15-22"""
    layers = [nn.Conv1d(in_channels, out_channels, kernel_size=1, bias=bias)]
    if bn:
        layers.append(nn.BatchNorm1d(out_channels))
    if activation == "default":
        activation = nn.LeakyReLU(negative_slope=0.01)
    if activation:
        layers.append(activation)
    return nn.Sequential(*layers)


class GlobalDenseBaseModule(nn.Module):
    """core/base_conv/dense.py:150-185: MLP over cat(x, pos^T), max / mean over the points, `pos = None` afterwards"""

    def __init__(self, nn, aggr="max", bn=True, activation=None, **kwargs):
        super().__init__()
        self.nn = MLP2D(nn, bn=bn, activation=activation, bias=False)
        if aggr.lower() not in ["mean", "max"]:
            raise Exception("The aggregation provided is unrecognized {}".format(aggr))
        self._aggr = aggr.lower()

    def forward(self, data, **kwargs):
        x, pos = data.x, data.pos
        x = self.nn(torch.cat([x, pos.transpose(1, 2).contiguous()], dim=1).unsqueeze(-1))
        x = x.squeeze(-1).max(-1)[0] if self._aggr == "max" else x.squeeze(-1).mean(-1)
        return SimpleNamespace(x=x.unsqueeze(-1), pos=None)


class PointNet2_D(nn.Module):
    """models/segmentation/pointnet2.py:19-113 with the `pointnet2_charlesmsg` architecture resolved
    (conf/models/segmentation/pointnet2.yaml:95-130): two multi-scale-grouping SA levels, the global module, three
    feature-propagation levels wired as UnetBasedModel nests them (skip connections to the level's input), and the
    mlp_cls head (Conv1D 128 -> 128 + BN, Dropout 0.5, Conv1D 128 -> classes).  forward(pos [B,N,3], x [B,N,FEAT])
    returns the [B*N, classes] logits of pointnet2.py:103; `loss` is its cross-entropy (:108-110)."""

    def __init__(self, feat_dim=3, num_classes=50, dropout=0.5):
        super().__init__()
        F_ = feat_dim
        self.down = nn.ModuleList([
            PointNetMSGDown(npoint=512, radii=[0.1, 0.2, 0.4], nsample=[32, 64, 128],
                            down_conv_nn=[[F_ + 3, 32, 32, 64], [F_ + 3, 64, 64, 128], [F_ + 3, 64, 96, 128]]),
            PointNetMSGDown(npoint=128, radii=[0.4, 0.8], nsample=[64, 128],
                            down_conv_nn=[[64 + 128 + 128 + 3, 128, 128, 256], [64 + 128 + 128 + 3, 128, 196, 256]])])
        self.inner = GlobalDenseBaseModule([256 * 2 + 3, 256, 512, 1024])
        self.up = nn.ModuleList([DenseFPModule([1024 + 256 * 2, 256, 256]),
                                 DenseFPModule([256 + 128 * 2 + 64, 256, 128]),
                                 DenseFPModule([128 + F_, 128, 128])])
        self.FC_layer = nn.Sequential(Conv1D(128, 128, bn=True, bias=False), nn.Dropout(p=dropout),
                                      Conv1D(128, num_classes, activation=None, bias=True, bn=False))
        self._num_classes = num_classes

    def forward(self, pos, x):
        d0 = SimpleNamespace(pos=pos, x=None if x is None else x.transpose(1, 2).contiguous())
        d1 = self.down[0](d0)
        d2 = self.down[1](d1)
        u = self.inner(d2)
        u = self.up[0]((u, d2))
        u = self.up[1]((u, d1))
        u = self.up[2]((u, d0))
        return self.FC_layer(u.x).transpose(1, 2).contiguous().view((-1, self._num_classes))

    def loss(self, logits, labels):
        return F.cross_entropy(logits, torch.flatten(labels).long(), ignore_index=-1)  # IGNORE_LABEL, pointnet2.py:108-110
