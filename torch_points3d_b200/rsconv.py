"""Host-side mirror of the dense Relation-Shape convolution layers, the third caller of the sampling / grouping path
(same FPS + ball query + `tp.grouping_operation` chain, no kernel of its own):

  OriginalRSConv, RSConvOriginalMSGDown    torch_points3d/modules/RSConv/dense.py:202-397   (the `RSConv_MSN` benchmark model)
  RSConvMapper, SharedRSConv               torch_points3d/modules/RSConv/dense.py:18-104
  RSConvSharedMSGDown, RSConvMSGDown       torch_points3d/modules/RSConv/dense.py:107-199, 400-482
  RSConv_MSN                               torch_points3d/models/segmentation/rsconv.py:14-124 over
                                           conf/models/segmentation/rsconv.yaml:57-108 (the network the reference benchmarks)
  Convolution, RSConvDown                  torch_points3d/modules/RSConv/message_passing.py:10-60  (MESSAGE_PASSING format,
                                           the `RSConv_2LD` / `RSConv_4LD` models; edge gathers and per-target max of csrc/edges.cu)

Same constructor arguments, tensor layouts (pos [B,N,3], x [B,C,N]) and parameter names; the chain around `conv`
(sampler -> centres -> one neighbour table per scale -> concatenation of the scales) is PointNetMSGDown's
(core/base_conv/dense.py:60-87), the sampling / searching / grouping steps run on libp3d_b200.so, the 1x1 convolutions
and batch norms are stock PyTorch as in the reference.

Two properties of the reference that the numbers depend on are kept:
  * RSConvOriginalMSGDown centres the grouped coordinates IN PLACE (dense.py:346-348), so the block it calls "raw"
    coordinates holds the centred ones too: the relation vector is [|d|, d of the first neighbour, d, d];
  * the three 1x1 convolutions of RSConvOriginalMSGDown are shared by its scales (and therefore appear once per scale
    in the state_dict), the batch norms are per scale.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .message_passing import BaseConvolutionDown
from .partial_dense import MLP
from .pointnet2 import Conv1D, DenseFPModule, GlobalDenseBaseModule, MLP2D, PointNetMSGDown
from .spatial_ops import DenseFPSSampler, DenseRadiusNeighbourFinder, FPSSampler, RadiusNeighbourFinder


def _relation(distance_of, centre, first, second):
    """the 10-channel relation vector [|delta|, centre (3), first (3), second (3)] of dense.py:96,243"""
    return torch.cat((torch.norm(distance_of, p=2, dim=1).unsqueeze(1), centre, first, second), dim=1)


class OriginalRSConv(nn.Module):
    """dense.py:202-258.  input [B, 3 + 3 + C, npoint, nsample] -> [B, C_out, npoint]"""

    def __init__(self, mapping=None, first_layer=False, radius=None, activation=None):
        super().__init__()
        self.nn = nn.ModuleList()
        self._radius = radius
        self.mapping_func1, self.mapping_func2, self.cr_mapping = mapping[0], mapping[1], mapping[2]
        self.first_layer = first_layer
        if first_layer:
            self.xyz_raising = mapping[3]
            self.bn_xyz_raising = nn.BatchNorm2d(self.xyz_raising.out_channels)
            self.nn.append(self.bn_xyz_raising)
        self.bn_mapping = nn.BatchNorm2d(self.mapping_func1.out_channels)
        self.bn_rsconv = nn.BatchNorm2d(self.cr_mapping.in_channels)
        self.bn_channel_raising = nn.BatchNorm1d(self.cr_mapping.out_channels)
        self.nn.extend([self.bn_mapping, self.bn_rsconv, self.bn_channel_raising])
        self.activation = activation if activation is not None else nn.ReLU(inplace=True)

    def forward(self, grouped):
        act = self.activation
        first, second, x = grouped[:, 0:3], grouped[:, 3:6], grouped[:, 3:]
        nsample = x.shape[3]
        relation = _relation(second, first[:, :, :, 0:1].repeat(1, 1, 1, nsample), first, second)
        relation = self.mapping_func2(act(self.bn_mapping(self.mapping_func1(relation))))
        if self.first_layer:
            x = act(self.bn_xyz_raising(self.xyz_raising(x)))
        x = F.max_pool2d(act(self.bn_rsconv(torch.mul(relation, x))), kernel_size=(1, nsample)).squeeze(3)
        return act(self.bn_channel_raising(self.cr_mapping(x)))

    def __repr__(self):
        return "{}({})".format(self.__class__.__name__, self.nn.__repr__())


class _DenseRSConvDown(PointNetMSGDown):
    """what the three RSConv down modules share: FPS sampler + multi-scale dense ball query, and the
    BaseDenseConvolutionDown chain (inherited forward)"""

    def _setup(self, npoint, radii, nsample, down_conv_nn, use_xyz, fused_grouping):
        nn.Module.__init__(self)
        assert len(radii) == len(nsample)
        self.sampler = DenseFPSSampler(num_to_sample=npoint)
        self.neighbour_finder = DenseRadiusNeighbourFinder(radii, nsample)
        self.use_xyz, self.npoint, self.radii = use_xyz, npoint, radii
        self.fused_grouping = fused_grouping
        self.mlps = nn.ModuleList()


class RSConvOriginalMSGDown(_DenseRSConvDown):
    """dense.py:261-397.  `down_conv_nn` = [C_in, C_mid, C_out] of the relation MLP, or for the first layer
    [[C_in, C_mid, C_out], [feat_in, feat_out]] (the second pair lifts the input features); `channel_raising_nn` =
    [in, out] of the 1x1 Conv1d applied after the max over the neighbours."""

    def __init__(self, npoint=None, radii=None, nsample=None, down_conv_nn=None, channel_raising_nn=None, bn=True,
                 bias=True, use_xyz=True, activation=None, fused_grouping=True, **kwargs):
        self._setup(npoint, radii, nsample, down_conv_nn, use_xyz, fused_grouping)
        self.mappings = nn.ModuleList()
        self._first_layer = len(down_conv_nn) == 2
        relation_nn = down_conv_nn[0] if self._first_layer else down_conv_nn

        def conv2d(cin, cout):
            layer = nn.Conv2d(cin, cout, kernel_size=(1, 1), stride=(1, 1), bias=bias)
            nn.init.kaiming_normal_(layer.weight)
            if bias:
                nn.init.constant_(layer.bias, 0)
            return layer

        lift = conv2d(*down_conv_nn[-1]) if self._first_layer else None
        mapping = [conv2d(relation_nn[0], relation_nn[1]), conv2d(relation_nn[1], relation_nn[2])]
        raising = nn.Conv1d(channel_raising_nn[0], channel_raising_nn[1], kernel_size=1, stride=1, bias=bias)
        nn.init.kaiming_normal_(raising.weight)
        nn.init.constant_(raising.bias, 0)  # unconditional in the reference as well (:315): bias=False fails here
        mapping.append(raising)
        if lift is not None:
            mapping.append(lift)
        self.mappings.extend(mapping)
        for radius in radii:
            self.mlps.append(OriginalRSConv(mapping=mapping, first_layer=self._first_layer, radius=radius))

    def _prepare_features(self, xyz, new_xyz, features=None, idx=None):
        """[B, 3 + 3 + C, npoint, nsample]: centred coordinates twice (see the module docstring), then the features"""
        if self.fused_grouping and (self.use_xyz or features is None):
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            grouped = ops.group_features(features, xyz, new_xyz, idx, use_xyz=True)   # [centred xyz, features]
            return torch.cat([grouped[:, :3], grouped], dim=1)
        centred = ops.grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return torch.cat([centred, centred], dim=1)
        grouped_features = ops.grouping_operation(features, idx)
        return torch.cat([centred, centred, grouped_features], dim=1) if self.use_xyz else grouped_features

    def conv(self, x, pos, new_pos, radius_idx, scale_idx):
        assert scale_idx < len(self.mlps)
        return self.mlps[scale_idx](self._prepare_features(pos, new_pos, x, radius_idx))


class RSConvMapper(nn.Module):
    """dense.py:18-68: the relation MLP, the optional lift of the input features (first layer) and the norm applied
    to their product"""

    def __init__(self, down_conv_nn, use_xyz, bn=True, activation=None, *args, **kwargs):
        super().__init__()
        self._down_conv_nn, self._use_xyz = down_conv_nn, use_xyz
        self._first_layer = len(down_conv_nn) == 2
        self.nn = nn.ModuleDict()
        if self._first_layer:
            f_in, f_mid, f_out = down_conv_nn[0]
            self.nn["features_nn"] = MLP2D(down_conv_nn[1], bn=bn, bias=False)
        else:
            f_in, f_mid, f_out = down_conv_nn
        self.nn["mlp_msg"] = MLP2D([f_in, f_mid, f_out], bn=bn, bias=False)
        self.nn["norm"] = nn.Sequential(nn.BatchNorm2d(f_out),
                                        activation if activation is not None else nn.LeakyReLU(negative_slope=0.01))
        self._f_out = f_out

    @property
    def f_out(self):
        return self._f_out

    def forward(self, features, msg):
        msg = self.nn["mlp_msg"](msg)
        if self._first_layer:
            features = self.nn["features_nn"](features)
        return self.nn["norm"](torch.mul(features, msg))


class SharedRSConv(nn.Module):
    """dense.py:71-104.  aggr_features [B, 3 + 3 + C, npoint, nsample], centroids [B, 3, npoint, 1]"""

    def __init__(self, mapper, radius):
        super().__init__()
        self._mapper, self._radius = mapper, radius

    def forward(self, aggr_features, centroids):
        absolute, delta = aggr_features[:, :3], aggr_features[:, 3:6]
        centre = centroids.repeat(1, 1, 1, absolute.shape[-1])
        return self._mapper(aggr_features[:, 3:], _relation(delta, centre, absolute, delta))

    def __repr__(self):
        return "{}(radius={})".format(self.__class__.__name__, self._radius)


class RSConvSharedMSGDown(_DenseRSConvDown):
    """dense.py:107-199: ONE mapper shared by all scales, max over the neighbours, then the channel-raising
    Conv1d + BatchNorm1d + activation.  The grouped coordinates are kept absolute AND centred here, so the two
    gathers stay separate launches (adding the centre back to the fused kernel's output would not be bit-equal)."""

    PER_SCALE_MAPPER = False

    def __init__(self, npoint=None, radii=None, nsample=None, down_conv_nn=None, channel_raising_nn=None, bn=True,
                 use_xyz=True, activation=None, bias=True, **kwargs):
        self._setup(npoint, radii, nsample, down_conv_nn, use_xyz, fused_grouping=True)
        activation = activation if activation is not None else nn.ReLU()
        if not self.PER_SCALE_MAPPER:
            self._mapper = RSConvMapper(down_conv_nn, activation=activation, use_xyz=use_xyz)
        self.mlp_out = nn.Sequential(
            nn.Conv1d(channel_raising_nn[0], channel_raising_nn[-1], kernel_size=1, stride=1, bias=True),
            nn.BatchNorm1d(channel_raising_nn[-1]), activation)
        for radius in radii:
            mapper = RSConvMapper(down_conv_nn, activation=activation, use_xyz=use_xyz) if self.PER_SCALE_MAPPER \
                else self._mapper
            self.mlps.append(SharedRSConv(mapper, radius))
        if self.PER_SCALE_MAPPER:
            self._mapper = mapper

    def _prepare_features(self, x, pos, new_pos, idx):
        absolute = ops.grouping_operation(pos.transpose(1, 2).contiguous(), idx)
        centroids = new_pos.transpose(1, 2).unsqueeze(-1)
        centred = absolute - centroids
        if x is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return torch.cat([absolute, centred], dim=1), centroids
        grouped_features = ops.grouping_operation(x, idx)
        if not self.use_xyz:
            return grouped_features, centroids
        return torch.cat([absolute, centred, grouped_features], dim=1), centroids

    def conv(self, x, pos, new_pos, radius_idx, scale_idx):
        assert scale_idx < len(self.mlps)
        aggr_features, centroids = self._prepare_features(x, pos, new_pos, radius_idx)
        new_features = self.mlps[scale_idx](aggr_features, centroids)
        new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        return self.mlp_out(new_features.squeeze(-1))


class RSConvMSGDown(RSConvSharedMSGDown):
    """dense.py:400-482: as RSConvSharedMSGDown with one mapper PER scale (`_mapper` names the last one, so its
    parameters appear twice in the state_dict, as in the reference)"""

    PER_SCALE_MAPPER = True


class Convolution(nn.Module):
    """modules/RSConv/message_passing.py:10-52 (torch_geometric MessagePassing, aggr="max"): the relation vector
    h_ij = [|v_ij|, v_ij, pos_i, pos_j] with v_ij = pos_i - pos_j goes through `local_nn`, weighs the source features
    x_j (the source positions when there are no features), the per-target max goes through the activation and the
    optional channel-raising `global_nn`.  Targets without edges get 0, as from torch_scatter."""

    _RELATION_ORDER = [9, 6, 7, 8, 0, 1, 2, 3, 4, 5]  # ops.edge_encoding yields [pos_i, pos_j, v_ij, |v_ij|]

    def __init__(self, local_nn, activation=None, global_nn=None, aggr="max", **kwargs):
        super().__init__()
        if aggr != "max":
            raise ValueError("only the per-target max of the reference's RSConv layer is built (aggr=%r)" % (aggr,))
        self.local_nn = MLP(local_nn)
        self.activation = activation if activation is not None else nn.ReLU()
        self.global_nn = MLP(global_nn) if global_nn is not None else None

    def forward(self, x, pos, edge_index):
        pos_src, pos_dst = (pos, pos) if torch.is_tensor(pos) else pos
        src, dst = edge_index[0], edge_index[1]
        relation = ops.edge_encoding(pos_src, pos_dst, src, dst)[:, self._RELATION_ORDER]
        x_j = pos_src[src] if x is None else x[src]
        out = self.activation(ops.segment_max(self.local_nn(relation) * x_j, dst, pos_dst.shape[0]))
        return self.global_nn(out) if self.global_nn is not None else out


class RSConvDown(BaseConvolutionDown):
    """modules/RSConv/message_passing.py:55-60: FPSSampler(ratio) -> RadiusNeighbourFinder(radius) -> Convolution.
    See BaseConvolutionDown for the order of the position pair."""

    def __init__(self, ratio=None, radius=None, local_nn=None, down_conv_nn=None, *args, **kwargs):
        super().__init__(FPSSampler(ratio), RadiusNeighbourFinder(radius), *args, **kwargs)
        self._conv = Convolution(local_nn=local_nn, global_nn=down_conv_nn)

    def conv(self, x, pos, edge_index, batch):
        return self._conv(x, pos, edge_index)


class RSConv_MSN(nn.Module):
    """The multi-scale-neighbourhood RSConv segmentation network: RSConvLogicModel (models/segmentation/rsconv.py:14-124)
    with its architecture entry `RSConv_MSN` (conf/models/segmentation/rsconv.yaml:57-108) written out.  Four
    RSConvOriginalMSGDown levels (1024 / 256 / 64 / 16 centres, three radii each), two global modules (mean over the
    points of the last and of the second-to-last level), four feature-propagation levels back to the input points
    (the fifth `up_conv` entry of the file is empty and never called: kept, parameter-free), and a head over
    [propagated features, both global vectors broadcast to every point].  Module names follow
    UnwrappedUnetBasedModel (`down_modules`, `inner_modules`, `up_modules`, `FC_layer`).

    forward(pos [B,N,3] with N >= 1024, x [B,N,FEAT]) returns the [B*N, classes] logits of rsconv.py:104; `loss` is
    its cross-entropy with the ignore label (:110-112)."""

    def __init__(self, feat_dim=3, num_classes=50, **kwargs):
        super().__init__()
        levels = [  # npoint, radii, nsample, relation MLP (+ feature lift on the first level), channel raising
            (1024, [0.075, 0.1, 0.125], [16, 32, 48], [[10, 64 // 2, 16], [feat_dim + 3, 16]], [16, 64]),
            (256, [0.1, 0.15, 0.2], [16, 48, 64], [10, 128 // 4, 64 * 3 + 3], [64 * 3 + 3, 128]),
            (64, [0.2, 0.3, 0.4], [16, 32, 48], [10, 256 // 4, 128 * 3 + 3], [128 * 3 + 3, 256]),
            (16, [0.4, 0.6, 0.8], [16, 24, 32], [10, 512 // 4, 256 * 3 + 3], [256 * 3 + 3, 512])]
        self.down_modules = nn.ModuleList(
            RSConvOriginalMSGDown(npoint=n, radii=r, nsample=k, down_conv_nn=d, channel_raising_nn=c, index=i, **kwargs)
            for i, (n, r, k, d, c) in enumerate(levels))
        self.inner_modules = nn.ModuleList([GlobalDenseBaseModule([512 * 3 + 3, 128], aggr="mean"),
                                            GlobalDenseBaseModule([256 * 3 + 3, 128], aggr="mean")])
        self.up_modules = nn.ModuleList(DenseFPModule(up_conv_nn=w, bn=True, bias=False, skip=True, index=i) for i, w in enumerate(
            [[512 * 3 + 256 * 3, 512, 512], [128 * 3 + 512, 512, 512], [64 * 3 + 512, 256, 256],
             [256 + feat_dim, 128, 128], []]))
        self.FC_layer = nn.Sequential(Conv1D(128 * 2 + 2 * 64, 128, bn=True, bias=False),
                                      Conv1D(128, num_classes, activation=None, bias=True, bn=False))
        self._num_classes = num_classes

    def forward(self, pos, x):
        from types import SimpleNamespace
        data = SimpleNamespace(pos=pos, x=None if x is None else x.transpose(1, 2).contiguous())
        skips = [data]
        for module in self.down_modules[:-1]:
            data = module(data)
            skips.append(data)
        data = self.down_modules[-1](data)
        whole = self.inner_modules[0](data)                  # mean over the 16 points of the last level
        coarse = self.inner_modules[1](skips[3])             # mean over the 64 points of the level above
        for module in self.up_modules[:-1]:
            data = module((data, skips.pop()))
        n = data.x.shape[-1]
        last = torch.cat([data.x, whole.x.repeat(1, 1, n), coarse.x.repeat(1, 1, n)], dim=1)
        return self.FC_layer(last).transpose(1, 2).contiguous().view((-1, self._num_classes))

    def loss(self, logits, labels):
        return F.cross_entropy(logits, torch.flatten(labels).long(), ignore_index=-1)
