"""Host-side mirror of the rigid KPConv layer:
  KPConv_ops     torch_points3d/modules/KPConv/convolution_ops.py:19-107
  KPConvLayer    torch_points3d/modules/KPConv/kernels.py:20-104
  SimpleBlock    torch_points3d/modules/KPConv/blocks.py:15-97 (on-device neighbour search)
  ResnetBBlock   blocks.py:100-225        KPDualBlock  blocks.py:228-292
  KPConvPaper    torch_points3d/models/segmentation/kpconv.py:22-140 with conf/models/segmentation/kpconv.yaml:4-76
Same constructor arguments and parameter names (K_points stays a non-trainable Parameter, BatchNorm sits inside
FastBatchNorm1d as `.batch_norm`), so a reference state_dict loads key for key.
"""
import os

import numpy as np
import torch
from torch.nn.parameter import Parameter

from . import ops
from .ops import KPConv_ops  # noqa: F401  (re-export under the reference's name)
from .partial_dense import FastBatchNorm1d, FPModule_PD
from .spatial_ops import PARTIAL_DENSE, RadiusNeighbourFinder

# 15 kernel points of k_015_center.ply (unit radius): DATA of the reference model, stored next to
# the code the way modules/KPConv/kernels/dispositions/ does.
_DISPOSITION = os.path.join(os.path.dirname(os.path.abspath(__file__)), "k_015_center.npy")


def load_kernels(radius, num_kpoints=15, fixed="center", random_rotation=True, rng=None):
    """kernel_utils.load_kernels (modules/KPConv/kernel_utils.py:189-282) for the shipped
    15-point 'center' disposition: scale by `radius`, random rotation, N(0, 0.01 r) jitter."""
    if num_kpoints != 15 or fixed != "center":
        raise NotImplementedError("only the k_015_center disposition is shipped")
    kernel = np.load(_DISPOSITION).astype(np.float64) * radius
    if not random_rotation:
        return kernel.astype(np.float32)
    rng = np.random.default_rng() if rng is None else rng
    u = rng.uniform(-1, 1, 3)
    u /= np.linalg.norm(u) + 1e-9
    v = rng.uniform(-1, 1, 3)
    v -= np.dot(u, v) * u
    v /= np.linalg.norm(v) + 1e-9
    R = np.stack((u, v, np.cross(u, v)), axis=-1)
    kernel = kernel @ R + rng.normal(scale=radius * 0.01, size=kernel.shape)
    return kernel.astype(np.float32)


def add_ones(query_points, x, add_one):
    """kernels.py:10-17"""
    if add_one:
        ones = torch.ones(query_points.shape[0], dtype=torch.float, device=query_points.device).unsqueeze(-1)
        x = torch.cat([ones.to(x.dtype), x], dim=-1) if x is not None else ones
    return x


class KPConvLayer(torch.nn.Module):
    """kernels.py:20-104"""

    _INFLUENCE_TO_RADIUS = 1.5

    def __init__(self, num_inputs, num_outputs, point_influence, n_kernel_points=15, fixed="center",
                 KP_influence="linear", aggregation_mode="sum", dimension=3, add_one=False,
                 random_rotation=True, **kwargs):
        super().__init__()
        self.kernel_radius = self._INFLUENCE_TO_RADIUS * point_influence
        self.point_influence = point_influence
        self.add_one = add_one
        self.num_inputs = num_inputs + self.add_one * 1
        self.num_outputs = num_outputs
        self.KP_influence = KP_influence
        self.n_kernel_points = n_kernel_points
        self.aggregation_mode = aggregation_mode
        K_points_numpy = load_kernels(self.kernel_radius, n_kernel_points, fixed=fixed,
                                      random_rotation=random_rotation)
        self.K_points = Parameter(torch.from_numpy(K_points_numpy.reshape((n_kernel_points, dimension))).float(),
                                  requires_grad=False)
        weights = torch.empty([n_kernel_points, self.num_inputs, num_outputs], dtype=torch.float)
        torch.nn.init.xavier_normal_(weights)
        self.weight = Parameter(weights)

    def forward(self, query_points, support_points, neighbors, x):
        x = add_ones(support_points, x, self.add_one)
        return ops.KPConv_ops(query_points, support_points, neighbors, x, self.K_points, self.weight,
                              self.point_influence, self.KP_influence, self.aggregation_mode)

    def __repr__(self):
        return "KPConvLayer(InF: %i, OutF: %i, kernel_pts: %i, radius: %.2f, KP_influence: %s, Add_one: %s)" % (
            self.num_inputs, self.num_outputs, self.n_kernel_points, self.kernel_radius, self.KP_influence,
            self.add_one)


def fitting_loss(sq_distance, radius):
    """losses.py:4-16"""
    kpmin = sq_distance.min(dim=1)[0]
    return torch.mean(kpmin / (radius ** 2))


def repulsion_loss(deformed_kpoints, radius):
    """losses.py:19-36 (the division by the radius in the reference discards its result, :26: kept that way)"""
    n_points = deformed_kpoints.shape[1]
    repulsive_loss = 0
    for i in range(n_points):
        with torch.no_grad():
            other_points = torch.cat([deformed_kpoints[:, :i, :], deformed_kpoints[:, i + 1:, :]], dim=1)
        distances = torch.sqrt(torch.sum((other_points - deformed_kpoints[:, i:i + 1, :]) ** 2, dim=-1))
        repulsion_force = torch.sum(torch.pow(torch.relu(1.5 - distances), 2), dim=1)
        repulsive_loss += torch.mean(repulsion_force)
    return repulsive_loss


def permissive_loss(deformed_kpoints, radius):
    """losses.py:39-45"""
    norm_deformed_normalized = torch.norm(deformed_kpoints, p=2, dim=-1) / float(radius)
    return torch.mean(norm_deformed_normalized[norm_deformed_normalized > 1.0])


class KPConvDeformableLayer(torch.nn.Module):
    """modules/KPConv/kernels.py:107-256: a rigid KPConv (offset_weights) predicts per-point kernel offsets (and
    modulations), the deformable operator applies them; fitting / repulsion (or permissive) regularisers are kept
    in `internal_losses`.  The fitting loss reads the fused min over neighbours instead of the [n,M,K] tensor."""

    PERMISSIVE_LOSS_KEY = "permissive_loss"
    FITTING_LOSS_KEY = "fitting_loss"
    REPULSION_LOSS_KEY = "repulsion_loss"
    _INFLUENCE_TO_RADIUS = 1.5

    def __init__(self, num_inputs, num_outputs, point_influence, n_kernel_points=15, fixed="center",
                 KP_influence="linear", aggregation_mode="sum", dimension=3, modulated=False, loss_mode="fitting",
                 add_one=False, random_rotation=True, **kwargs):
        super().__init__()
        assert dimension == 3
        self.kernel_radius = self._INFLUENCE_TO_RADIUS * point_influence
        self.point_influence = point_influence
        self.add_one = add_one
        self.num_inputs = num_inputs + self.add_one * 1
        self.num_outputs = num_outputs
        self.KP_influence, self.n_kernel_points, self.aggregation_mode = KP_influence, n_kernel_points, aggregation_mode
        self.modulated, self.loss_mode = modulated, loss_mode
        self.internal_losses = {self.PERMISSIVE_LOSS_KEY: 0.0, self.FITTING_LOSS_KEY: 0.0, self.REPULSION_LOSS_KEY: 0.0}
        kp = load_kernels(self.kernel_radius, n_kernel_points, fixed=fixed, random_rotation=random_rotation)
        self.K_points = torch.nn.Parameter(torch.from_numpy(kp.reshape(n_kernel_points, 3)).to(torch.float),
                                           requires_grad=False)
        offset_dim = (4 if modulated else 3) * n_kernel_points
        offset_weights = torch.empty([n_kernel_points, self.num_inputs, offset_dim], dtype=torch.float)
        torch.nn.init.xavier_normal_(offset_weights)
        self.offset_weights = torch.nn.Parameter(offset_weights)
        self.offset_bias = torch.nn.Parameter(torch.zeros(offset_dim, dtype=torch.float))
        weights = torch.empty([n_kernel_points, self.num_inputs, num_outputs], dtype=torch.float)
        torch.nn.init.xavier_normal_(weights)
        self.weight = torch.nn.Parameter(weights)

    def forward(self, query_points, support_points, neighbors, x):
        x = add_ones(support_points, x, self.add_one)
        offset_feat = ops.KPConv_ops(query_points, support_points, neighbors, x, self.K_points, self.offset_weights,
                                     self.point_influence, self.KP_influence, self.aggregation_mode) + self.offset_bias
        nk = self.n_kernel_points
        if self.modulated:
            offsets = offset_feat[:, :3 * nk].reshape((-1, nk, 3))
            modulations = 2 * torch.sigmoid(offset_feat[:, 3 * nk:])
        else:
            offsets = offset_feat.reshape((-1, nk, 3))
            modulations = None
        offsets = offsets * self.point_influence
        new_feat, kpmin, K_points_deformed = ops.KPConv_deform_fitting(
            query_points, support_points, neighbors, x, self.K_points, offsets, modulations, self.weight,
            self.point_influence, self.KP_influence, self.aggregation_mode)
        if self.loss_mode == "fitting":
            self.internal_losses[self.FITTING_LOSS_KEY] = torch.mean(kpmin / (self.kernel_radius ** 2))
            self.internal_losses[self.REPULSION_LOSS_KEY] = repulsion_loss(K_points_deformed, self.point_influence)
        elif self.loss_mode == "permissive":
            self.internal_losses[self.PERMISSIVE_LOSS_KEY] = permissive_loss(K_points_deformed, self.kernel_radius)
        else:
            raise NotImplementedError("Loss mode %s not recognised. Only permissive and fitting are valid"
                                      % self.loss_mode)
        return new_feat

    def get_internal_losses(self):
        return self.internal_losses

    def __repr__(self):
        return "KPConvDeformableLayer(InF: %i, OutF: %i, kernel_pts: %i, radius: %.2f, KP_influence: %s)" % (
            self.num_inputs, self.num_outputs, self.n_kernel_points, self.kernel_radius, self.KP_influence)


class Cloud:
    """minimal stand-in for the torch_geometric `Data` objects the reference blocks pass around: an attribute bag
    (`pos`, `x`, `batch`, `idx_neighboors`, `block_idx`, `up_idx`, ...) whose `clone` copies the attribute table"""

    def __init__(self, pos=None, x=None, batch=None, idx_neighboors=None, nb_key=None, block_idx=0, **extra):
        self.pos, self.x, self.batch, self.idx_neighboors = pos, x, batch, idx_neighboors
        self.nb_key = nb_key  # (radius, max_num_neighbors) of a self-neighbour table of `pos`, if one is attached
        self.block_idx = block_idx  # position in the precomputed multiscale list (blocks.py:67-68,93)
        self.__dict__.update(extra)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def clone(self):
        out = Cloud()
        out.__dict__.update(self.__dict__)
        return out


def grid_subsample(pos, batch, size):
    """GridSampling3D(size, mode="mean") restricted to what the strided blocks read back (blocks.py:74-87:
    `query_data.pos`, `query_data.batch`): one launch of the voxel-clustering kernel (csrc/grid.cu).  Voxels come out
    in the reference's order, increasing (batch, z, y, x) of the rounded coordinates; positions are fp32 means summed
    in point order.  Returns (q_pos [n,3], q_batch [n])."""
    vox = ops.grid_cluster(pos, size, batch, want_mean=True)
    return vox.mean_pos, vox.batch


class GridSampler:
    def __init__(self, size):
        self._size = size

    def __call__(self, pos, batch):
        return grid_subsample(pos, batch, self._size)

    def __repr__(self):
        return "GridSampler(%g)" % self._size


class SimpleBlock(torch.nn.Module):
    """blocks.py:15-97: [grid subsampling of the queries when strided] -> neighbour search (partial-dense ball query,
    radius 2.5 (rigid) / 5.0 (deformable) * sigma * prev_grid_size) -> KPConv -> BN -> LeakyReLU(0.1)."""

    DEFORMABLE_DENSITY = 5.0
    RIGID_DENSITY = 2.5

    def __init__(self, down_conv_nn=None, grid_size=None, prev_grid_size=None, sigma=1.0, max_num_neighbors=16,
                 activation=None, bn_momentum=0.02, bn=FastBatchNorm1d, deformable=False, add_one=False, **kwargs):
        super().__init__()
        assert len(down_conv_nn) == 2
        num_inputs, num_outputs = down_conv_nn
        if deformable:
            density_parameter = self.DEFORMABLE_DENSITY
            self.kp_conv = KPConvDeformableLayer(num_inputs, num_outputs, point_influence=prev_grid_size * sigma,
                                                 add_one=add_one, **kwargs)
        else:
            density_parameter = self.RIGID_DENSITY
            self.kp_conv = KPConvLayer(num_inputs, num_outputs, point_influence=prev_grid_size * sigma,
                                       add_one=add_one, **kwargs)
        search_radius = density_parameter * sigma * prev_grid_size
        self.neighbour_finder = RadiusNeighbourFinder(search_radius, max_num_neighbors, conv_type=PARTIAL_DENSE)
        if bn is True:
            bn = FastBatchNorm1d
        self.bn = bn(num_outputs, momentum=bn_momentum) if bn else None
        self.activation = activation if activation is not None else torch.nn.LeakyReLU(negative_slope=0.1)
        self.sampler = GridSampler(grid_size) if prev_grid_size != grid_size else None

    def forward(self, data, precomputed=None, **kwargs):
        if precomputed:  # blocks.py:70-79: geometry of this block was built ahead of time (MultiScaleTransform)
            query = precomputed[data.block_idx]
            q_pos, q_batch, idx_neighboors = query.pos, getattr(query, "batch", None), query.idx_neighboors
            key = getattr(query, "nb_key", None)  # absent on a reference-format pyramid (multiscale.py)
        else:
            if self.sampler is not None:
                q_pos, q_batch = self.sampler(data.pos, data.batch)
            else:
                q_pos, q_batch = data.pos, data.batch
            key = (self.neighbour_finder._radius, self.neighbour_finder._max_num_neighbors)
            if self.sampler is None and data.idx_neighboors is not None and data.nb_key == key:
                idx_neighboors = data.idx_neighboors  # same points, radius and width as the previous block: same table
            else:
                idx_neighboors = self.neighbour_finder(data.pos, q_pos, batch_x=data.batch, batch_y=q_batch)
            if self.sampler is not None:
                key = None  # the table describes self-neighbourhoods of q_pos only when this block was not strided
        x = self.kp_conv(q_pos, data.pos, idx_neighboors, data.x)
        if self.bn is not None:
            x = self.bn(x)
        return Cloud(q_pos, self.activation(x), q_batch, idx_neighboors, key, data.block_idx + 1)


class ResnetBBlock(torch.nn.Module):
    """blocks.py:100-214: unary_1 -> SimpleBlock(KPConv d_2 -> d_2) -> unary_2, plus the shortcut
    (max over the neighbour rows when strided, a Linear+BN when the width changes)."""

    def __init__(self, down_conv_nn=None, grid_size=None, prev_grid_size=None, sigma=1, max_num_neighbors=16,
                 activation=None, has_bottleneck=True, bn_momentum=0.02, bn=FastBatchNorm1d, deformable=False,
                 add_one=False, **kwargs):
        super().__init__()
        if bn is True:
            bn = FastBatchNorm1d
        assert len(down_conv_nn) == 2 or len(down_conv_nn) == 3, "down_conv_nn should be of size 2 or 3"
        if len(down_conv_nn) == 2:
            num_inputs, num_outputs = down_conv_nn
            d_2 = num_outputs // 4
        else:
            num_inputs, d_2, num_outputs = down_conv_nn
        act = activation if activation is not None else torch.nn.LeakyReLU(negative_slope=0.1)
        self.is_strided = prev_grid_size != grid_size
        self.has_bottleneck = has_bottleneck
        kp_size = [d_2, d_2] if has_bottleneck else [num_inputs, num_outputs]
        self.kp_conv = SimpleBlock(down_conv_nn=kp_size, grid_size=grid_size, prev_grid_size=prev_grid_size,
                                   sigma=sigma, max_num_neighbors=max_num_neighbors, activation=act,
                                   bn_momentum=bn_momentum, bn=bn, deformable=deformable, add_one=add_one, **kwargs)
        Lin, BN = ops.Linear, bn
        if has_bottleneck:
            if bn:
                self.unary_1 = torch.nn.Sequential(Lin(num_inputs, d_2, bias=False), BN(d_2, momentum=bn_momentum), act)
                self.unary_2 = torch.nn.Sequential(Lin(d_2, num_outputs, bias=False),
                                                   BN(num_outputs, momentum=bn_momentum), act)
            else:
                self.unary_1 = torch.nn.Sequential(Lin(num_inputs, d_2, bias=False), act)
                self.unary_2 = torch.nn.Sequential(Lin(d_2, num_outputs, bias=False), act)
        if num_inputs != num_outputs:
            self.shortcut_op = (torch.nn.Sequential(Lin(num_inputs, num_outputs, bias=False),
                                                    BN(num_outputs, momentum=bn_momentum))
                                if bn else Lin(num_inputs, num_outputs, bias=False))
        else:
            self.shortcut_op = torch.nn.Identity()
        self.activation = act

    def forward(self, data, precomputed=None, **kwargs):
        shortcut_x = data.x
        inner = data.clone()
        if self.has_bottleneck:
            inner.x = self.unary_1(inner.x)
        output = self.kp_conv(inner, precomputed=precomputed)
        if self.has_bottleneck:
            output.x = self.unary_2(output.x)
        if self.is_strided:  # max over the neighbour rows of cat(x, zeros): one gather+max kernel
            shortcut_x = ops.neighbor_max(shortcut_x, output.idx_neighboors)
        output.x = output.x + self.shortcut_op(shortcut_x)
        return output

    @property
    def sampler(self):
        return self.kp_conv.sampler

    @property
    def neighbour_finder(self):
        return self.kp_conv.neighbour_finder


class KPConvEncoder(torch.nn.Module):
    """The down-conv stack of KPConvPaper (conf/models/segmentation/kpconv.yaml:14-57): five KPDualBlocks,
    [SimpleBlock, ResnetBBlock] then four [strided ResnetBBlock, ResnetBBlock], widths in_feat * 2^level,
    grids first_subsampling * 2^level, max_num_neighbors from the YAML (clipped to `max_nn_cap`)."""

    MAX_NN = [[25, 25], [25, 30], [30, 38], [38, 38], [38, 38]]

    def __init__(self, feat_dim, in_feat=64, first_subsampling=0.04, levels=5, max_nn_cap=None, **kwargs):
        super().__init__()
        g, f = first_subsampling, in_feat
        self.blocks = torch.nn.ModuleList()
        for lv in range(levels):
            nn_ = [min(v, max_nn_cap) if max_nn_cap else v for v in self.MAX_NN[lv]]
            if lv == 0:
                self.blocks.append(SimpleBlock([feat_dim, f], grid_size=g, prev_grid_size=g, max_num_neighbors=nn_[0],
                                               **kwargs))
                self.blocks.append(ResnetBBlock([f, 2 * f], grid_size=g, prev_grid_size=g, max_num_neighbors=nn_[1],
                                                **kwargs))
            else:
                w = f * 2 ** lv
                self.blocks.append(ResnetBBlock([w, w], grid_size=g * 2 ** lv, prev_grid_size=g * 2 ** (lv - 1),
                                                max_num_neighbors=nn_[0], **kwargs))
                self.blocks.append(ResnetBBlock([w, 2 * w], grid_size=g * 2 ** lv, prev_grid_size=g * 2 ** lv,
                                                max_num_neighbors=nn_[1], **kwargs))
        self.out_channels = in_feat * 2 ** levels

    def forward(self, pos, x, batch, precomputed=None):
        data = Cloud(pos, x, batch)
        for block in self.blocks:
            data = block(data, precomputed=precomputed)
        return data

    def strategies(self):
        """samplers and neighbour finders of the blocks in order -- what the reference hands to MultiScaleTransform
        (models/base_architectures/unet.py collects the same lists from its down modules)"""
        return {"sampler": [b.sampler for b in self.blocks], "neighbour_finder": [b.neighbour_finder for b in self.blocks]}


def _is_list(v):
    return isinstance(v, (list, tuple))


class KPDualBlock(torch.nn.Module):
    """blocks.py:228-292: a pair of KPConv blocks (usually strided + non strided) built by name."""

    def __init__(self, block_names=None, down_conv_nn=None, grid_size=None, prev_grid_size=None, has_bottleneck=None,
                 max_num_neighbors=None, deformable=False, add_one=False, **kwargs):
        super().__init__()
        assert len(block_names) == len(down_conv_nn)
        self.blocks = torch.nn.ModuleList()
        classes = {"SimpleBlock": SimpleBlock, "ResnetBBlock": ResnetBBlock}
        for i, class_name in enumerate(block_names):
            block_kwargs = {key: (arg[i] if _is_list(arg) else arg) for key, arg in kwargs.items()}
            self.blocks.append(classes[class_name](
                down_conv_nn=down_conv_nn[i], grid_size=grid_size[i], prev_grid_size=prev_grid_size[i],
                has_bottleneck=has_bottleneck[i], max_num_neighbors=max_num_neighbors[i],
                deformable=deformable[i] if _is_list(deformable) else deformable,
                add_one=add_one[i] if _is_list(add_one) else add_one, **block_kwargs))

    def forward(self, data, precomputed=None, **kwargs):
        for block in self.blocks:
            data = block(data, precomputed=precomputed)
        return data

    @property
    def sampler(self):
        return [b.sampler for b in self.blocks]

    @property
    def neighbour_finder(self):
        return [b.neighbour_finder for b in self.blocks]


class KPConvPaper(torch.nn.Module):
    """The KPConv segmentation network (models/segmentation/kpconv.py:22-140) with its architecture file
    (conf/models/segmentation/kpconv.yaml:4-76) resolved: five KPDualBlocks down (the yaml's down_conv_nn /
    grid_size / prev_grid_size / block_names / has_bottleneck / max_num_neighbors lists), no innermost module, four
    FPModule_PD up (KNNInterpolate(k = 1) + skip + MLP, core/base_conv/partial_dense.py:103-146), then the mlp_cls
    head (Linear -> FastBatchNorm1d -> LeakyReLU(0.2), Linear to the classes, LogSoftmax).  Module names follow
    UnwrappedUnetBasedModel (`down_modules`, `up_modules`, `FC_layer`).

    forward(pos, x, batch, multiscale=None, upsample=None) mirrors set_input + forward: a ones column is prepended to
    the features (kpconv.py:88), the down modules push their outputs on a stack, the up modules pop them.
    `multiscale` / `upsample` are the precomputed pyramid halves (MultiScaleTransform); without them every block
    searches on the fly."""

    MAX_NN = [[25, 25], [25, 30], [30, 38], [38, 38], [38, 38]]

    def __init__(self, feat_dim, num_classes, in_feat=64, first_subsampling=0.04, bn_momentum=0.02, levels=5,
                 max_nn_cap=None, deformable=False, **kwargs):
        super().__init__()
        g, f = first_subsampling, in_feat
        self.down_modules = torch.nn.ModuleList()
        for lv in range(levels):
            nn_ = [min(v, max_nn_cap) if max_nn_cap else v for v in self.MAX_NN[lv]]
            w = f * 2 ** lv
            if lv == 0:
                args = dict(block_names=["SimpleBlock", "ResnetBBlock"], down_conv_nn=[[feat_dim + 1, f], [f, 2 * f]],
                            grid_size=[g, g], prev_grid_size=[g, g], has_bottleneck=[False, True])
            else:
                args = dict(block_names=["ResnetBBlock", "ResnetBBlock"], down_conv_nn=[[w, w], [w, 2 * w]],
                            grid_size=[g * 2 ** lv, g * 2 ** lv], prev_grid_size=[g * 2 ** (lv - 1), g * 2 ** lv],
                            has_bottleneck=[True, True])
            self.down_modules.append(KPDualBlock(max_num_neighbors=nn_, deformable=deformable, n_kernel_points=15,
                                                 bn_momentum=bn_momentum, **args, **kwargs))
        self.up_modules = torch.nn.ModuleList()
        for lv in range(levels - 1, 0, -1):  # yaml up_conv_nn: [32f+16f, 8f], [8f+8f, 4f], [4f+4f, 2f], [2f+2f, f]
            below = f * 2 ** (lv + 1) if lv == levels - 1 else f * 2 ** lv
            skip = f * 2 ** lv
            self.up_modules.append(FPModule_PD(1, [below + skip, f * 2 ** (lv - 1)], bn_momentum=bn_momentum))
        self.FC_layer = torch.nn.Sequential()
        self.FC_layer.add_module("1", torch.nn.Sequential(ops.Linear(f, f, bias=False),
                                                          FastBatchNorm1d(f, momentum=bn_momentum),
                                                          torch.nn.LeakyReLU(0.2)))
        self.FC_layer.add_module("Class", torch.nn.Linear(f, num_classes, bias=False))
        self.FC_layer.add_module("Softmax", torch.nn.LogSoftmax(-1))

    def forward(self, pos, x, batch, multiscale=None, upsample=None):
        data = Cloud(pos, add_ones(pos, x, True), batch)
        stack_down = []
        for i in range(len(self.down_modules) - 1):
            data = self.down_modules[i](data, precomputed=multiscale)
            stack_down.append(data)
        data = self.down_modules[-1](data, precomputed=multiscale)
        for i in range(len(self.up_modules)):
            # no innermost module in this architecture: every level is an interpolation (has_innermost = False is
            # what the reference's `len(data.x) == data.batch.max() + 1` evaluates to, without its device read)
            data = self.up_modules[i]((data, stack_down.pop()), precomputed=upsample, has_innermost=False)
        return self.FC_layer(data.x)

    def strategies(self):
        """what the reference's model hands to MultiScaleTransform (models/base_architectures/unet.py
        `_spatial_ops_dict`): samplers and neighbour finders of every down block in order, upsample ops innermost first"""
        samplers, finders = [], []
        for m in self.down_modules:
            samplers += m.sampler
            finders += m.neighbour_finder
        return {"sampler": samplers, "neighbour_finder": finders,
                "upsample_op": [m.upsample_op for m in self.up_modules]}


class MultiScaleTransform:
    """core/data_transform/transforms.py:579-651 on the device: runs every block's sampler and neighbour finder once,
    ahead of the network, and returns the `multiscale` list the blocks read through `precomputed=`
    (modules/KPConv/blocks.py:70-79).  Differences from the reference's DataLoader-worker version: the whole batch is
    processed at once by the batched operators (no per-sample loop, so MultiScaleBatch.from_data_list's index
    shifting -- and its aliasing of every sample's shadow index with the first point of the next sample,
    datasets/multiscale_data.py:128-136 -- does not arise: the shadow index is the support count of the batch), and a
    block that repeats the previous block's search on unchanged points shares its table."""

    def __init__(self, strategies):
        self.strategies = strategies  # {"sampler": [...], "neighbour_finder": [...], optional "upsample_op": [...]}
        self.num_layers = len(strategies["sampler"])

    @torch.no_grad()
    def upsample(self, pos, batch, multiscale):
        """the `upsample` half of the pyramid (transforms.py:616-629, 649-650): one KNNInterpolate.precompute per
        strided level -- query = the subsampled level, support = the level above it -- innermost level first"""
        ups = self.strategies.get("upsample_op") or []
        out, k = [], 0
        support = Cloud(pos, None, batch)
        for sampler, level in zip(self.strategies["sampler"], multiscale):
            if sampler is not None and ups:
                if k >= len(ups):
                    raise ValueError("You are missing some upsample blocks in your network")
                out.append(ups[k].precompute(level, support))
                k += 1
            support = level
        out.reverse()
        return out

    @torch.no_grad()
    def __call__(self, pos, batch):
        # pass 1: every level's query set.  Each grid sampling reads its voxel count back (it sizes the outputs), so
        # the samplers run first, back to back, while the stream is nearly empty; the neighbour searches -- the bulk
        # of the work -- are queued afterwards and never sit in front of one of those reads.
        levels = []
        s_pos, s_batch = pos, batch
        for sampler in self.strategies["sampler"]:
            q_pos, q_batch = sampler(s_pos, s_batch) if sampler is not None else (s_pos, s_batch)
            levels.append((s_pos, s_batch, q_pos, q_batch))
            s_pos, s_batch = q_pos, q_batch
        # pass 2: neighbour tables
        multiscale = []
        prev_key, prev_table = None, None
        for (s_pos, s_batch, q_pos, q_batch), sampler, finder in zip(levels, self.strategies["sampler"],
                                                                    self.strategies["neighbour_finder"]):
            key = (finder._radius, finder._max_num_neighbors)
            if sampler is None and prev_key == key:
                table = prev_table
            else:
                table = finder(s_pos, q_pos, batch_x=s_batch, batch_y=q_batch)
            multiscale.append(Cloud(q_pos, None, q_batch, table, key if sampler is None else None))
            prev_key, prev_table = (key, table) if sampler is None else (None, None)
        return multiscale
