"""Host-side mirror of the rigid KPConv layer:
  KPConv_ops     torch_points3d/modules/KPConv/convolution_ops.py:19-107
  KPConvLayer    torch_points3d/modules/KPConv/kernels.py:20-104
  SimpleBlock    torch_points3d/modules/KPConv/blocks.py:15-97 (rigid, on-device neighbour search)
Same constructor arguments; K_points stays a non-trainable Parameter so that reference
checkpoints load.
"""
import os

import numpy as np
import torch
from torch.nn.parameter import Parameter

from . import ops
from .ops import KPConv_ops  # noqa: F401  (re-export under the reference's name)
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


class SimpleBlock(torch.nn.Module):
    """blocks.py:15-97 (rigid): neighbour search (partial-dense ball query) -> KPConv -> BN ->
    LeakyReLU(0.1).  `sampler(pos, batch) -> (q_pos, q_batch)` supplies the strided query set
    (GridSampling3D in the reference, a 'next' row of SURVEY.md section 8f)."""

    RIGID_DENSITY = 2.5

    def __init__(self, down_conv_nn=None, grid_size=None, prev_grid_size=None, sigma=1.0, max_num_neighbors=16,
                 activation=None, bn_momentum=0.02, bn=True, add_one=False, sampler=None, **kwargs):
        super().__init__()
        assert len(down_conv_nn) == 2
        num_inputs, num_outputs = down_conv_nn
        self.kp_conv = KPConvLayer(num_inputs, num_outputs, point_influence=prev_grid_size * sigma,
                                   add_one=add_one, **kwargs)
        search_radius = self.RIGID_DENSITY * sigma * prev_grid_size
        self.neighbour_finder = RadiusNeighbourFinder(search_radius, max_num_neighbors, conv_type=PARTIAL_DENSE)
        self.bn = torch.nn.BatchNorm1d(num_outputs, momentum=bn_momentum) if bn else None
        self.activation = activation if activation is not None else torch.nn.LeakyReLU(negative_slope=0.1)
        self.sampler = sampler if prev_grid_size != grid_size else None

    def forward(self, pos, x, batch):
        if self.sampler is not None:
            q_pos, q_batch = self.sampler(pos, batch)
        else:
            q_pos, q_batch = pos, batch
        idx_neighboors = self.neighbour_finder(pos, q_pos, batch_x=batch, batch_y=q_batch)
        out = self.kp_conv(q_pos, pos, idx_neighboors, x)
        if self.bn is not None:
            out = self.bn(out)
        return q_pos, self.activation(out), q_batch, idx_neighboors
