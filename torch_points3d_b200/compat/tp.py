"""`import torch_points3d_b200.compat.tp as tp` -- the torch_points_kernels 0.7.0 surface the
reference calls (7 call sites + 6 in RSConv), served by libp3d_b200.so."""
from ..ops import ball_query, furthest_point_sample, grouping_operation, three_interpolate, three_nn  # noqa: F401
from ..ops import knn_dense as knn  # noqa: F401

__all__ = ["furthest_point_sample", "ball_query", "grouping_operation", "three_nn", "three_interpolate", "knn"]
