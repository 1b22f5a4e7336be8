"""`from torch_points3d_b200.compat.cluster import fps, knn, radius, knn_interpolate` -- the
torch_geometric.nn (torch_cluster 1.5.9) functions the reference calls on this path."""
from ..ops import fps, knn_interpolate, radius  # noqa: F401
from ..ops import knn_edges as knn  # noqa: F401

__all__ = ["fps", "knn", "radius", "knn_interpolate"]
