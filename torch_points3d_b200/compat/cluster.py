"""`from torch_points3d_b200.compat.cluster import fps, knn, radius, knn_interpolate, global_max_pool, global_mean_pool,
PointConv` -- the torch_geometric.nn (torch_cluster 1.5.9) names the reference imports on this path
(core/base_conv/message_passing.py:10-17, core/base_conv/partial_dense.py:10-17, core/spatial_ops/*.py,
modules/pointnet2/message_passing.py:1), served by libp3d_b200.so."""
from ..ops import fps, knn_interpolate, radius  # noqa: F401
from ..ops import knn_edges as knn  # noqa: F401


def _pool(x, batch, size, aggr):
    from ..partial_dense import _global_pool
    if size is None:
        size = int(batch[-1].item()) + 1 if batch.numel() else 0   # the library's own `batch.max() + 1` read
    return _global_pool(x, batch, int(size), aggr)


def global_max_pool(x, batch, size=None):
    """torch_geometric.nn.global_max_pool over a sorted batch vector: [N, C] -> [size, C]"""
    return _pool(x, batch, size, "max")


def global_mean_pool(x, batch, size=None):
    """torch_geometric.nn.global_mean_pool over a sorted batch vector: [N, C] -> [size, C]"""
    return _pool(x, batch, size, "mean")


def __getattr__(name):
    if name == "PointConv":             # imported on demand: message_passing imports spatial_ops, which may import this
        from ..message_passing import PointConv
        return PointConv
    raise AttributeError(name)


__all__ = ["fps", "knn", "radius", "knn_interpolate", "global_max_pool", "global_mean_pool", "PointConv"]
