"""Drop-in namespaces with the signatures of the third-party wheels the reference imports:
`compat.tp` for `torch_points_kernels`, `compat.cluster` for `torch_geometric.nn` / `torch_cluster`."""
from . import cluster, tp  # noqa: F401
