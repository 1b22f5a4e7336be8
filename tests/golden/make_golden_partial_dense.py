"""Run tests/golden/partial_dense_cases.py on the REFERENCE's own core/base_conv/partial_dense.py and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): core/base_conv/{base_conv,partial_dense}.py,
core/common_modules/base_modules.py, core/spatial_ops/*.py, utils/enums.py.  torch_geometric / torch_scatter /
torch_points_kernels are absent from this image: `knn`, `knn_interpolate`, `global_*_pool`, `scatter_add` are replaced
by the torch stand-ins of partial_dense_cases (their published definitions), `Batch` / `Data` by an attribute bag.
/root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_partial_dense.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import partial_dense_cases as cases  # noqa: E402


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class Bag:
    """the slice of torch_geometric's Data / Batch the reference modules touch"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    @property
    def num_nodes(self):
        return self.__dict__["num_nodes"] if "num_nodes" in self.__dict__ else self.pos.shape[0]

    def clone(self):
        return Bag(**self.__dict__)


def main():
    def unused(*a, **kw):
        raise AssertionError("not on the tested paths")

    _module("torch_geometric")
    _module("torch_geometric.nn", fps=unused, knn=cases.knn, radius=unused, voxel_grid=unused,
            knn_interpolate=cases.knn_interpolate, global_max_pool=cases.global_max_pool,
            global_mean_pool=cases.global_mean_pool)
    _module("torch_geometric.nn.pool")
    _module("torch_geometric.nn.pool.consecutive", consecutive_cluster=unused)
    _module("torch_geometric.nn.pool.pool", pool_pos=unused, pool_batch=unused)
    _module("torch_geometric.data", Data=Bag, Batch=Bag)
    _module("torch_scatter", scatter_add=cases.scatter_add)
    _module("torch_points_kernels", furthest_point_sample=unused, ball_query=unused)
    for pkg in ("torch_points3d", "torch_points3d.utils", "torch_points3d.core"):
        _module(pkg)
    _module("torch_points3d.utils.config", is_list=lambda e: isinstance(e, list))
    _load("torch_points3d.utils.enums", os.path.join(REF, "utils/enums.py"))
    _load("torch_points3d.utils.debugging_vars", os.path.join(REF, "utils/debugging_vars.py"))
    base = _load("torch_points3d.core.common_modules.base_modules", os.path.join(REF, "core/common_modules/base_modules.py"))
    _module("torch_points3d.core.common_modules", **{k: v for k, v in vars(base).items() if not k.startswith("_")})
    sys.modules["torch_points3d.core.common_modules.base_modules"] = base
    so = _module("torch_points3d.core.spatial_ops")
    for f in ("sampling", "neighbour_finder", "interpolate"):
        mod = _load("torch_points3d.core.spatial_ops." + f, os.path.join(REF, "core/spatial_ops", f + ".py"))
        so.__dict__.update({k: v for k, v in vars(mod).items() if not k.startswith("_")})
    _module("torch_points3d.core.base_conv")
    _load("torch_points3d.core.base_conv.base_conv", os.path.join(REF, "core/base_conv/base_conv.py"))
    ref = _load("torch_points3d.core.base_conv.partial_dense", os.path.join(REF, "core/base_conv/partial_dense.py"))
    ns = types.SimpleNamespace(BasePartialDenseConvolutionDown=ref.BasePartialDenseConvolutionDown,
                               GlobalPartialDenseBaseModule=ref.GlobalPartialDenseBaseModule, FPModule_PD=ref.FPModule_PD)
    out = cases.run_cases(ns, Bag)
    np.savez_compressed(os.path.join(HERE, "partial_dense_chain.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
