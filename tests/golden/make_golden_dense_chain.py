"""Run tests/golden/dense_chain_cases.py on the REFERENCE's own dense PointNet++ modules and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): core/base_conv/dense.py, modules/pointnet2/dense.py,
core/base_conv/base_conv.py, core/common_modules/{base_modules,dense_modules}.py, core/spatial_ops/*.py,
utils/enums.py, utils/model_building_utils/activation_resolver.py.  The third-party modules they import are absent
from this image: `torch_points_kernels` is replaced by dense_chain_cases.Kernels (oracle restatements),
torch_geometric / torch_scatter by minimal stand-ins (`Data` is an attribute bag; nothing else is reached).
/root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_dense_chain.py
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
import dense_chain_cases  # noqa: E402


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []  # a package, so that submodules resolve through sys.modules
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def load_reference(kernels):
    def unused(*a, **kw):
        raise AssertionError("not on the tested paths")

    _module("torch_geometric")
    _module("torch_geometric.nn", fps=unused, knn=unused, radius=unused, voxel_grid=unused, knn_interpolate=unused,
            global_max_pool=unused, global_mean_pool=unused)
    _module("torch_geometric.nn.pool")
    _module("torch_geometric.nn.pool.consecutive", consecutive_cluster=unused)
    _module("torch_geometric.nn.pool.pool", pool_pos=unused, pool_batch=unused)
    _module("torch_geometric.data", Data=Data)
    _module("torch_scatter", scatter_add=unused)
    _module("torch_points_kernels", furthest_point_sample=kernels.furthest_point_sample, ball_query=kernels.ball_query,
            grouping_operation=kernels.grouping_operation, three_nn=kernels.three_nn,
            three_interpolate=kernels.three_interpolate)
    for pkg in ("torch_points3d", "torch_points3d.utils", "torch_points3d.core", "torch_points3d.core.common_modules",
                "torch_points3d.utils.model_building_utils", "torch_points3d.modules", "torch_points3d.modules.pointnet2"):
        _module(pkg)
    _module("torch_points3d.utils.config", is_list=lambda e: isinstance(e, list), is_dict=lambda e: isinstance(e, dict))
    _load("torch_points3d.utils.enums", os.path.join(REF, "utils/enums.py"))
    _load("torch_points3d.utils.debugging_vars", os.path.join(REF, "utils/debugging_vars.py"))
    _load("torch_points3d.utils.model_building_utils.activation_resolver",
          os.path.join(REF, "utils/model_building_utils/activation_resolver.py"))
    _load("torch_points3d.core.common_modules.base_modules", os.path.join(REF, "core/common_modules/base_modules.py"))
    _load("torch_points3d.core.common_modules.dense_modules", os.path.join(REF, "core/common_modules/dense_modules.py"))
    so = _module("torch_points3d.core.spatial_ops")
    for f in ("sampling", "neighbour_finder", "interpolate"):
        mod = _load("torch_points3d.core.spatial_ops." + f, os.path.join(REF, "core/spatial_ops", f + ".py"))
        so.__dict__.update({k: v for k, v in vars(mod).items() if not k.startswith("_")})
    bc = _module("torch_points3d.core.base_conv")
    base = _load("torch_points3d.core.base_conv.base_conv", os.path.join(REF, "core/base_conv/base_conv.py"))
    bc.__dict__.update({k: v for k, v in vars(base).items() if not k.startswith("_")})
    _load("torch_points3d.core.base_conv.dense", os.path.join(REF, "core/base_conv/dense.py"))
    return _load("torch_points3d.modules.pointnet2.dense", os.path.join(REF, "modules/pointnet2/dense.py"))


def main():
    kernels = dense_chain_cases.Kernels()
    ref = load_reference(kernels)
    ns = types.SimpleNamespace(PointNetMSGDown=ref.PointNetMSGDown, DenseFPModule=ref.DenseFPModule,
                               GlobalDenseBaseModule=ref.GlobalDenseBaseModule)
    out = dense_chain_cases.run_cases(ns, Data)
    out["kernel_calls"] = np.array(sorted("%s=%d" % kv for kv in kernels.calls.items()))
    np.savez_compressed(os.path.join(HERE, "dense_chain.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
