"""Run tests/golden/hostlogic_cases.py on the REFERENCE's own core/spatial_ops classes and store what they do.

The three reference files (sampling.py, neighbour_finder.py, interpolate.py) are loaded BY FILE PATH from
/root/reference -- nothing is copied -- with the third-party modules they import (torch_geometric, torch_scatter,
torch_points_kernels: absent from this image) replaced by the recording stand-ins of hostlogic_cases.Kernels, so what
is captured is the reference's HOST logic: constructor rules, attributes, exception types and messages, which kernel
is called with which scalars, the interpolation arithmetic.  /root/reference does not exist on the GPU box: the test
only reads the committed JSON.

    python tests/golden/make_golden_hostlogic.py
"""
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
import hostlogic_cases  # noqa: E402


def _module(name, **attrs):
    m = types.ModuleType(name)
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
    """the slice of torch_geometric.data.Data the reference's KNNInterpolate touches"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    k = hostlogic_cases.Kernels()

    def scatter_add(src, index, dim=0, dim_size=None):
        out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
        return out.index_add_(0, index, src)

    def unused(*a, **kw):
        raise AssertionError("not on the tested paths")

    _module("torch_geometric")
    _module("torch_geometric.nn", fps=k.fps, knn=k.knn, radius=k.radius, voxel_grid=unused, knn_interpolate=unused)
    _module("torch_geometric.nn.pool")
    _module("torch_geometric.nn.pool.consecutive", consecutive_cluster=unused)
    _module("torch_geometric.nn.pool.pool", pool_pos=unused, pool_batch=unused)
    _module("torch_geometric.data", Data=Data)
    _module("torch_scatter", scatter_add=scatter_add)
    _module("torch_points_kernels", furthest_point_sample=k.furthest_point_sample, ball_query=k.ball_query)
    for pkg in ("torch_points3d", "torch_points3d.utils", "torch_points3d.core"):
        _module(pkg)
    # utils/config.py needs omegaconf for ListConfig only; is_list is `isinstance(entity, list) or ListConfig` (:58-59)
    _module("torch_points3d.utils.config", is_list=lambda e: isinstance(e, list))
    _load("torch_points3d.utils.enums", os.path.join(REF, "utils/enums.py"))
    _load("torch_points3d.utils.debugging_vars", os.path.join(REF, "utils/debugging_vars.py"))
    ns = types.SimpleNamespace()
    for f in ("sampling", "neighbour_finder", "interpolate"):
        mod = _load("ref_spatial_ops_" + f, os.path.join(REF, "core/spatial_ops", f + ".py"))
        for name, obj in vars(mod).items():
            if isinstance(obj, type):
                setattr(ns, name, obj)
    result = hostlogic_cases.run_cases(ns, k, data_cls=Data)
    with open(os.path.join(HERE, "spatial_ops_hostlogic.json"), "w") as f:
        json.dump(result, f, indent=1, sort_keys=True)
    print(json.dumps(result, indent=1, sort_keys=True)[:3000])


if __name__ == "__main__":
    main()
