"""Run tests/golden/kpconv_blocks_cases.py on the REFERENCE's own KPConv blocks and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): modules/KPConv/{blocks,kernels,kernel_utils,plyutils,
convolution_ops,losses}.py, core/common_modules/{base_modules,gathering}.py, core/spatial_ops/*.py, utils/enums.py.
KPConv_ops is the reference's own torch code, executed on CPU.  Absent third-party modules: `torch_points_kernels.
ball_query` and `GridSampling3D` are replaced by the oracle-backed stand-ins of kpconv_blocks_cases, matplotlib /
torch_geometric / torch_scatter by empty stand-ins (nothing of them is reached), `models.base_model.
BaseInternalLossModule` by a bare nn.Module subclass.  /root/reference does not exist on the GPU box: the test only
reads the committed .npz.

    python tests/golden/make_golden_kpconv_blocks.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import kpconv_blocks_cases as cases  # noqa: E402


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
    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def clone(self):
        return Bag(**self.__dict__)


def load_reference():
    """the reference's modules/KPConv/blocks.py module object, loaded by file path over the stand-ins above"""
    def unused(*a, **kw):
        raise AssertionError("not on the tested paths")

    def tp_ball_query(radius, nsample, x, y, mode="dense", batch_x=None, batch_y=None, sort=False):
        return cases.ball_query(radius, nsample, x, y, mode=mode, batch_x=batch_x, batch_y=batch_y)

    _module("matplotlib")
    _module("matplotlib.pyplot")
    _module("torch_geometric")
    _module("torch_geometric.nn", fps=unused, knn=unused, radius=unused, voxel_grid=unused, knn_interpolate=unused,
            global_max_pool=unused, global_mean_pool=unused)
    _module("torch_geometric.nn.pool")
    _module("torch_geometric.nn.pool.consecutive", consecutive_cluster=unused)
    _module("torch_geometric.nn.pool.pool", pool_pos=unused, pool_batch=unused)
    _module("torch_geometric.data", Data=Bag, Batch=Bag)
    _module("torch_scatter", scatter_add=unused)
    _module("torch_points_kernels", furthest_point_sample=unused, ball_query=tp_ball_query)
    for pkg in ("torch_points3d", "torch_points3d.utils", "torch_points3d.core", "torch_points3d.models",
                "torch_points3d.modules", "torch_points3d.modules.KPConv", "torch_points3d.core.base_conv"):
        _module(pkg)
    _module("torch_points3d.utils.config", is_list=lambda e: isinstance(e, list))
    _load("torch_points3d.utils.enums", os.path.join(REF, "utils/enums.py"))
    _load("torch_points3d.utils.debugging_vars", os.path.join(REF, "utils/debugging_vars.py"))
    base = _load("torch_points3d.core.common_modules.base_modules", os.path.join(REF, "core/common_modules/base_modules.py"))
    _module("torch_points3d.core.common_modules", **{k: v for k, v in vars(base).items() if not k.startswith("_")})
    sys.modules["torch_points3d.core.common_modules.base_modules"] = base
    _load("torch_points3d.core.common_modules.gathering", os.path.join(REF, "core/common_modules/gathering.py"))
    so = _module("torch_points3d.core.spatial_ops")
    for f in ("sampling", "neighbour_finder", "interpolate"):
        mod = _load("torch_points3d.core.spatial_ops." + f, os.path.join(REF, "core/spatial_ops", f + ".py"))
        so.__dict__.update({k: v for k, v in vars(mod).items() if not k.startswith("_")})
    _module("torch_points3d.core.data_transform", GridSampling3D=cases.GridSampling3D)
    _module("torch_points3d.core.base_conv.message_passing", GlobalBaseModule=type("GlobalBaseModule", (torch.nn.Module,), {}))
    _module("torch_points3d.models.base_model", BaseInternalLossModule=type("BaseInternalLossModule", (torch.nn.Module,), {}))
    kp = "torch_points3d.modules.KPConv."
    for f in ("plyutils", "kernel_utils", "losses", "convolution_ops", "kernels"):
        _load(kp + f, os.path.join(REF, "modules/KPConv", f + ".py"))
    return _load(kp + "blocks", os.path.join(REF, "modules/KPConv/blocks.py"))


def main():
    ref = load_reference()
    np.random.seed(0)  # load_kernels draws a random rotation; the kernel points are overwritten by seeded values anyway
    ns = types.SimpleNamespace(SimpleBlock=ref.SimpleBlock, ResnetBBlock=ref.ResnetBBlock, KPDualBlock=ref.KPDualBlock)
    out = cases.run_cases(ns, Bag)
    np.savez_compressed(os.path.join(HERE, "kpconv_blocks.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
