"""Run tests/golden/mp_chain_cases.py on the REFERENCE's own MESSAGE_PASSING PointNet++ modules and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): modules/pointnet2/message_passing.py,
core/base_conv/{base_conv,message_passing}.py, core/common_modules/base_modules.py, core/spatial_ops/*.py,
utils/enums.py.  torch_geometric is absent from this image; stood in for: `radius` (oracle), `knn` /
`knn_interpolate` / `global_*_pool` (torch restatements of their definitions) and `PointConv` -- restated below from
torch_geometric 1.7.2 (remove self loops, add (i, i) for i < min(#source, #target), message = local_nn([x_j,
pos_j - pos_i]), max over the edges of a target with 0 for targets without edges, global_nn).  The fixture therefore
pins BaseMSConvolutionDown.forward, GlobalBaseModule and FPModule on the reference's own code; PointConv itself stays
a restatement on both sides.  /root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_mp_chain.py
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
import mp_chain_cases as cases  # noqa: E402


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


class PointConv(torch.nn.Module):
    def __init__(self, local_nn=None, global_nn=None, add_self_loops=True, **kwargs):
        super().__init__()
        self.local_nn, self.global_nn, self.add_self_loops = local_nn, global_nn, add_self_loops

    def forward(self, x, pos, edge_index):
        pos_src, pos_dst = (pos, pos) if torch.is_tensor(pos) else pos
        if isinstance(x, tuple):
            x = x[0]
        j, i = edge_index[0], edge_index[1]
        if self.add_self_loops:
            keep = j != i
            loops = torch.arange(min(pos_src.shape[0], pos_dst.shape[0]), dtype=j.dtype)
            j, i = torch.cat([j[keep], loops]), torch.cat([i[keep], loops])
        msg = pos_src[j] - pos_dst[i]
        if x is not None:
            msg = torch.cat([x[j], msg], dim=1)
        if self.local_nn is not None:
            msg = self.local_nn(msg)
        out = cases.segment_max(msg, i, pos_dst.shape[0])
        return self.global_nn(out) if self.global_nn is not None else out


def main():
    def unused(*a, **kw):
        raise AssertionError("not on the tested paths")

    _module("torch_geometric")
    _module("torch_geometric.nn", fps=unused, knn=cases.knn, radius=cases.radius, voxel_grid=unused,
            knn_interpolate=cases.knn_interpolate, global_max_pool=cases.global_max_pool,
            global_mean_pool=cases.global_mean_pool, PointConv=PointConv)
    _module("torch_geometric.nn.pool")
    _module("torch_geometric.nn.pool.consecutive", consecutive_cluster=unused)
    _module("torch_geometric.nn.pool.pool", pool_pos=unused, pool_batch=unused)
    _module("torch_geometric.data", Data=Bag, Batch=Bag)
    _module("torch_scatter", scatter_add=unused)
    _module("torch_points_kernels", furthest_point_sample=unused, ball_query=unused)
    for pkg in ("torch_points3d", "torch_points3d.utils", "torch_points3d.core", "torch_points3d.core.base_conv"):
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
    _load("torch_points3d.core.base_conv.base_conv", os.path.join(REF, "core/base_conv/base_conv.py"))
    mpb = _load("torch_points3d.core.base_conv.message_passing", os.path.join(REF, "core/base_conv/message_passing.py"))
    ref = _load("ref_pointnet2_mp", os.path.join(REF, "modules/pointnet2/message_passing.py"))
    ns = types.SimpleNamespace(SAModule=ref.SAModule, GlobalBaseModule=mpb.GlobalBaseModule, FPModule=mpb.FPModule)
    out = cases.run_cases(ns, Bag)
    np.savez_compressed(os.path.join(HERE, "mp_chain.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
