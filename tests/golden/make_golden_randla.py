"""Run tests/golden/randla_cases.py on the REFERENCE's own RandLA-Net modules and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): modules/RandLANet/modules.py,
core/base_conv/{base_conv,message_passing}.py, core/common_modules/base_modules.py, core/spatial_ops/*.py,
utils/enums.py.  torch_geometric is absent from this image: `knn` is the brute-force stand-in of randla_cases, and
`MessagePassing` is the minimal restatement below of what torch_geometric 1.7.2's base class does for these modules
(flow source_to_target: `_j` arguments are gathered with edge_index[0] from the first element of a pair, `_i`
arguments with edge_index[1] from the second; messages are summed per target; `update` gets the sums).
/root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_randla.py
"""
import importlib.util
import inspect
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
import randla_cases as cases  # noqa: E402


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


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0):
        super().__init__()
        assert aggr in ("add", "max") and flow == "source_to_target"
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        size = [None, None]                       # [number of sources, number of targets], as __set_size__ infers them
        args = {}
        for name in inspect.signature(self.message).parameters:
            value = kwargs.get(name[:-2])
            if value is None:
                args[name] = None
                continue
            dim = 0 if name.endswith("_j") else 1
            if isinstance(value, tuple):
                if size[1 - dim] is None:
                    size[1 - dim] = value[1 - dim].shape[0]
                value = value[dim]
            if size[dim] is None:
                size[dim] = value.shape[0]
            args[name] = value[j] if dim == 0 else value[i]
        msg = self.message(**args)
        num_targets = size[1] if size[1] is not None else size[0]
        if self.aggr == "max":   # torch_scatter's max: 0 for a target without edges
            pooled = torch.zeros((num_targets, msg.shape[1]), dtype=msg.dtype)
            for t in torch.unique(i).tolist():
                pooled[t] = msg[i == t].max(dim=0).values
        else:
            pooled = torch.zeros((num_targets, msg.shape[1]), dtype=msg.dtype).index_add_(0, i, msg)
        return self.update(pooled)


def load_reference_base(radius=None):
    """stand-ins + the reference's base classes of the MESSAGE_PASSING format, loaded by file path"""
    def unused(*a, **kw):
        raise AssertionError("not on the tested paths")

    _module("torch_geometric")
    _module("torch_geometric.nn", fps=unused, knn=cases.knn, radius=radius if radius is not None else unused, voxel_grid=unused, knn_interpolate=unused,
            global_max_pool=unused, global_mean_pool=unused, MessagePassing=MessagePassing)
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
    _load("torch_points3d.core.base_conv.message_passing", os.path.join(REF, "core/base_conv/message_passing.py"))


def main():
    load_reference_base()
    ref = _load("ref_randla_modules", os.path.join(REF, "modules/RandLANet/modules.py"))
    ns = types.SimpleNamespace(RandlaKernel=ref.RandlaKernel, RandlaConv=ref.RandlaConv, RandLANetRes=ref.RandLANetRes)
    out = cases.run_cases(ns, Bag)
    np.savez_compressed(os.path.join(HERE, "randla_chain.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
