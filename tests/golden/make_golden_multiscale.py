"""Run tests/golden/multiscale_cases.py on the REFERENCE's own pyramid code and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): datasets/multiscale_data.py (MultiScaleData,
MultiScaleBatch, from_data_list_token), core/spatial_ops/interpolate.py (KNNInterpolate.precompute), and the class
MultiScaleTransform of core/data_transform/transforms.py:579-651 -- that one class is compiled from the file's own
text (ast) because the module around it imports sklearn KD-trees, numba and the registration datasets, none of which
this path touches.  torch_geometric is absent from this image: `Data` / `Batch` below restate what torch_geometric
1.7.2's classes do for these code paths (attribute table, `keys` without None / dunder entries, `__cat_dim__` / `__inc__`
rules, `num_nodes`, clone / contiguous, and a plain `Batch.from_data_list`; `ptr` and `follow_batch` are not
modelled); `knn` and `scatter_add` are brute-force stand-ins.
/root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_multiscale.py
"""
import ast
import copy
import functools
import importlib.util
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
import multiscale_cases as cases  # noqa: E402


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


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, normal=None, face=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        self.normal, self.face = normal, face
        for key, item in kwargs.items():
            if key == "num_nodes":
                self.__num_nodes__ = item
            else:
                self[key] = item

    @classmethod
    def from_dict(cls, dictionary):
        data = cls()
        for key, item in dictionary.items():
            data[key] = item
        return data

    def __getitem__(self, key):
        return getattr(self, key, None)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    @property
    def keys(self):
        return [k for k in self.__dict__ if self[k] is not None and k[:2] != "__" and k[-2:] != "__"]

    def __contains__(self, key):
        return key in self.keys

    def __iter__(self):
        for key in sorted(self.keys):
            yield key, self[key]

    def __call__(self, *keys):
        for key in sorted(self.keys) if not keys else keys:
            if key in self:
                yield key, self[key]

    def __cat_dim__(self, key, value):
        return -1 if re.search("(index|face)", key) else 0

    def __inc__(self, key, value):
        return self.num_nodes if re.search("(index|face)", key) else 0

    @property
    def num_nodes(self):
        if hasattr(self, "__num_nodes__"):
            return self.__num_nodes__
        for key, item in self("x", "pos", "normal", "batch"):
            return item.size(self.__cat_dim__(key, item))
        if self.edge_index is not None:
            return int(self.edge_index.max()) + 1
        return None

    @num_nodes.setter
    def num_nodes(self, value):
        self.__num_nodes__ = value

    def apply(self, func, *keys):
        for key, item in self(*keys):
            if torch.is_tensor(item):
                self[key] = func(item)
        return self

    def contiguous(self, *keys):
        return self.apply(lambda t: t.contiguous(), *keys)

    def clone(self):
        return self.__class__.from_dict({k: v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
                                         for k, v in self.__dict__.items()})


class Batch(Data):
    def __init__(self, batch=None, **kwargs):
        super().__init__(**kwargs)
        self.batch = batch
        self.__data_class__ = Data
        self.__slices__ = None

    @staticmethod
    def from_data_list(data_list, follow_batch=[]):
        keys = sorted(set().union(*[set(d.keys) for d in data_list]))
        assert "batch" not in keys
        batch = Batch()
        cols, shift, vec = {k: [] for k in keys}, {k: 0 for k in keys}, []
        for i, data in enumerate(data_list):
            for key in data.keys:
                item = data[key]
                if torch.is_tensor(item) and item.dtype != torch.bool:
                    item = item + shift[key]
                shift[key] += data.__inc__(key, item)
                cols[key].append(item)
            vec.append(torch.full((data.num_nodes,), i, dtype=torch.long))
        for key in keys:
            item = cols[key][0]
            batch[key] = torch.cat(cols[key], dim=data_list[0].__cat_dim__(key, item)) if torch.is_tensor(item) \
                else torch.tensor(cols[key])
        batch.batch = torch.cat(vec)
        return batch.contiguous()


def scatter_add(src, index, dim=0, dim_size=None):
    assert dim == 0
    return torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, src)


def main():
    def unused(*a, **kw):
        raise AssertionError("not on the tested paths")

    _module("torch_geometric", is_debug_enabled=lambda: False)
    _module("torch_geometric.data", Data=Data, Batch=Batch)
    _module("torch_geometric.nn", knn=cases.knn, knn_interpolate=unused)
    _module("torch_scatter", scatter_add=scatter_add)
    ms = _load("ref_multiscale_data", os.path.join(REF, "datasets/multiscale_data.py"))
    interp = _load("ref_interpolate", os.path.join(REF, "core/spatial_ops/interpolate.py"))
    path = os.path.join(REF, "core/data_transform/transforms.py")
    tree = ast.parse(open(path).read(), filename=path)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MultiScaleTransform")
    scope = {"partial": functools.partial, "Data": Data, "MultiScaleData": ms.MultiScaleData, "torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), scope)
    ns = types.SimpleNamespace(Data=Data, MultiScaleData=ms.MultiScaleData, MultiScaleBatch=ms.MultiScaleBatch,
                               MultiScaleTransform=scope["MultiScaleTransform"],
                               from_data_list_token=ms.from_data_list_token, KNNInterpolate=interp.KNNInterpolate)
    out = cases.run_cases(ns)
    np.savez_compressed(os.path.join(HERE, "multiscale.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
