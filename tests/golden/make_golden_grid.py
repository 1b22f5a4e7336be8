"""Run tests/golden/grid_cases.py on the REFERENCE's own core/data_transform/grid_transform.py and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied).  torch_geometric / torch_cluster / torch_scatter are
absent from this image: `Data` is the restatement in make_golden_multiscale.py, the four operators are the
restatements of grid_cases.py (same functions on both sides of the comparison).
/root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_grid.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
import grid_cases as cases  # noqa: E402
from make_golden_multiscale import Data, _load, _module  # noqa: E402


def main():
    _module("torch_geometric")
    _module("torch_geometric.data", Data=Data)
    _module("torch_geometric.nn", voxel_grid=lambda pos, batch, size: cases.voxel_key(pos, batch, size))
    _module("torch_geometric.nn.pool")
    _module("torch_geometric.nn.pool.consecutive", consecutive_cluster=cases.consecutive_cluster)
    _module("torch_scatter", scatter_mean=cases.scatter_mean, scatter_add=cases.scatter_add)
    _module("torch_cluster", grid_cluster=lambda pos, size: cases.voxel_key(pos, None, size))
    ref = _load("ref_grid_transform", os.path.join(REF, "core/data_transform/grid_transform.py"))
    ns = types.SimpleNamespace(Data=Data, GridSampling3D=ref.GridSampling3D, group_data=ref.group_data,
                               SaveOriginalPosId=ref.SaveOriginalPosId)
    out = cases.run_cases(ns)
    np.savez_compressed(os.path.join(HERE, "grid_transform.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
