"""Host-side mirror of torch_points3d/core/data_transform/grid_transform.py:21-141 (shuffle_data, group_data,
GridSampling3D) over the voxel-clustering kernels of libp3d_b200.so (csrc/grid.cu) -- SURVEY.md section 8f row 1.

The reference transform works on `torch_geometric.data.Data`; any object that iterates as (key, tensor) pairs,
supports item assignment and has `num_nodes` works here, and `Data` below is a minimal stand-in with that
interface (torch_geometric is not a dependency of this package).
"""
import re

import torch

from . import ops

_INTEGER_LABEL_KEYS = ["y", "instance_labels"]  # grid_transform.py:20


class Data:
    """attribute bag with the slice of the torch_geometric `Data` interface the transform uses"""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def __iter__(self):
        for k in sorted(self.keys):
            yield k, self.__dict__[k]

    def __contains__(self, key):
        return self.__dict__.get(key) is not None

    def __getitem__(self, key):
        return self.__dict__[key]

    def __setitem__(self, key, value):
        self.__dict__[key] = value

    @property
    def num_nodes(self):
        return self.pos.shape[0]

    def clone(self):
        return Data(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})


def shuffle_data(data):
    """grid_transform.py:23-30"""
    num_points = data.pos.shape[0]
    shuffle_idx = torch.randperm(num_points, device=data.pos.device)
    for key in set(data.keys):
        item = data[key]
        if torch.is_tensor(item) and num_points == item.shape[0]:
            data[key] = item[shuffle_idx]
    return data


def group_data(data, cluster=None, unique_pos_indices=None, mode="last", skip_keys=[], voxels=None):
    """grid_transform.py:33-81.  `voxels` is the `ops.VoxelClusters` the cluster ids came from (the segment list
    the mean / vote kernels run over); GridSampling3D passes it, a caller holding only `cluster` gets it rebuilt."""
    assert mode in ["mean", "last"]
    if mode == "mean" and cluster is None:
        raise ValueError("In mean mode the cluster argument needs to be specified")
    if mode == "last" and unique_pos_indices is None:
        raise ValueError("In last mode the unique_pos_indices argument needs to be specified")
    if mode == "mean" and voxels is None:
        voxels = ops.voxels_from_cluster(cluster, unique_pos_indices)

    num_nodes = data.num_nodes
    for key, item in list(data):
        if bool(re.search("edge", key)):
            raise ValueError("Edges not supported. Wrong data type.")
        if key in skip_keys:
            continue
        if torch.is_tensor(item) and item.size(0) == num_nodes:
            if mode == "last" or key == "batch" or key == SaveOriginalPosId.KEY:
                data[key] = item[unique_pos_indices]
            elif mode == "mean":
                is_item_bool = item.dtype == torch.bool
                if is_item_bool:
                    item = item.int()
                if key in _INTEGER_LABEL_KEYS:
                    data[key] = ops.segment_vote(item, voxels)
                elif key == "pos" and voxels.mean_pos is not None:
                    data[key] = voxels.mean_pos  # already produced by the clustering kernel
                else:
                    out = ops.segment_mean(item, voxels)
                    # scatter_mean keeps integer dtypes (floor division); floats stay fp32
                    data[key] = out if item.dtype.is_floating_point else out.floor().to(item.dtype)
                if is_item_bool:
                    data[key] = data[key].bool()
    return data


class GridSampling3D:
    """grid_transform.py:84-141: clusters points into voxels of edge `size`; mode "mean" averages positions and
    features (majority vote for labels), mode "last" keeps one point per voxel after a random shuffle."""

    def __init__(self, size, quantize_coords=False, mode="mean", verbose=False):
        self._grid_size = size
        self._quantize_coords = quantize_coords
        self._mode = mode

    def _process(self, data):
        if self._mode == "last":
            data = shuffle_data(data)
        pos = data.pos
        batch = data["batch"] if "batch" in data else None
        voxels = ops.grid_cluster(pos, self._grid_size, batch, want_mean=(self._mode == "mean"),
                                  want_coords=self._quantize_coords)
        data = group_data(data, voxels.cluster, voxels.unique_pos_indices, mode=self._mode, voxels=voxels)
        if self._quantize_coords:
            data.coords = voxels.coords
        data.grid_size = torch.tensor([self._grid_size])
        return data

    def __call__(self, data):
        if isinstance(data, list):
            return [self._process(d) for d in data]
        return self._process(data)

    def __repr__(self):
        return "{}(grid_size={}, quantize_coords={}, mode={})".format(
            self.__class__.__name__, self._grid_size, self._quantize_coords, self._mode)


class SaveOriginalPosId:
    """grid_transform.py:144-167: adds the index of every point so that it can be traced through the samplers"""

    KEY = "origin_id"

    def _process(self, data):
        if hasattr(data, self.KEY):
            return data
        setattr(data, self.KEY, torch.arange(0, data.pos.shape[0], device=data.pos.device))
        return data

    def __call__(self, data):
        if isinstance(data, list):
            return [self._process(d) for d in data]
        return self._process(data)

    def __repr__(self):
        return self.__class__.__name__
