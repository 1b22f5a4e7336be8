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


def _per_point_items(data):
    """(key, tensor) of every attribute that has one row per point"""
    n = data.pos.shape[0]
    return [(k, v) for k, v in list(data) if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n]


def shuffle_data(data):
    """grid_transform.py:23-30: one random permutation applied to every per-point tensor"""
    perm = torch.randperm(data.pos.shape[0], device=data.pos.device)
    for key, item in _per_point_items(data):
        data[key] = item[perm]
    return data


def _voxel_average(key, item, voxels):
    """what mode="mean" stores for one attribute (grid_transform.py:66-79): labels vote, everything else is averaged;
    bool attributes go through int32 and back, integer attributes keep their dtype (floor of the mean)"""
    was_bool = item.dtype == torch.bool
    if was_bool:
        item = item.int()
    if key in _INTEGER_LABEL_KEYS:
        out = ops.segment_vote(item, voxels)
    elif key == "pos" and voxels.mean_pos is not None:
        out = voxels.mean_pos  # already produced by the clustering kernel
    else:
        out = ops.segment_mean(item, voxels)
        if not item.dtype.is_floating_point:
            out = out.floor().to(item.dtype)
    return out.bool() if was_bool else out


def group_data(data, cluster=None, unique_pos_indices=None, mode="last", skip_keys=[], voxels=None):
    """grid_transform.py:33-81: reduce every per-point attribute to one row per voxel.  mode "last" (and, in any
    mode, `batch` and the saved origin ids) takes the voxel's representative point `unique_pos_indices`; mode "mean"
    averages, with a majority vote for the label keys.  `voxels` is the `ops.VoxelClusters` the cluster ids came
    from (the segment list the mean / vote kernels run over); GridSampling3D passes it, a caller holding only
    `cluster` gets it rebuilt."""
    assert mode in ["mean", "last"]
    if mode == "mean" and cluster is None:
        raise ValueError("In mean mode the cluster argument needs to be specified")
    if mode == "last" and unique_pos_indices is None:
        raise ValueError("In last mode the unique_pos_indices argument needs to be specified")

    by_representative = (SaveOriginalPosId.KEY, "batch")
    num_nodes = data.num_nodes
    for key, item in list(data):
        if re.search("edge", key):
            raise ValueError("Edges not supported. Wrong data type.")
        if key in skip_keys or not torch.is_tensor(item) or item.size(0) != num_nodes:
            continue
        if mode == "last" or key in by_representative:
            data[key] = item[unique_pos_indices]
            continue
        if voxels is None:
            voxels = ops.voxels_from_cluster(cluster, unique_pos_indices)
        data[key] = _voxel_average(key, item, voxels)
    return data


class GridSampling3D:
    """grid_transform.py:84-141: clusters points into voxels of edge `size`; mode "mean" averages positions and
    features (majority vote for labels), mode "last" keeps one point per voxel after a random shuffle;
    `quantize_coords` adds the int32 voxel coordinates as `coords`.  One launch of the clustering kernel yields ids,
    representatives, mean positions and coordinates together."""

    def __init__(self, size, quantize_coords=False, mode="mean", verbose=False):
        self._grid_size, self._quantize_coords, self._mode = size, quantize_coords, mode

    def _process(self, data):
        averaging = self._mode != "last"
        if not averaging:
            data = shuffle_data(data)
        voxels = ops.grid_cluster(data.pos, self._grid_size, data["batch"] if "batch" in data else None,
                                  want_mean=averaging, want_coords=self._quantize_coords)
        data = group_data(data, voxels.cluster, voxels.unique_pos_indices, mode=self._mode, voxels=voxels)
        if self._quantize_coords:
            data.coords = voxels.coords
        data.grid_size = torch.tensor([self._grid_size])
        return data

    def __call__(self, data):
        return [self._process(d) for d in data] if isinstance(data, list) else self._process(data)

    def __repr__(self):
        return "%s(grid_size=%s, quantize_coords=%s, mode=%s)" % (type(self).__name__, self._grid_size,
                                                                  self._quantize_coords, self._mode)


class SaveOriginalPosId:
    """grid_transform.py:144-167: stores every point's index as `origin_id` (unless already there) so that it can be
    traced through the samplers"""

    KEY = "origin_id"

    def _process(self, data):
        if not hasattr(data, self.KEY):
            setattr(data, self.KEY, torch.arange(data.pos.shape[0], device=data.pos.device))
        return data

    def __call__(self, data):
        return [self._process(d) for d in data] if isinstance(data, list) else self._process(data)

    def __repr__(self):
        return type(self).__name__
