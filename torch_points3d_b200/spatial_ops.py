"""Host-side mirror of torch_points3d/core/spatial_ops (sampling.py, neighbour_finder.py,
interpolate.py): same class names, constructor arguments, call conventions and error
behaviour, with the third-party kernels replaced by libp3d_b200.so (ops.py).

Only the classes on the hot path are mirrored: the voxel-grid sampler and
the debugging histogram hooks stay in the reference.
"""
import math
from abc import ABC, abstractmethod
from types import SimpleNamespace
from typing import List, Union

import torch

from . import ops

DENSE, PARTIAL_DENSE, MESSAGE_PASSING = "dense", "partial_dense", "message_passing"  # utils/enums.py:10-14


def _is_list(v) -> bool:
    return isinstance(v, (list, tuple))


# ----------------------------------------------------------------- samplers (sampling.py) ---
class BaseSampler(ABC):
    """sampling.py:13-49: exactly one of ratio / num_to_sample / subsampling_param."""

    def __init__(self, ratio=None, num_to_sample=None, subsampling_param=None):
        if num_to_sample is not None:
            if (ratio is not None) or (subsampling_param is not None):
                raise ValueError("Can only specify ratio or num_to_sample or subsampling_param, not several !")
            self._num_to_sample = num_to_sample
        elif ratio is not None:
            self._ratio = ratio
        elif subsampling_param is not None:
            self._subsampling_param = subsampling_param
        else:
            raise Exception('At least ["ratio, num_to_sample, subsampling_param"] should be defined')

    def __call__(self, pos, x=None, batch=None):
        return self.sample(pos, batch=batch, x=x)

    def _get_num_to_sample(self, batch_size) -> int:
        if hasattr(self, "_num_to_sample"):
            return self._num_to_sample
        return math.floor(batch_size * self._ratio)

    def _get_ratio_to_sample(self, batch_size) -> float:
        if hasattr(self, "_ratio"):
            return self._ratio
        return self._num_to_sample / float(batch_size)

    @abstractmethod
    def sample(self, pos, x=None, batch=None):
        pass


class FPSSampler(BaseSampler):
    """sampling.py:53-63 -> torch_geometric.nn.fps.  `random_start` defaults to the reference's
    behaviour (True); parity tests and benchmarks pass False (start index 0 of each cloud)."""

    def __init__(self, ratio=None, num_to_sample=None, subsampling_param=None, random_start=True):
        super().__init__(ratio, num_to_sample, subsampling_param)
        self._random_start = random_start

    def sample(self, pos, batch, **kwargs):
        if len(pos.shape) != 2:
            raise ValueError(" This class is for sparse data and expects the pos tensor to be of dimension 2")
        return ops.fps(pos, batch, ratio=self._get_ratio_to_sample(pos.shape[0]), random_start=self._random_start)


class DenseFPSSampler(BaseSampler):
    """sampling.py:84-100 -> tp.furthest_point_sample; pos [B,N,3] -> idx [B,num_sample]."""

    def sample(self, pos, **kwargs):
        if len(pos.shape) != 3:
            raise ValueError(" This class is for dense data and expects the pos tensor to be of dimension 2")
        return ops.furthest_point_sample(pos, self._get_num_to_sample(pos.shape[1]))


class RandomSampler(BaseSampler):
    """sampling.py:103-112 (RandLA-Net style): draws with replacement over the whole batch.  `sorted_indices=True`
    returns the draw in ascending order, so that `batch[idx]` stays sorted -- which the batched searches downstream
    require (torch_cluster's knn states the same precondition; the reference feeds it the unsorted draw)."""

    def __init__(self, ratio=None, num_to_sample=None, subsampling_param=None, sorted_indices=False):
        super().__init__(ratio, num_to_sample, subsampling_param)
        self._sorted = sorted_indices

    def sample(self, pos, batch=None, **kwargs):
        if len(pos.shape) != 2:
            raise ValueError(" This class is for sparse data and expects the pos tensor to be of dimension 2")
        idx = torch.randint(0, pos.shape[0], (self._get_num_to_sample(pos.shape[0]),), device=pos.device)
        return torch.sort(idx)[0] if self._sorted else idx


class DenseRandomSampler(BaseSampler):
    """sampling.py:115-127."""

    def sample(self, pos, **kwargs):
        if len(pos.shape) != 3:
            raise ValueError(" This class is for dense data and expects the pos tensor to be of dimension 2")
        return torch.randint(0, pos.shape[1], (self._get_num_to_sample(pos.shape[1]),), device=pos.device)


# --------------------------------------------------- neighbour finders (neighbour_finder.py) ---
class BaseNeighbourFinder(ABC):
    def __call__(self, x, y, batch_x, batch_y):
        return self.find_neighbours(x, y, batch_x, batch_y)

    @abstractmethod
    def find_neighbours(self, x, y, batch_x, batch_y):
        pass

    def __repr__(self):
        return str(self.__class__.__name__) + " " + str(self.__dict__)


class RadiusNeighbourFinder(BaseNeighbourFinder):
    """neighbour_finder.py:25-39.  MESSAGE_PASSING -> edge list; anything else is routed to
    ball_query(mode=conv_type) exactly as the reference's always-true `elif` does (:34)."""

    def __init__(self, radius: float, max_num_neighbors: int = 64, conv_type=MESSAGE_PASSING):
        self._radius = radius
        self._max_num_neighbors = max_num_neighbors
        self._conv_type = conv_type.lower()

    def find_neighbours(self, x, y, batch_x=None, batch_y=None):
        if self._conv_type == MESSAGE_PASSING:
            return ops.radius(x, y, self._radius, batch_x, batch_y, max_num_neighbors=self._max_num_neighbors)
        return ops.ball_query(self._radius, self._max_num_neighbors, x, y, mode=self._conv_type, batch_x=batch_x,
                              batch_y=batch_y, return_dist=False)[0]


class KNNNeighbourFinder(BaseNeighbourFinder):
    """neighbour_finder.py:42-47 -> knn edge list [2, E] (row = y index, col = x index)."""

    def __init__(self, k):
        self.k = k

    def find_neighbours(self, x, y, batch_x, batch_y):
        return ops.knn_edges(x, y, self.k, batch_x, batch_y)


class DilatedKNNNeighbourFinder(BaseNeighbourFinder):
    """neighbour_finder.py:50-68: k*dilation nearest, then k of them drawn at random."""

    def __init__(self, k, dilation):
        self.k = k
        self.dilation = dilation
        self.initialFinder = KNNNeighbourFinder(k * dilation)

    def find_neighbours(self, x, y, batch_x, batch_y):
        row, col = self.initialFinder.find_neighbours(x, y, batch_x, batch_y)
        index = torch.randint(self.k * self.dilation, (len(y), self.k), device=row.device, dtype=torch.long)
        arange = torch.arange(len(y), dtype=torch.long, device=row.device) * (self.k * self.dilation)
        index = (index + arange.view(-1, 1)).view(-1)
        return row[index], col[index]


class BaseMSNeighbourFinder(ABC):
    def __call__(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        return self.find_neighbours(x, y, batch_x=batch_x, batch_y=batch_y, scale_idx=scale_idx)

    @abstractmethod
    def find_neighbours(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        pass

    @property
    @abstractmethod
    def num_scales(self):
        pass


class MultiscaleRadiusNeighbourFinder(BaseMSNeighbourFinder):
    """neighbour_finder.py:89-153: list/scalar broadcasting of radius and max_num_neighbors."""

    def __init__(self, radius: Union[float, List[float]], max_num_neighbors: Union[int, List[int]] = 64):
        if not _is_list(max_num_neighbors) and _is_list(radius):
            self._radius = list(radius)
            self._max_num_neighbors = [max_num_neighbors for _ in self._radius]
        elif not _is_list(radius) and _is_list(max_num_neighbors):
            self._max_num_neighbors = list(max_num_neighbors)
            self._radius = [radius for _ in self._max_num_neighbors]
        elif _is_list(max_num_neighbors):
            if len(max_num_neighbors) != len(radius):
                raise ValueError("Both lists max_num_neighbors and radius should be of the same length")
            self._max_num_neighbors = list(max_num_neighbors)
            self._radius = list(radius)
        else:
            self._max_num_neighbors = [max_num_neighbors]
            self._radius = [radius]

    def find_neighbours(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        if scale_idx >= self.num_scales:
            raise ValueError("Scale %i is out of bounds %i" % (scale_idx, self.num_scales))
        return ops.radius(x, y, self._radius[scale_idx], batch_x, batch_y,
                          max_num_neighbors=self._max_num_neighbors[scale_idx])

    @property
    def num_scales(self):
        return len(self._radius)

    def __call__(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        return self.find_neighbours(x, y, batch_x, batch_y, scale_idx)


class DenseRadiusNeighbourFinder(MultiscaleRadiusNeighbourFinder):
    """neighbour_finder.py:156-178 -> tp.ball_query(r, K, x, y)[0], idx [B,S,K] int64."""

    def find_neighbours(self, x, y, scale_idx=0):
        if scale_idx >= self.num_scales:
            raise ValueError("Scale %i is out of bounds %i" % (scale_idx, self.num_scales))
        return ops.ball_query(self._radius[scale_idx], self._max_num_neighbors[scale_idx], x, y,
                              return_dist=False)[0]

    def __call__(self, x, y, scale_idx=0, **kwargs):
        return self.find_neighbours(x, y, scale_idx)


# ------------------------------------------------------------- KNNInterpolate (interpolate.py) ---
class KNNInterpolate:
    """interpolate.py:7-70.  `query` / `support` are any objects with .pos, .x and optionally
    .batch (torch_geometric Data in the reference)."""

    def __init__(self, k):
        self.k = k

    @staticmethod
    def _batch(data):
        b = getattr(data, "batch", None)
        if b is None:
            b = torch.zeros((data.pos.shape[0],), dtype=torch.long, device=data.pos.device)
        return b

    def precompute(self, query, support):
        pos_x, pos_y = query.pos, support.pos
        batch_y, batch_x = self._batch(support), self._batch(query)
        with torch.no_grad():
            y_idx, x_idx = ops.knn_edges(pos_x, pos_y, self.k, batch_x=batch_x, batch_y=batch_y)
            diff = pos_x[x_idx] - pos_y[y_idx]
            squared_distance = (diff * diff).sum(dim=-1, keepdim=True)
            weights = 1.0 / torch.clamp(squared_distance, min=1e-16)
        normalisation = torch.zeros((pos_y.size(0), 1), dtype=weights.dtype, device=weights.device)
        normalisation.index_add_(0, y_idx, weights)
        return SimpleNamespace(num_nodes=pos_y.shape[0], x_idx=x_idx, y_idx=y_idx, weights=weights,
                               normalisation=normalisation)

    def __call__(self, query, support, precomputed=None):
        if precomputed:
            num_points = support.pos.size(0)
            if num_points != precomputed.num_nodes:
                raise ValueError("Precomputed indices do not match with the data given to the transform")
            x = query.x
            y = torch.zeros((num_points, x.shape[1]), dtype=x.dtype, device=x.device)
            y.index_add_(0, precomputed.y_idx, x[precomputed.x_idx] * precomputed.weights)
            return y / precomputed.normalisation
        return ops.knn_interpolate(query.x, query.pos, support.pos, self._batch(query), self._batch(support),
                                   k=self.k)
