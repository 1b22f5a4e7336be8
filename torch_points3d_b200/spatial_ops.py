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

_SPARSE_MSG = " This class is for sparse data and expects the pos tensor to be of dimension 2"
_DENSE_MSG = " This class is for dense data and expects the pos tensor to be of dimension 2"  # (sic, sampling.py:98)


def _is_list(v) -> bool:
    return isinstance(v, (list, tuple))


def _expect_rank(pos, rank, message):
    if pos.dim() != rank:
        raise ValueError(message)


# ----------------------------------------------------------------- samplers (sampling.py) ---
class BaseSampler(ABC):
    """sampling.py:13-49.  The sample count comes from exactly one source: `num_to_sample` (exclusive with the other
    two), else `ratio`, else `subsampling_param`; the attribute that was given is the only one that exists afterwards
    (callers test for it with hasattr, like the reference)."""

    def __init__(self, ratio=None, num_to_sample=None, subsampling_param=None):
        given = {"_num_to_sample": num_to_sample, "_ratio": ratio, "_subsampling_param": subsampling_param}
        if num_to_sample is not None and not (ratio is None and subsampling_param is None):
            raise ValueError("Can only specify ratio or num_to_sample or subsampling_param, not several !")
        for name, value in given.items():  # precedence: num_to_sample, ratio, subsampling_param
            if value is not None:
                setattr(self, name, value)
                return
        raise Exception('At least ["ratio, num_to_sample, subsampling_param"] should be defined')

    def __call__(self, pos, x=None, batch=None):
        return self.sample(pos, batch=batch, x=x)

    def _get_num_to_sample(self, batch_size) -> int:
        """the fixed count, or floor(ratio * size)"""
        fixed = getattr(self, "_num_to_sample", None)
        return fixed if fixed is not None else math.floor(batch_size * self._ratio)

    def _get_ratio_to_sample(self, batch_size) -> float:
        """the given ratio, or count / size"""
        ratio = getattr(self, "_ratio", None)
        return ratio if ratio is not None else self._num_to_sample / float(batch_size)

    @abstractmethod
    def sample(self, pos, x=None, batch=None):
        """indices of the sampled points"""


class FPSSampler(BaseSampler):
    """sampling.py:53-63 -> torch_geometric.nn.fps.  `random_start` defaults to the reference's
    behaviour (True); parity tests and benchmarks pass False (start index 0 of each cloud)."""

    def __init__(self, ratio=None, num_to_sample=None, subsampling_param=None, random_start=True):
        super().__init__(ratio, num_to_sample, subsampling_param)
        self._random_start = random_start

    def sample(self, pos, batch, **kwargs):
        _expect_rank(pos, 2, _SPARSE_MSG)
        share = self._get_ratio_to_sample(pos.shape[0])
        return ops.fps(pos, batch, ratio=share, random_start=self._random_start)


class DenseFPSSampler(BaseSampler):
    """sampling.py:84-100 -> tp.furthest_point_sample; pos [B,N,3] -> idx [B,num_sample]."""

    def sample(self, pos, **kwargs):
        _expect_rank(pos, 3, _DENSE_MSG)
        return ops.furthest_point_sample(pos, self._get_num_to_sample(pos.shape[1]))


class RandomSampler(BaseSampler):
    """sampling.py:103-112 (RandLA-Net style): draws with replacement over the whole batch.  `sorted_indices=True`
    returns the draw in ascending order, so that `batch[idx]` stays sorted -- which the batched searches downstream
    require (torch_cluster's knn states the same precondition; the reference feeds it the unsorted draw)."""

    def __init__(self, ratio=None, num_to_sample=None, subsampling_param=None, sorted_indices=False):
        super().__init__(ratio, num_to_sample, subsampling_param)
        self._sorted = sorted_indices

    def sample(self, pos, batch=None, **kwargs):
        _expect_rank(pos, 2, _SPARSE_MSG)
        total = pos.shape[0]
        draw = torch.randint(0, total, (self._get_num_to_sample(total),), device=pos.device)
        return draw.sort().values if self._sorted else draw


class DenseRandomSampler(BaseSampler):
    """sampling.py:115-127: one draw shared by every cloud of the batch."""

    def sample(self, pos, **kwargs):
        _expect_rank(pos, 3, _DENSE_MSG)
        per_cloud = pos.shape[1]
        return torch.randint(0, per_cloud, (self._get_num_to_sample(per_cloud),), device=pos.device)


# --------------------------------------------------- neighbour finders (neighbour_finder.py) ---
class BaseNeighbourFinder(ABC):
    def __call__(self, x, y, batch_x, batch_y):
        return self.find_neighbours(x, y, batch_x, batch_y)

    @abstractmethod
    def find_neighbours(self, x, y, batch_x, batch_y):
        """neighbours of every y among x"""

    def __repr__(self):
        return "%s %s" % (type(self).__name__, self.__dict__)


class RadiusNeighbourFinder(BaseNeighbourFinder):
    """neighbour_finder.py:25-39.  MESSAGE_PASSING -> edge list; anything else is routed to
    ball_query(mode=conv_type) exactly as the reference's always-true `elif` does (:34)."""

    def __init__(self, radius: float, max_num_neighbors: int = 64, conv_type=MESSAGE_PASSING):
        self._radius, self._max_num_neighbors = radius, max_num_neighbors
        self._conv_type = conv_type.lower()

    def find_neighbours(self, x, y, batch_x=None, batch_y=None):
        r, cap = self._radius, self._max_num_neighbors
        if self._conv_type == MESSAGE_PASSING:
            return ops.radius(x, y, r, batch_x, batch_y, max_num_neighbors=cap)
        table, _ = ops.ball_query(r, cap, x, y, mode=self._conv_type, batch_x=batch_x, batch_y=batch_y,
                                  return_dist=False)
        return table


class KNNNeighbourFinder(BaseNeighbourFinder):
    """neighbour_finder.py:42-47 -> knn edge list [2, E] (row = y index, col = x index)."""

    def __init__(self, k):
        self.k = k

    def find_neighbours(self, x, y, batch_x, batch_y):
        return ops.knn_edges(x, y, self.k, batch_x, batch_y)


class DilatedKNNNeighbourFinder(BaseNeighbourFinder):
    """neighbour_finder.py:50-68: the k * dilation nearest per query, of which k are drawn at random (with
    replacement, one independent draw per query)."""

    def __init__(self, k, dilation):
        self.k, self.dilation = k, dilation
        self.initialFinder = KNNNeighbourFinder(k * dilation)

    def find_neighbours(self, x, y, batch_x, batch_y):
        row, col = self.initialFinder.find_neighbours(x, y, batch_x, batch_y)
        wide, n = self.k * self.dilation, len(y)
        pick = torch.randint(wide, (n, self.k), device=row.device, dtype=torch.long)
        pick += wide * torch.arange(n, dtype=torch.long, device=row.device).unsqueeze(1)  # offset of query i's block
        pick = pick.reshape(-1)
        return row[pick], col[pick]


class BaseMSNeighbourFinder(ABC):
    def __call__(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        return self.find_neighbours(x, y, batch_x=batch_x, batch_y=batch_y, scale_idx=scale_idx)

    @abstractmethod
    def find_neighbours(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        """neighbours at one of the scales"""

    @property
    @abstractmethod
    def num_scales(self):
        """number of (radius, max_num_neighbors) pairs"""


class MultiscaleRadiusNeighbourFinder(BaseMSNeighbourFinder):
    """neighbour_finder.py:89-153: a scalar on one side is repeated to the length of the list on the other side;
    two lists must have the same length; two scalars make one scale."""

    def __init__(self, radius: Union[float, List[float]], max_num_neighbors: Union[int, List[int]] = 64):
        many_r, many_k = _is_list(radius), _is_list(max_num_neighbors)
        if many_r and many_k and len(radius) != len(max_num_neighbors):
            raise ValueError("Both lists max_num_neighbors and radius should be of the same length")
        scales = len(radius) if many_r else (len(max_num_neighbors) if many_k else 1)
        self._radius = list(radius) if many_r else [radius] * scales
        self._max_num_neighbors = list(max_num_neighbors) if many_k else [max_num_neighbors] * scales

    def _scale(self, scale_idx):
        if scale_idx >= self.num_scales:
            raise ValueError("Scale %i is out of bounds %i" % (scale_idx, self.num_scales))
        return self._radius[scale_idx], self._max_num_neighbors[scale_idx]

    def find_neighbours(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        r, cap = self._scale(scale_idx)
        return ops.radius(x, y, r, batch_x, batch_y, max_num_neighbors=cap)

    @property
    def num_scales(self):
        return len(self._radius)

    def __call__(self, x, y, batch_x=None, batch_y=None, scale_idx=0):
        return self.find_neighbours(x, y, batch_x, batch_y, scale_idx)


class DenseRadiusNeighbourFinder(MultiscaleRadiusNeighbourFinder):
    """neighbour_finder.py:156-178 -> tp.ball_query(r, K, x, y)[0], idx [B,S,K] int64."""

    def find_neighbours(self, x, y, scale_idx=0):
        r, cap = self._scale(scale_idx)
        table, _ = ops.ball_query(r, cap, x, y, return_dist=False)
        return table

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
            b = data.pos.new_zeros((data.pos.shape[0],), dtype=torch.long)
        return b

    def precompute(self, query, support):
        """the k query points nearest to every support point with their 1 / d^2 weights and the per-support-point sum
        of the weights (interpolate.py:11-33), to be applied by __call__(..., precomputed=...)"""
        src, dst = query.pos, support.pos
        with torch.no_grad():
            y_idx, x_idx = ops.knn_edges(src, dst, self.k, batch_x=self._batch(query), batch_y=self._batch(support))
            d2 = (src[x_idx] - dst[y_idx]).square().sum(dim=-1, keepdim=True)
            weights = 1.0 / d2.clamp(min=1e-16)
        normalisation = weights.new_zeros((dst.shape[0], 1)).index_add_(0, y_idx, weights)
        return SimpleNamespace(num_nodes=dst.shape[0], x_idx=x_idx, y_idx=y_idx, weights=weights,
                               normalisation=normalisation)

    def __call__(self, query, support, precomputed=None):
        if not precomputed:
            return ops.knn_interpolate(query.x, query.pos, support.pos, self._batch(query), self._batch(support),
                                       k=self.k)
        n = support.pos.shape[0]
        if n != precomputed.num_nodes:
            raise ValueError("Precomputed indices do not match with the data given to the transform")
        feats = query.x
        weighted = feats[precomputed.x_idx] * precomputed.weights
        return feats.new_zeros((n, feats.shape[1])).index_add_(0, precomputed.y_idx, weighted) / precomputed.normalisation
