"""Case list shared by tests/golden/make_golden_hostlogic.py (which runs it on the REFERENCE's own
core/spatial_ops classes, loaded by file path with the third-party kernels replaced by recording stand-ins) and
tests/test_reference_hostlogic.py (which runs it on torch_points3d_b200.spatial_ops with `ops.*` replaced by the same
stand-ins).  Everything compared is HOST logic: constructor rules, attributes, exceptions and their messages, which
kernel is called with which scalars, and the interpolation arithmetic around the knn call.
"""
import math
from types import SimpleNamespace

import torch


class Kernels:
    """recording stand-ins for the third-party kernels; `log` holds (name, scalars...) tuples"""

    def __init__(self):
        self.log = []

    def fps(self, pos, batch=None, ratio=0.5, **kw):
        self.log.append(["fps", float(ratio), list(pos.shape), batch is not None])
        return torch.arange(int(math.ceil(ratio * pos.shape[0])))

    def furthest_point_sample(self, pos, npoint, **kw):
        self.log.append(["furthest_point_sample", int(npoint), list(pos.shape)])
        return torch.arange(npoint).unsqueeze(0).repeat(pos.shape[0], 1)

    def radius(self, x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, **kw):
        self.log.append(["radius", float(r), int(max_num_neighbors), batch_x is not None, batch_y is not None])
        return torch.zeros((2, 1), dtype=torch.long)

    def ball_query(self, radius, nsample, x, y, mode="dense", batch_x=None, batch_y=None, **kw):
        self.log.append(["ball_query", float(radius), int(nsample), str(mode), batch_x is not None, batch_y is not None])
        return torch.zeros(tuple(y.shape[:-1]) + (nsample,), dtype=torch.long), None

    def knn(self, x, y, k, batch_x=None, batch_y=None, **kw):
        """brute force: the k nearest x of every y inside its cloud, ascending distance; [2, E] = (y index, x index)"""
        self.log.append(["knn", int(k), batch_x is not None, batch_y is not None])
        bx = batch_x if batch_x is not None else torch.zeros(x.shape[0], dtype=torch.long)
        by = batch_y if batch_y is not None else torch.zeros(y.shape[0], dtype=torch.long)
        d = torch.cdist(y.double(), x.double())
        d[by.unsqueeze(1) != bx.unsqueeze(0)] = float("inf")
        col = d.topk(k, dim=1, largest=False).indices
        row = torch.arange(y.shape[0]).unsqueeze(1).expand(-1, k)
        return torch.stack([row.reshape(-1), col.reshape(-1)])


def _attempt(fn):
    """result of fn() or the exception it raises, in a JSON-able form"""
    try:
        return {"ok": fn()}
    except Exception as e:  # noqa: BLE001
        kind = type(e).__name__
        return {"raises": kind if kind in ("ValueError", "NotImplementedError", "TypeError") else "Exception",
                "message": str(e)}


def _sampler_state(s):
    return {k: getattr(s, k) for k in ("_num_to_sample", "_ratio", "_subsampling_param") if hasattr(s, k)}


def run_cases(ns, kernels, data_cls=SimpleNamespace):
    """ns: namespace with the spatial_ops classes; kernels: the Kernels instance those classes call into"""
    out = {}
    g = torch.Generator().manual_seed(11)

    # ---- samplers: constructor rule for every combination of the three arguments, count rules
    combos = {}
    for ratio in (None, 0.25):
        for num in (None, 6):
            for sub in (None, 0.1):
                key = "ratio=%s,num=%s,sub=%s" % (ratio, num, sub)
                combos[key] = _attempt(lambda: _sampler_state(
                    ns.DenseFPSSampler(ratio=ratio, num_to_sample=num, subsampling_param=sub)))
    out["sampler_constructor"] = combos
    out["sampler_counts"] = {
        "ratio": [[ns.DenseFPSSampler(ratio=r)._get_num_to_sample(n), ns.DenseFPSSampler(ratio=r)._get_ratio_to_sample(n)]
                  for r in (0.25, 0.5, 0.3) for n in (10, 17, 1024)],
        "num": [[ns.DenseFPSSampler(num_to_sample=m)._get_num_to_sample(n),
                 ns.DenseFPSSampler(num_to_sample=m)._get_ratio_to_sample(n)] for m in (1, 5, 512) for n in (10, 17, 1024)],
    }
    sparse, dense = torch.rand((40, 3), generator=g), torch.rand((2, 20, 3), generator=g)
    batch = torch.arange(2).repeat_interleave(20)
    kernels.log.clear()
    out["sampler_calls"] = {
        "fps_ratio": _attempt(lambda: list(ns.FPSSampler(ratio=0.25).sample(sparse, batch).shape)),
        "fps_num": _attempt(lambda: list(ns.FPSSampler(num_to_sample=8)(sparse, batch=batch).shape)),
        "fps_dense_input": _attempt(lambda: ns.FPSSampler(ratio=0.25).sample(dense, None)),
        "dense_fps_num": _attempt(lambda: list(ns.DenseFPSSampler(num_to_sample=5).sample(dense).shape)),
        "dense_fps_ratio": _attempt(lambda: list(ns.DenseFPSSampler(ratio=0.3)(dense).shape)),
        "dense_fps_sparse_input": _attempt(lambda: ns.DenseFPSSampler(num_to_sample=5).sample(sparse)),
        "random_shape": _attempt(lambda: list(ns.RandomSampler(ratio=0.25).sample(sparse, batch).shape)),
        "random_dense_input": _attempt(lambda: ns.RandomSampler(ratio=0.25).sample(dense, None)),
        "dense_random_shape": _attempt(lambda: list(ns.DenseRandomSampler(num_to_sample=7).sample(dense).shape)),
        "dense_random_sparse_input": _attempt(lambda: ns.DenseRandomSampler(num_to_sample=7).sample(sparse)),
    }
    out["sampler_kernel_log"] = [list(e) for e in kernels.log]

    # ---- neighbour finders
    ms = {}
    for name, (r, k) in {"list_scalar": ([0.1, 0.2, 0.4], 16), "scalar_list": (0.5, [4, 8]),
                         "list_list": ([0.1, 0.2], [4, 8]), "scalar_scalar": (0.3, 12), "default_k": (0.3, None),
                         "mismatch": ([0.1, 0.2, 0.3], [4, 8])}.items():
        def build(r=r, k=k):
            f = ns.MultiscaleRadiusNeighbourFinder(r) if k is None else ns.MultiscaleRadiusNeighbourFinder(r, k)
            return {"radius": list(f._radius), "max_num_neighbors": list(f._max_num_neighbors), "num_scales": f.num_scales}
        ms[name] = _attempt(build)
    out["multiscale_constructor"] = ms
    kernels.log.clear()
    x, y = sparse, sparse[:10]
    bx, by = batch, batch[:10]
    finder_calls = {
        "ms_scale1": _attempt(lambda: list(ns.MultiscaleRadiusNeighbourFinder([0.1, 0.2], [4, 8])(
            x, y, batch_x=bx, batch_y=by, scale_idx=1).shape)),
        "ms_out_of_bounds": _attempt(lambda: ns.MultiscaleRadiusNeighbourFinder([0.1, 0.2], 8).find_neighbours(
            x, y, scale_idx=2)),
        "dense_scale0": _attempt(lambda: list(ns.DenseRadiusNeighbourFinder([0.2, 0.4], [8, 16])(
            dense, dense[:, :5], scale_idx=0).shape)),
        "dense_scale1": _attempt(lambda: list(ns.DenseRadiusNeighbourFinder([0.2, 0.4], [8, 16]).find_neighbours(
            dense, dense[:, :5], scale_idx=1).shape)),
        "dense_out_of_bounds": _attempt(lambda: ns.DenseRadiusNeighbourFinder(0.2, 8).find_neighbours(
            dense, dense[:, :5], scale_idx=1)),
        "radius_mp": _attempt(lambda: list(ns.RadiusNeighbourFinder(0.2, 9)(x, y, bx, by).shape)),
        "radius_partial": _attempt(lambda: list(ns.RadiusNeighbourFinder(0.2, 9, conv_type="PARTIAL_DENSE")(
            x, y, bx, by).shape)),
        "radius_dense": _attempt(lambda: list(ns.RadiusNeighbourFinder(0.2, 9, conv_type="DENSE").find_neighbours(
            dense, dense[:, :5]).shape)),
        "radius_default_k": _attempt(lambda: ns.RadiusNeighbourFinder(0.2)._max_num_neighbors),
        "knn": _attempt(lambda: list(ns.KNNNeighbourFinder(3)(x, y, bx, by).shape)),
        "repr_knn": _attempt(lambda: repr(ns.KNNNeighbourFinder(3))),
        "repr_radius": _attempt(lambda: repr(ns.RadiusNeighbourFinder(0.2, 9, conv_type="DENSE"))),
    }

    def dilated():
        f = ns.DilatedKNNNeighbourFinder(2, 3)
        row, col = f.find_neighbours(x, y, bx, by)
        wide = 6
        full = kernels.knn(x, y, wide, bx, by)  # what the inner finder returned: block i = candidates of query i
        inside = all(int(c) in full[1, int(r) * wide:(int(r) + 1) * wide].tolist() for r, c in zip(row, col))
        return {"rows": row.tolist(), "picked_from_own_block": bool(inside), "inner_k": f.initialFinder.k}
    finder_calls["dilated"] = _attempt(dilated)
    out["finder_calls"] = finder_calls
    out["finder_kernel_log"] = [list(e) for e in kernels.log]

    # ---- KNNInterpolate: the arithmetic around the knn call (weights, normalisation, applying a table)
    q = data_cls(pos=torch.rand((9, 3), generator=g), x=torch.randn((9, 4), generator=g),
                 batch=torch.tensor([0] * 5 + [1] * 4))
    s = data_cls(pos=torch.rand((15, 3), generator=g), x=torch.randn((15, 2), generator=g),
                 batch=torch.tensor([0] * 8 + [1] * 7))
    q.num_nodes, s.num_nodes = 9, 15
    interp = ns.KNNInterpolate(3)
    table = interp.precompute(q, s)
    up = interp(q, s, precomputed=table)
    out["interpolate"] = {
        "num_nodes": int(table.num_nodes), "x_idx": table.x_idx.tolist(), "y_idx": table.y_idx.tolist(),
        "weights": [float(v) for v in table.weights.reshape(-1)],
        "normalisation": [float(v) for v in table.normalisation.reshape(-1)],
        "upsampled": [[float(v) for v in row] for row in up],
        "wrong_size": _attempt(lambda: interp(q, data_cls(pos=torch.rand(14, 3), x=None, num_nodes=14), precomputed=table)),
    }
    return out
