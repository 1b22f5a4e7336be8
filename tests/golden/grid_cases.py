"""Cases shared by tests/golden/make_golden_grid.py (run on the REFERENCE's own grid_transform.py) and
tests/test_reference_grid_transform.py (run on torch_points3d_b200.grid_transform with the device operators replaced
by the stand-ins below): which attribute gets the mean, the majority vote, the representative point; bool and integer
attributes; untouched attributes; quantised coordinates; `last` mode after the seeded shuffle; list inputs; error
types and messages.

The third-party operators are the same restatements on both sides (torch_cluster 1.5.9 grid_cluster / torch_geometric
1.7.2 voxel_grid + consecutive_cluster / torch_scatter 2.0.8 scatter_add + scatter_mean): what is pinned is the host
logic of the transform, not their numerics."""
import numpy as np
import torch


def voxel_key(coords, batch=None, size=None):
    """grid_cluster / voxel_grid with cell size 1: x fastest, the batch id as a 4th coordinate"""
    full = coords if batch is None else torch.cat([coords, batch.to(coords.dtype).unsqueeze(1)], dim=1)
    start, end = full.min(0).values, full.max(0).values
    cells = (end - start).long() + 1
    strides = torch.cat([torch.ones(1, dtype=torch.long), torch.cumprod(cells, 0)[:-1]])
    return ((full - start).long() * strides).sum(1)


def consecutive_cluster(src):
    """consecutive ids in increasing key order; the LAST point of a cluster is its representative"""
    _, inv = torch.unique(src, sorted=True, return_inverse=True)
    perm = np.empty(int(inv.max()) + 1, dtype=np.int64)
    perm[inv.numpy()] = np.arange(inv.shape[0], dtype=np.int64)  # in order: the last write wins
    return inv, torch.from_numpy(perm)


def scatter_add(src, index, dim=0, dim_size=None):
    n = int(index.max()) + 1 if dim_size is None else dim_size
    return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, src)


def scatter_mean(src, index, dim=0, dim_size=None):
    total = scatter_add(src, index, dim, dim_size)
    count = torch.bincount(index, minlength=total.shape[0]).clamp(min=1).reshape((-1,) + (1,) * (src.dim() - 1))
    if src.dtype.is_floating_point:
        return total / count.to(src.dtype)
    return torch.div(total, count, rounding_mode="floor")


def _np(t):
    return t.detach().clone().numpy() if torch.is_tensor(t) else np.asarray(t)


def _attempt(fn):
    try:
        fn()
        return "no error"
    except Exception as e:  # noqa: BLE001
        return "{}: {}".format(type(e).__name__, e)


def _cloud(ns, n, seed, batch=True):
    g = torch.Generator().manual_seed(seed)
    kw = dict(pos=torch.randn(n, 3, generator=g) * 0.1, x=torch.randn(n, 4, generator=g),
              y=torch.randint(3, 8, (n,), generator=g), instance_labels=torch.randint(-1, 3, (n,), generator=g),
              mask=torch.rand(n, generator=g) > 0.4, rgb=torch.randint(0, 256, (n, 3), generator=g),
              other=torch.arange(7), name="cloud")
    if batch:
        kw["batch"] = (torch.arange(n) >= n // 2).long()
    return ns.Data(**kw)


def _record(out, tag, data):
    keys = []
    for key, item in data:
        keys.append(key)
        if torch.is_tensor(item):
            out["%s.%s" % (tag, key)] = _np(item)
            out["%s.%s.dtype" % (tag, key)] = np.asarray(str(item.dtype))
        else:
            out["%s.%s" % (tag, key)] = np.asarray(item)
    out[tag + ".keys"] = np.asarray(",".join(sorted(keys)))


def run_cases(ns):
    out = {}
    save = ns.SaveOriginalPosId()
    _record(out, "mean_batch", ns.GridSampling3D(0.05, quantize_coords=True)(save(_cloud(ns, 60, 1))))
    _record(out, "mean_single", ns.GridSampling3D(0.08)(_cloud(ns, 40, 2, batch=False)))
    torch.manual_seed(3)
    _record(out, "last", ns.GridSampling3D(0.05, mode="last", quantize_coords=True)(save(_cloud(ns, 50, 3))))
    both = ns.GridSampling3D(0.1)([_cloud(ns, 20, 4), _cloud(ns, 30, 5, batch=False)])
    out["list.len"] = np.asarray(len(both))
    _record(out, "list0", both[0])
    _record(out, "list1", both[1])
    twice = ns.GridSampling3D(0.05)
    d = twice(twice(_cloud(ns, 60, 6)))
    _record(out, "twice", d)
    out["repr"] = np.asarray(repr(ns.GridSampling3D(0.02, quantize_coords=True, mode="last")))
    # SaveOriginalPosId: an existing attribute is kept, lists are mapped
    kept = _cloud(ns, 10, 7)
    kept.origin_id = torch.arange(10) * 2
    out["save.kept"] = _np(save(kept).origin_id)
    out["save.list"] = _np(save([_cloud(ns, 5, 8)])[0].origin_id)
    out["save.repr"] = np.asarray(repr(save))
    # group_data called directly
    base = _cloud(ns, 30, 9)
    cluster, unique = consecutive_cluster(voxel_key(torch.round(base.pos / 0.1), base.batch))
    _record(out, "direct_mean", ns.group_data(_cloud(ns, 30, 9), cluster, unique, mode="mean"))
    _record(out, "direct_last_skip", ns.group_data(_cloud(ns, 30, 9), cluster, unique, mode="last", skip_keys=["x", "y"]))
    _record(out, "direct_mean_skip", ns.group_data(_cloud(ns, 30, 9), cluster, None, mode="mean",
                                                  skip_keys=["batch"]))
    out["err.mode"] = np.asarray(_attempt(lambda: ns.group_data(_cloud(ns, 30, 9), cluster, unique, mode="first")))
    out["err.mean"] = np.asarray(_attempt(lambda: ns.group_data(_cloud(ns, 30, 9), None, unique, mode="mean")))
    out["err.last"] = np.asarray(_attempt(lambda: ns.group_data(_cloud(ns, 30, 9), cluster, None, mode="last")))
    edged = _cloud(ns, 30, 9)
    edged.edge_attr = torch.zeros(30, 2)
    out["err.edge"] = np.asarray(_attempt(lambda: ns.group_data(edged, cluster, unique, mode="mean")))
    return out
