"""Cases shared by tests/golden/make_golden_multiscale.py (run on the REFERENCE's own MultiScaleTransform /
MultiScaleBatch / from_data_list_token) and tests/test_reference_multiscale.py (run on
torch_points3d_b200.multiscale).  The pyramid wire format of SURVEY section 8f row 4: per-sample precomputation, the
collation's index shifting (support counts for `idx_neighboors`, query / support counts for `x_idx` / `y_idx`, the
per-sample shadow index shifted with the rest, negative entries left alone), the upsample list innermost first.

`ns` provides Data, MultiScaleData, MultiScaleBatch, MultiScaleTransform, from_data_list_token, KNNInterpolate.  The
operators inside are the same brute-force stand-ins on both sides (what is pinned is the format, not the searches)."""
import numpy as np
import torch


class Strided:
    """keeps every `step`-th point; callable the reference's way (Data -> Data) and the mirror's ((pos, batch) ->
    (pos, batch))"""

    def __init__(self, step):
        self.step = step

    def __call__(self, *args):
        if len(args) == 1:
            data = args[0]
            data.pos = data.pos[:: self.step]
            return data
        pos, batch = args
        return pos[:: self.step], batch[:: self.step]


class Ball:
    """partial-dense ball query: lowest support indices inside the radius first, padded with the shadow index N"""

    def __init__(self, radius, width):
        self.radius, self.width = radius, width

    def __call__(self, s_pos, q_pos, batch_x=None, batch_y=None):
        n = s_pos.shape[0]
        d2 = (q_pos[:, None, :] - s_pos[None, :, :]).square().sum(-1)
        inside = (d2 <= self.radius ** 2) & (batch_y[:, None] == batch_x[None, :])
        table = torch.full((q_pos.shape[0], self.width), n, dtype=torch.long)
        for q in range(q_pos.shape[0]):
            hits = torch.nonzero(inside[q]).flatten()[: self.width]
            table[q, : hits.numel()] = hits
        return table


def knn(x, y, k, batch_x=None, batch_y=None, **kw):
    """brute force: the k nearest x of every y inside its cloud, ascending distance; [2, E] = (y index, x index)"""
    bx = batch_x if batch_x is not None else torch.zeros(x.shape[0], dtype=torch.long)
    by = batch_y if batch_y is not None else torch.zeros(y.shape[0], dtype=torch.long)
    d = torch.cdist(y.double(), x.double())
    d[by.unsqueeze(1) != bx.unsqueeze(0)] = float("inf")
    col = d.topk(k, dim=1, largest=False).indices
    row = torch.arange(y.shape[0]).unsqueeze(1).expand(-1, k)
    return torch.stack([row.reshape(-1), col.reshape(-1)])


def _np(t):
    return t.detach().clone().numpy() if torch.is_tensor(t) else np.asarray(t)


def _attempt(fn):
    try:
        fn()
        return "no error"
    except Exception as e:  # noqa: BLE001
        return "{}: {}".format(type(e).__name__, e)


def _strategies(ns, layers=3, ups=2):
    radii = [(0.5, 6), (0.8, 5), (1.2, 4)][:layers]
    return {"sampler": ([None] + [Strided(2)] * (layers - 1)),
            "neighbour_finder": [Ball(r, m) for r, m in radii],
            "upsample_op": [ns.KNNInterpolate(1) for _ in range(ups)]}


def _samples(ns, transform, sizes, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i, n in enumerate(sizes):
        d = ns.Data(pos=torch.rand(n, 3, generator=g), x=torch.rand(n, 2, generator=g),
                    y=torch.full((n,), i, dtype=torch.long))
        out.append(transform(d))
    return out


def run_cases(ns):
    out = {}
    sizes = [11, 7, 9]
    samples = _samples(ns, ns.MultiScaleTransform(_strategies(ns)), sizes, seed=5)
    # ---- per-sample pyramids as the transform leaves them
    for i, ms in enumerate(samples):
        out["s%d.num_scales" % i] = np.asarray(ms.num_scales)
        out["s%d.num_upsample" % i] = np.asarray(ms.num_upsample)
        for l, level in enumerate(ms.multiscale):
            out["s%d.l%d.pos" % (i, l)] = _np(level.pos)
            out["s%d.l%d.idx_neighboors" % (i, l)] = _np(level.idx_neighboors)
        for l, up in enumerate(ms.upsample):
            for key in ("x_idx", "y_idx", "weights", "normalisation"):
                out["s%d.u%d.%s" % (i, l, key)] = _np(getattr(up, key))
            out["s%d.u%d.num_nodes" % (i, l)] = np.asarray(int(up.num_nodes))
    samples[1].multiscale[0].idx_neighboors[0, 0] = -1  # a negative entry is never shifted
    # ---- collation
    batch = ns.MultiScaleBatch.from_data_list(samples)
    out["b.num_scales"] = np.asarray(batch.num_scales)
    out["b.num_upsample"] = np.asarray(batch.num_upsample)
    for key in ("pos", "x", "y", "batch"):
        out["b." + key] = _np(getattr(batch, key))
    for l, level in enumerate(batch.multiscale):
        for key in ("pos", "batch", "idx_neighboors"):
            out["b.l%d.%s" % (l, key)] = _np(getattr(level, key))
    for l, up in enumerate(batch.upsample):
        for key in ("x_idx", "y_idx", "weights", "normalisation", "batch"):
            out["b.u%d.%s" % (l, key)] = _np(getattr(up, key))
        out["b.u%d.num_nodes" % l] = np.asarray(int(up.num_nodes))
    # ---- the token collation on hand-made objects: bool tensors are never shifted, `*index*` keys are joined
    # along their last axis and shifted by the node count, python numbers become one tensor
    hand = []
    g = torch.Generator().manual_seed(9)
    for n in (4, 3, 5):
        hand.append(ns.Data(pos=torch.rand(n, 3, generator=g), edge_index=torch.randint(0, n, (2, n + 1), generator=g),
                            flag=torch.rand(n, generator=g) > 0.5, scale=float(n) / 2, count=n))
    tok = ns.from_data_list_token(hand)
    for key in ("pos", "edge_index", "flag", "scale", "count", "batch"):
        out["tok." + key] = _np(getattr(tok, key))
    # ---- errors
    short = _samples(ns, ns.MultiScaleTransform(_strategies(ns, layers=2, ups=1)), [5], seed=6)
    full = _samples(ns, ns.MultiScaleTransform(_strategies(ns)), [6], seed=7)
    out["err.scales"] = np.asarray(_attempt(lambda: ns.MultiScaleBatch.from_data_list(full + short)))
    out["err.upsample_blocks"] = np.asarray(_attempt(
        lambda: _samples(ns, ns.MultiScaleTransform(_strategies(ns, layers=3, ups=1)), [6], seed=8)))
    # no upsample operators at all: the pyramid is built without an `upsample` list
    bare = _samples(ns, ns.MultiScaleTransform(_strategies(ns, layers=2, ups=0)), [6, 4], seed=10)
    out["bare.num_upsample"] = np.asarray(bare[0].num_upsample)
    bb = ns.MultiScaleBatch.from_data_list(bare)
    out["bare.b.num_upsample"] = np.asarray(bb.num_upsample)
    out["bare.b.l1.idx_neighboors"] = _np(bb.multiscale[1].idx_neighboors)
    return out
