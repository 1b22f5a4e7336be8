"""The reference's MultiScaleBatch wire format:

  MultiScaleTransform                 core/data_transform/transforms.py:579-651   per-sample pyramid (`multiscale`, `upsample`)
  MultiScaleBatch.from_data_list      datasets/multiscale_data.py:56-104           collation of the per-sample pyramids
  from_data_list_token                datasets/multiscale_data.py:107-165          index shifting rule of that collation

so that a pyramid made the reference's way -- per sample, collated afterwards -- can be fed to the blocks'
`precomputed=` path (modules/KPConv/blocks.py:67-81) and to FPModule_PD (core/base_conv/partial_dense.py:121-131),
and so that this package can emit that format.  The per-sample searches run on libp3d_b200.so (ops.py); the collation
is tensor bookkeeping (torch.cat + offsets) wherever the tensors live.

Wire format after collation, per level of `multiscale`: `pos [n,3]`, `batch [n]` (sample id), `idx_neighboors [n,M]`
int64 where every non-negative entry of sample i was shifted by the number of SUPPORT points of samples 0..i-1.  The
shift is applied to the per-sample shadow index as well (it is non-negative), so the shadow of sample i ends up equal
to the first support point of sample i+1; only the last sample's shadow is the real `N_support` (Appendix C.4 of the
survey: a property of the reference's collation, reproduced here because the tables are an INPUT of the operators).
Per level of `upsample` (innermost first): `x_idx`, `y_idx` shifted by the cumulated query / support counts,
`weights`, `normalisation`, `batch`, `num_nodes`.
"""
import re

import torch

from .partial_dense import _Data


class MultiScaleData(_Data):
    """datasets/multiscale_data.py:9-53"""

    @property
    def num_scales(self):
        ms = getattr(self, "multiscale", None)
        return len(ms) if ms else 0

    @property
    def num_upsample(self):
        up = getattr(self, "upsample", None)
        return len(up) if up else 0

    def clone(self):
        out = MultiScaleData()
        out.__dict__.update(self.__dict__)
        return out

    def apply(self, func, *keys):
        """datasets/multiscale_data.py:21-35: `func` on every tensor attribute (or on `keys`) and on every tensor of
        every level of both lists"""
        for key in (keys or list(self.keys)):
            item = getattr(self, key, None)
            if torch.is_tensor(item):
                setattr(self, key, func(item))
        for levels in (getattr(self, "multiscale", None) or [], getattr(self, "upsample", None) or []):
            for level in levels:
                for key, item in list(level.__dict__.items()):
                    if torch.is_tensor(item):
                        setattr(level, key, func(item))
        return self

    def to(self, device, *keys, **kwargs):
        return self.apply(lambda t: t.to(device, **kwargs), *keys)

    def contiguous(self, *keys):
        return self.apply(lambda t: t.contiguous(), *keys)


def _num_nodes(data):
    n = getattr(data, "num_nodes", None)
    if n is not None:
        return int(n)
    for key in ("x", "pos", "norm", "batch"):  # torch_geometric Data.num_nodes
        item = getattr(data, key, None)
        if torch.is_tensor(item):
            return item.shape[0]
    return None


def _inc(data, key, item):
    """Data.__inc__ with the per-object overrides MultiScaleTransform installs (transforms.py:595-603,631-647)"""
    special = getattr(data, "_inc", None) or {}
    if key in special:
        return special[key]
    return _num_nodes(data) if re.search("(index|face)", key) else 0


def from_data_list_token(data_list):
    """datasets/multiscale_data.py:107-165: torch_geometric's Batch.from_data_list except that negative entries are
    never shifted.  (The reference shifts the samples' tensors in place; this returns new tensors.)"""
    keys = sorted(set().union(*[set(k for k in d.keys if not k.startswith("_") and k != "num_nodes") for d in data_list]))
    assert "batch" not in keys
    cols = {key: [] for key in keys}
    cumsum = {key: 0 for key in keys}
    batch_vec, total_nodes, have_nodes = [], 0, True
    for i, data in enumerate(data_list):
        for key in keys:
            item = getattr(data, key, None)
            if item is None:
                continue
            if torch.is_tensor(item) and item.dtype != torch.bool and cumsum[key] > 0:
                item = torch.where(item >= 0, item + cumsum[key], item)
            cumsum[key] += _inc(data, key, item)
            cols[key].append(item)
        n = _num_nodes(data)
        if n is None:
            have_nodes = False
        else:
            dev = next((v.device for v in data.__dict__.values() if torch.is_tensor(v)), None)
            batch_vec.append(torch.full((n,), i, dtype=torch.long, device=dev))
            total_nodes += n
    out = _Data()
    for key, items in cols.items():
        if not items:
            continue
        if torch.is_tensor(items[0]):
            dim = -1 if re.search("(index|face)", key) else 0
            setattr(out, key, torch.cat(items, dim=dim).contiguous())
        elif isinstance(items[0], (int, float)):
            setattr(out, key, torch.tensor(items))
        else:
            raise ValueError("Unsupported attribute type {} : {}".format(type(items[0]), items[0]))
    out.batch = torch.cat(batch_vec) if have_nodes and batch_vec else None
    if have_nodes:
        out.num_nodes = total_nodes
    return out


class MultiScaleBatch(MultiScaleData):
    @staticmethod
    def from_data_list(data_list):
        """datasets/multiscale_data.py:56-104"""
        num_scales = data_list[0].num_scales
        num_upsample = data_list[0].num_upsample
        for d in data_list:
            assert d.num_scales == num_scales, "All data objects should contain the same number of scales"
            assert d.num_upsample == num_upsample, "All data objects should contain the same number of scales"
        multiscale = [from_data_list_token([d.multiscale[s] for d in data_list]) for s in range(num_scales)]
        upsample = [from_data_list_token([d.upsample[s] for d in data_list]) for s in range(num_upsample)]
        plain = []
        for d in data_list:
            c = _Data(**{k: v for k, v in d.__dict__.items() if k not in ("multiscale", "upsample")})
            plain.append(c)
        flat = from_data_list_token(plain)
        batch = MultiScaleBatch()
        batch.__dict__.update(flat.__dict__)
        batch.multiscale, batch.upsample = multiscale, upsample
        return batch


class MultiScaleTransform:
    """core/data_transform/transforms.py:579-651, one SAMPLE at a time like the reference's DataLoader workers (the
    whole-batch on-device variant is kpconv.MultiScaleTransform).  `strategies` = {"sampler": [...],
    "neighbour_finder": [...], "upsample_op": [...]} as the model hands them out (KPConvPaper.strategies())."""

    def __init__(self, strategies):
        self.strategies = strategies
        self.num_layers = len(strategies["sampler"])

    @torch.no_grad()
    def __call__(self, data):
        ms_data = MultiScaleData(**data.__dict__)
        pos = data.pos.contiguous()
        zeros = lambda n: torch.zeros((n,), dtype=torch.long, device=pos.device)  # noqa: E731
        precomputed = [_Data(pos=pos)]
        upsample, upsample_index = [], 0
        ups = self.strategies.get("upsample_op") or []
        for index in range(self.num_layers):
            sampler, finder = self.strategies["sampler"][index], self.strategies["neighbour_finder"][index]
            support = precomputed[index]
            if sampler is not None:
                q_pos, _ = sampler(support.pos, zeros(support.pos.shape[0]))
                query = _Data(pos=q_pos.contiguous())
                if len(ups):
                    if upsample_index >= len(ups):
                        raise ValueError("You are missing some upsample blocks in your network")
                    pre_up = ups[upsample_index].precompute(query, support)
                    upsample_index += 1
                    pre = _Data(x_idx=pre_up.x_idx, y_idx=pre_up.y_idx, weights=pre_up.weights,
                                normalisation=pre_up.normalisation, num_nodes=pre_up.num_nodes)
                    pre._inc = {"x_idx": query.pos.shape[0], "y_idx": support.pos.shape[0]}
                    upsample.append(pre)
            else:
                query = _Data(pos=support.pos)
            s_pos, q_pos = support.pos, query.pos
            query.idx_neighboors = finder(s_pos, q_pos, batch_x=zeros(s_pos.shape[0]), batch_y=zeros(q_pos.shape[0]))
            query._inc = {"idx_neighboors": s_pos.shape[0]}
            precomputed.append(query)
        ms_data.multiscale = precomputed[1:]
        upsample.reverse()  # innermost level first
        ms_data.upsample = upsample
        return ms_data
