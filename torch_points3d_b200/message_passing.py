"""Host-side mirror of the MESSAGE_PASSING convolutions: PointConv as
modules/pointnet2/message_passing.py:9-26 uses it (torch_geometric 1.7.2 `PointConv`, aggr="max"), RandlaKernel
(modules/RandLANet/modules.py:9-54, aggr="add") and the down-convolution chain of
core/base_conv/message_passing.py:60-93, plus the single-scale chain BaseConvolutionDown.forward (:44-58) as RandLA-Net
drives it (RandlaConv, DilatedResidualBlock, RandLANetRes: modules/RandLANet/modules.py:57-130), FPModule (:146-165)
and GlobalBaseModule (:120-141).  The per-edge gathers and the per-target reductions are kernels of
libp3d_b200.so (csrc/edges.cu); the MLPs in between are stock torch modules, as in the reference.

Edge lists: `edge_index = [source j (support index), target i (query index)]`, sorted by target, exactly what
`torch.stack([col, row])` of the reference chains builds from `radius` / `knn` output.
"""
import torch
import torch.nn.functional as F

from . import ops
from .partial_dense import MLP, _Data, _global_pool, copy_from_to
from .spatial_ops import FPSSampler, KNNNeighbourFinder, MultiscaleRadiusNeighbourFinder, RandomSampler


class PointConv(torch.nn.Module):
    """torch_geometric.nn.PointConv (1.7.2): out_i = global_nn(max_j local_nn([x_j, pos_j - pos_i])).
    `add_self_loops` reproduces the library's handling, which the reference inherits by default: edges with
    source index == target index are removed and (i, i) is appended for i < min(#source, #target) -- on a
    bipartite (support, query) pair this links query i to support i, whatever those points are."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=True):
        super().__init__()
        self.local_nn, self.global_nn, self.add_self_loops = local_nn, global_nn, add_self_loops

    def forward(self, x, pos, edge_index):
        if isinstance(x, tuple):
            x = x[0]
        pos_src, pos_dst = (pos, pos) if torch.is_tensor(pos) else pos
        src, dst = edge_index[0], edge_index[1]
        T = pos_dst.shape[0]
        loops = 0
        if self.add_self_loops:
            keep = src != dst
            src, dst = src[keep], dst[keep]
            loops = min(pos_src.shape[0], pos_dst.shape[0])
            ar = torch.arange(loops, device=src.device, dtype=src.dtype)
            src_all, dst_all = torch.cat([src, ar]), torch.cat([dst, ar])
        else:
            src_all, dst_all = src, dst
        msg = ops.edge_message(x, pos_src, pos_dst, src_all, dst_all)
        if self.local_nn is not None:
            msg = self.local_nn(msg)
        E = src.shape[0]
        out = ops.segment_max(msg[:E], dst, T)
        if loops:
            has = torch.bincount(dst, minlength=T)[:loops] > 0
            loop_msg = msg[E:]
            head = torch.where(has.unsqueeze(1), torch.maximum(out[:loops], loop_msg), loop_msg)
            out = torch.cat([head, out[loops:]])
        if self.global_nn is not None:
            out = self.global_nn(out)
        return out


class RandlaKernel(torch.nn.Module):
    """modules/RandLANet/modules.py:9-54: local spatial encoding + attentive pooling, aggr="add"."""

    def __init__(self, point_pos_nn=None, attention_nn=None, global_nn=None, *args, **kwargs):
        super().__init__()
        self.point_pos_nn = MLP(point_pos_nn)
        self.attention_nn = MLP(attention_nn)
        self.global_nn = MLP(global_nn)

    def forward(self, x, pos, edge_index):
        pos_src, pos_dst = (pos, pos) if torch.is_tensor(pos) else pos
        src, dst = edge_index[0], edge_index[1]
        rel = ops.edge_encoding(pos_src, pos_dst, src, dst)              # [pos_i, pos_j, vij, dij]
        rij = self.point_pos_nn(rel)
        x_j = pos_src[src] if x is None else x[src]
        fij_hat = torch.cat([x_j, rij], dim=1)
        g_fij = self.attention_nn(fij_hat)
        s_ij = F.softmax(g_fij, -1)
        msg = s_ij * fij_hat
        return self.global_nn(ops.segment_sum(msg, dst, pos_dst.shape[0]))


class SAModule(torch.nn.Module):
    """modules/pointnet2/message_passing.py:9-37 over BaseMSConvolutionDown.forward
    (core/base_conv/message_passing.py:60-93): FPS -> one radius edge list per scale -> shared PointConv -> concat."""

    def __init__(self, ratio=None, radius=None, radius_num_point=None, down_conv_nn=None, random_start=True, **kwargs):
        super().__init__()
        self.sampler = FPSSampler(ratio=ratio, random_start=random_start)
        self.neighbour_finder = MultiscaleRadiusNeighbourFinder(radius, max_num_neighbors=radius_num_point)
        local_nn = MLP(down_conv_nn) if down_conv_nn is not None else None
        self._conv = PointConv(local_nn=local_nn, global_nn=None)

    def conv(self, x, pos, edge_index, batch):
        return self._conv(x, pos, edge_index)

    def forward(self, data, **kwargs):
        x, pos, batch = data.x, data.pos, data.batch
        idx = self.sampler(pos, batch)
        centres, centre_batch = pos[idx], batch[idx]
        per_scale = []
        for scale_idx in range(self.neighbour_finder.num_scales):
            row, col = self.neighbour_finder(pos, centres, batch_x=batch, batch_y=centre_batch, scale_idx=scale_idx)
            per_scale.append(self.conv(x, (pos, centres), torch.stack([col, row], dim=0), batch))
        out = _Data(idx=idx, x=torch.cat(per_scale, -1), pos=centres, batch=centre_batch)
        copy_from_to(data, out)  # labels and any other attribute travel with the level (message_passing.py:90)
        return out


class BaseConvolutionDown(torch.nn.Module):
    """core/base_conv/message_passing.py:35-58: idx = sampler(pos, batch); (row, col) = neighbour_finder(pos, pos[idx],
    batch, batch[idx]); edge_index = [col, row] = [support index, query index]; conv over that bipartite graph.

    The reference passes the position pair to `conv` as (pos[idx], pos) (:52) although edge_index[0] indexes `pos` and
    edge_index[1] indexes `pos[idx]` -- its multi-scale sibling (:86) and every PointNet++ MP module pass (pos, pos[idx]).
    With the pair swapped torch_geometric looks support indices up in the QUERY set: an out-of-range read as soon as
    the sampler keeps fewer points than it was given (RandlaConv with ratio < 1), a silent mix-up when it keeps all
    of them.  This mirror passes (source positions, target positions) = (pos, pos[idx]); `reference_pos_order=True`
    reproduces the literal argument order for the ratio-1 case.

    Sampler call: the reference writes `self.sampler(pos, batch)` (:47), and BaseSampler.__call__ is
    `(pos, x=None, batch=None)` (core/spatial_ops/sampling.py:33), so the batch vector lands in the `x` slot and an
    FPSSampler there samples the whole batch as ONE cloud (RandomSampler ignores `batch` either way).  This chain
    passes `batch=batch`: every cloud is sampled on its own and the sampled batch vector stays sorted, which the
    neighbour search needs.  SAModule below keeps the reference's literal call."""

    def __init__(self, sampler, neighbour_finder, *args, reference_pos_order=False, **kwargs):
        super().__init__()
        self.sampler = sampler
        self.neighbour_finder = neighbour_finder
        self._index = kwargs.get("index", None)
        self._reference_pos_order = reference_pos_order

    def conv(self, x, pos, edge_index, batch):
        raise NotImplementedError

    def forward(self, data, **kwargs):
        batch_obj = _Data()
        x, pos, batch = data.x, data.pos, data.batch
        idx = self.sampler(pos, batch=batch)
        row, col = self.neighbour_finder(pos, pos[idx], batch_x=batch, batch_y=batch[idx])
        edge_index = torch.stack([col, row], dim=0)
        batch_obj.idx = idx
        batch_obj.edge_index = edge_index
        pair = (pos[idx], pos) if self._reference_pos_order else (pos, pos[idx])
        batch_obj.x = self.conv(x, pair, edge_index, batch)
        batch_obj.pos = pos[idx]
        batch_obj.batch = batch[idx]
        copy_from_to(data, batch_obj)
        return batch_obj


class RandlaConv(BaseConvolutionDown):
    """modules/RandLANet/modules.py:57-67: RandomSampler(ratio) -> KNNNeighbourFinder(k) -> RandlaKernel.  The random
    draw is sorted (see spatial_ops.RandomSampler) so the sampled batch vector stays sorted for the knn search; pass
    `sampler=` (e.g. an FPSSampler) to drive the chain with another sampling strategy."""

    def __init__(self, ratio=None, k=None, *args, sampler=None, **kwargs):
        super().__init__(sampler if sampler is not None else RandomSampler(ratio, sorted_indices=True), KNNNeighbourFinder(k), *args,
                         **kwargs)
        if kwargs.get("index") == 0 and kwargs.get("nb_feature") is not None:
            kwargs["point_pos_nn"][-1] = kwargs.get("nb_feature")
            kwargs["attention_nn"][0] = kwargs["attention_nn"][-1] = kwargs.get("nb_feature") * 2
            kwargs["down_conv_nn"][0] = kwargs.get("nb_feature") * 2
        self._conv = RandlaKernel(*args, global_nn=kwargs["down_conv_nn"], **kwargs)

    def conv(self, x, pos, edge_index, batch):
        return self._conv(x, pos, edge_index)


class BaseResnetBlock(torch.nn.Module):
    """core/base_conv/message_passing.py:212-257"""

    def __init__(self, indim, outdim, convdim):
        super().__init__()
        self.indim, self.outdim, self.convdim = indim, outdim, convdim
        self.features_downsample_nn = MLP([self.indim, self.outdim // 4])
        self.features_upsample_nn = MLP([self.convdim, self.outdim])
        self.shortcut_feature_resize_nn = MLP([self.indim, self.outdim])
        self.activation = torch.nn.ReLU()

    def convs(self, data):
        raise NotImplementedError

    def forward(self, data, **kwargs):
        batch_obj = _Data()
        x = data.x
        shortcut = x
        x = self.features_downsample_nn(x)  # (computed and dropped by the reference as well: `convs` reads data.x)
        data = self.convs(data)
        x = data.x
        idx = data.idx
        x = self.features_upsample_nn(x)
        if idx is not None:
            shortcut = shortcut[idx]
        shortcut = self.shortcut_feature_resize_nn(shortcut)
        x = shortcut + x
        batch_obj.x = x
        batch_obj.pos = data.pos
        batch_obj.batch = data.batch
        copy_from_to(data, batch_obj)
        return batch_obj


class DilatedResidualBlock(BaseResnetBlock):
    """modules/RandLANet/modules.py:70-100: two RandlaConv (k = 16, hard-coded) inside a residual block.  Note that
    the block's shortcut gathers with conv2's sample indices only (`data.idx` of the last conv), as the reference does."""

    def __init__(self, indim, outdim, ratio1, ratio2, point_pos_nn1, point_pos_nn2, attention_nn1, attention_nn2,
                 global_nn1, global_nn2, *args, **kwargs):
        if kwargs.get("index") == 0 and kwargs.get("nb_feature") is not None:
            indim = kwargs.get("nb_feature")
        super().__init__(indim, outdim, outdim)
        self.conv1 = RandlaConv(ratio1, 16, *args, point_pos_nn=point_pos_nn1, attention_nn=attention_nn1,
                                down_conv_nn=global_nn1, **kwargs)
        kwargs["nb_feature"] = None
        self.conv2 = RandlaConv(ratio2, 16, *args, point_pos_nn=point_pos_nn2, attention_nn=attention_nn2,
                                down_conv_nn=global_nn2, **kwargs)

    def convs(self, data):
        data = self.conv1(data)
        data = self.conv2(data)
        return data


class RandLANetRes(torch.nn.Module):
    """modules/RandLANet/modules.py:103-127"""

    def __init__(self, indim, outdim, ratio, point_pos_nn, attention_nn, down_conv_nn, *args, **kwargs):
        super().__init__()
        self._conv = DilatedResidualBlock(indim, outdim, ratio[0], ratio[1], point_pos_nn[0], point_pos_nn[1],
                                          attention_nn[0], attention_nn[1], down_conv_nn[0], down_conv_nn[1], *args,
                                          **kwargs)

    def forward(self, data):
        return self._conv.forward(data)


class GlobalBaseModule(torch.nn.Module):
    """core/base_conv/message_passing.py:120-141: MLP over cat(x, pos), per-cloud max / mean pool"""

    def __init__(self, nn, aggr="max", *args, **kwargs):
        super().__init__()
        self.nn = MLP(nn)
        self.aggr = aggr

    def forward(self, data, num_clouds=None, **kwargs):
        batch_obj = _Data()
        x, pos, batch = data.x, data.pos, data.batch
        x = self.nn(torch.cat([x, pos], dim=1)) if pos is not None else self.nn(x)
        if num_clouds is None:
            num_clouds = int(batch[-1].item()) + 1  # sizes the output (global_max_pool reads it too)
        x = _global_pool(x, batch, num_clouds, self.aggr)
        batch_obj.x = x
        if pos is not None:
            batch_obj.pos = pos.new_zeros((x.size(0), 3))
        batch_obj.batch = torch.arange(x.size(0), device=batch.device)
        copy_from_to(data, batch_obj)
        return batch_obj


class FPModule(torch.nn.Module):
    """core/base_conv/message_passing.py:146-165 over BaseConvolutionUp.forward (:96-127): knn_interpolate the coarse
    level onto the skip level (libp3d_b200's knn), concatenate the skip features, MLP."""

    def __init__(self, up_k, up_conv_nn, *args, **kwargs):
        super().__init__()
        self.k = up_k
        # the reference's FPModule builds its base class without its keywords (`super().__init__(None)`, :165), so a
        # `skip=False` passed to it never arrives: the skip features are always concatenated
        self._skip = True
        bn_momentum = kwargs.get("bn_momentum", 0.1)
        self.nn = MLP(up_conv_nn, bn_momentum=bn_momentum, bias=False)

    def conv(self, x, pos, pos_skip, batch, batch_skip, *args):
        return ops.knn_interpolate(x, pos, pos_skip, batch, batch_skip, k=self.k)

    def forward(self, data, **kwargs):
        batch_obj = _Data()
        data, data_skip = data
        x, pos, batch = data.x, data.pos, data.batch
        x_skip, pos_skip, batch_skip = data_skip.x, data_skip.pos, data_skip.batch
        x = self.conv(x, pos, pos_skip, batch, batch_skip, None)
        if x_skip is not None and self._skip:
            x = torch.cat([x, x_skip], dim=1)
        batch_obj.x = self.nn(x) if hasattr(self, "nn") else x
        copy_from_to(data_skip, batch_obj)
        return batch_obj
