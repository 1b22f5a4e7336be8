"""Host-side mirror of the MESSAGE_PASSING convolutions: PointConv as
modules/pointnet2/message_passing.py:9-26 uses it (torch_geometric 1.7.2 `PointConv`, aggr="max"), RandlaKernel
(modules/RandLANet/modules.py:9-54, aggr="add") and the down-convolution chain of
core/base_conv/message_passing.py:60-93.  The per-edge gathers and the per-target reductions are kernels of
libp3d_b200.so (csrc/edges.cu); the MLPs in between are stock torch modules, as in the reference.

Edge lists: `edge_index = [source j (support index), target i (query index)]`, sorted by target, exactly what
`torch.stack([col, row])` of the reference chains builds from `radius` / `knn` output.
"""
import torch
import torch.nn.functional as F

from . import ops
from .spatial_ops import FPSSampler, MultiscaleRadiusNeighbourFinder


def MLP(channels, activation=None, bn_momentum=0.1, bias=True):
    """core/common_modules/base_modules.py:35-45"""
    activation = activation if activation is not None else torch.nn.LeakyReLU(0.2)
    return torch.nn.Sequential(*[
        torch.nn.Sequential(torch.nn.Linear(channels[i - 1], channels[i], bias=bias),
                            torch.nn.BatchNorm1d(channels[i], momentum=bn_momentum), activation)
        for i in range(1, len(channels))])


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


class _Out:
    pass


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
        out = _Out()
        out.idx = idx
        ms_x = []
        for scale_idx in range(self.neighbour_finder.num_scales):
            row, col = self.neighbour_finder(pos, pos[idx], batch_x=batch, batch_y=batch[idx], scale_idx=scale_idx)
            edge_index = torch.stack([col, row], dim=0)
            ms_x.append(self.conv(x, (pos, pos[idx]), edge_index, batch))
        out.x = torch.cat(ms_x, -1)
        out.pos, out.batch = pos[idx], batch[idx]
        return out
