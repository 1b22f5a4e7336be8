"""MESSAGE_PASSING operators against the numpy oracle: edge gathers bit-exact, segment
max bit-exact, segment sum bit-exact (same edge order), gradients against torch autograd of the plain formulation."""
import numpy as np
import pytest
import torch

import oracle

DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def _edges(rng, n_src, n_dst, max_deg):
    deg = rng.integers(0, max_deg + 1, n_dst)
    deg[rng.integers(0, n_dst, max(1, n_dst // 10))] = 0  # targets without edges
    dst = np.repeat(np.arange(n_dst), deg)
    src = rng.integers(0, n_src, dst.shape[0])
    return src.astype(np.int64), dst.astype(np.int64)


def test_oracle_segment_ops_small_known_answers():
    msg = np.array([[1, -5], [3, -7], [-2, -1]], np.float32)
    dst = np.array([0, 0, 2])
    assert oracle.segment_max(msg, dst, 4).tolist() == [[3, -5], [0, 0], [-2, -1], [0, 0]]
    assert oracle.segment_sum(msg, dst, 3).tolist() == [[4, -12], [0, 0], [-2, -1]]
    enc = oracle.edge_encoding(np.zeros((1, 3), np.float32), np.array([[3, 4, 0]], np.float32), np.array([0]), np.array([0]))
    assert enc.tolist() == [[3, 4, 0, 0, 0, 0, 3, 4, 0, 5]]


@pytest.mark.gpu
@pytest.mark.parametrize("n_src,n_dst,C,deg", [(5000, 1200, 64, 40), (300, 300, 3, 9), (1000, 50, 130, 64), (64, 2000, 1, 3)])
def test_edge_ops_against_oracle(n_src, n_dst, C, deg):
    from torch_points3d_b200 import ops
    rng = np.random.default_rng(n_src + C)
    src, dst = _edges(rng, n_src, n_dst, deg)
    ps, pd = rng.standard_normal((n_src, 3)).astype(np.float32), rng.standard_normal((n_dst, 3)).astype(np.float32)
    x = rng.standard_normal((n_src, C)).astype(np.float32)
    tx = T(x).requires_grad_(True)
    msg = ops.edge_message(tx, T(ps), T(pd), T(src), T(dst))
    assert np.array_equal(N(msg), oracle.edge_message(x, ps, pd, src, dst))
    assert np.array_equal(N(ops.edge_message(None, T(ps), T(pd), T(src), T(dst))), oracle.edge_message(None, ps, pd, src, dst))
    assert np.array_equal(N(ops.edge_encoding(T(ps), T(pd), T(src), T(dst))), oracle.edge_encoding(ps, pd, src, dst))
    g = rng.standard_normal(msg.shape).astype(np.float32)
    msg.backward(T(g))
    want = np.zeros_like(x)
    np.add.at(want, src, g[:, :C])
    assert np.abs(N(tx.grad) - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)
    # reductions
    m = rng.standard_normal((src.shape[0], C)).astype(np.float32)
    for fn, ofn in ((ops.segment_max, oracle.segment_max), (ops.segment_sum, oracle.segment_sum)):
        tm = T(m).requires_grad_(True)
        out = fn(tm, T(dst), n_dst)
        assert np.array_equal(N(out), ofn(m, dst, n_dst)), fn.__name__
        go = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(T(go))
        tm2 = T(m).requires_grad_(True)
        if fn is ops.segment_sum:
            ref = torch.zeros(n_dst, C, device=DEV).index_add(0, T(dst), tm2)
        else:
            ref = torch.zeros(n_dst, C, device=DEV).scatter_reduce(0, T(dst).unsqueeze(1).expand(-1, C), tm2, "amax",
                                                                   include_self=False)
        ref.backward(T(go))
        assert torch.equal(tm.grad, tm2.grad) or np.abs(N(tm.grad) - N(tm2.grad)).max() == 0, fn.__name__


@pytest.mark.gpu
def test_point_conv_and_randla_modules():
    from torch_points3d_b200 import message_passing as mp, ops, synthetic
    torch.manual_seed(0)
    pos_h, batch_h = synthetic.indoor_spheres(num_spheres=3, raw_points=3000, radius=0.5, seed=9)
    pos, batch = T(pos_h), T(batch_h)
    x = torch.randn(pos.shape[0], 6, device=DEV, requires_grad=True)
    data = mp._Data(x=x, pos=pos, batch=batch)
    sa = mp.SAModule(ratio=0.25, radius=[0.1, 0.2], radius_num_point=[16, 32], down_conv_nn=[6 + 3, 32, 64],
                     random_start=False).to(DEV).eval()
    out = sa(data)
    assert out.x.shape == (out.idx.shape[0], 128) and torch.isfinite(out.x).all()
    assert torch.equal(out.pos, pos[out.idx]) and torch.equal(out.batch, batch[out.idx])
    # scale 0 against the plain formulation of PointConv (message -> local_nn -> scatter max with the library's loops)
    row, col = sa.neighbour_finder(pos, pos[out.idx], batch_x=batch, batch_y=batch[out.idx], scale_idx=0)
    keep = col != row
    n = out.idx.shape[0]
    ar = torch.arange(n, device=DEV)
    src, dst = torch.cat([col[keep], ar]), torch.cat([row[keep], ar])
    msg = sa._conv.local_nn(torch.cat([x[src], pos[src] - pos[out.idx][dst]], 1))
    want = torch.zeros(n, 64, device=DEV).scatter_reduce(0, dst.unsqueeze(1).expand(-1, 64), msg, "amax", include_self=False)
    assert torch.allclose(out.x[:, :64], want, rtol=0, atol=1e-6)
    out.x.square().mean().backward()
    assert float(x.grad.abs().sum()) > 0
    # RandLA kernel on a knn graph
    q = pos[out.idx]
    row, col = ops.knn_edges(pos, q, 16, batch, batch[out.idx])
    rk = mp.RandlaKernel(point_pos_nn=[10, 8], attention_nn=[14, 14], global_nn=[14, 32]).to(DEV).eval()
    x6 = x.detach()
    y = rk(x6, (pos, q), torch.stack([col, row]))
    rel = torch.cat([q[row], pos[col], q[row] - pos[col], (q[row] - pos[col]).norm(dim=1, keepdim=True)], 1)
    fij = torch.cat([x6[col], rk.point_pos_nn(rel)], 1)
    want = rk.global_nn(torch.zeros(q.shape[0], 14, device=DEV).index_add(0, row, torch.softmax(rk.attention_nn(fij), -1) * fij))
    assert y.shape == (q.shape[0], 32) and torch.allclose(y, want, rtol=1e-4, atol=1e-5)
