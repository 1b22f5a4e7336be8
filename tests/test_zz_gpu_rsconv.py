"""GPU test of the dense RSConv down modules (torch_points3d_b200/rsconv.py): their feature preparation on
libp3d_b200.so against the CPU oracle, bit for bit, and one training step through each module.  The arithmetic of the
modules themselves is pinned on the CPU against the reference's own classes (tests/test_reference_rsconv.py)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import oracle
from torch_points3d_b200 import _cabi, rsconv, synthetic
from torch_points3d_b200.partial_dense import _Data

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def test_rsconv_feature_preparation_and_step():
    import __graft_entry__
    __graft_entry__.build()
    before = _cabi.launch_count()
    B, n, S, C = 4, 1024, 128, 6
    radii, nsample = [0.15, 0.3], [16, 32]
    xyz = synthetic.surface_clouds(B, n, seed=77)
    feat = np.random.default_rng(77).standard_normal((B, C, n)).astype(np.float32)
    pos, x = T(xyz), T(feat)
    oidx = oracle.fps_dense(xyz, S)
    new_xyz = np.take_along_axis(xyz, oidx[..., None], 1)
    centre = new_xyz.transpose(0, 2, 1)[..., None]
    xyz_t = np.ascontiguousarray(xyz.transpose(0, 2, 1))

    orig = rsconv.RSConvOriginalMSGDown(npoint=S, radii=radii, nsample=nsample, down_conv_nn=[[10, 16, 32], [C + 3, 32]],
                                        channel_raising_nn=[32, 64]).to(DEV)
    shared = rsconv.RSConvSharedMSGDown(npoint=S, radii=radii, nsample=nsample, down_conv_nn=[[10, 16, 32], [C + 3, 32]],
                                        channel_raising_nn=[32, 64]).to(DEV)
    assert np.array_equal(N(orig.sampler(pos)), oidx)
    for s, (r, k) in enumerate(zip(radii, nsample)):
        table = orig.neighbour_finder(pos, T(new_xyz), scale_idx=s)
        otable = oracle.ball_query_dense(r, k, xyz, new_xyz)[0]
        assert np.array_equal(N(table), otable)
        rel = oracle.group_fwd(xyz_t, otable) - centre
        gf = oracle.group_fwd(feat, otable)
        # RSConvOriginalMSGDown: centred coordinates twice (the reference centres in place), fused and op by op
        want = np.concatenate([rel, rel, gf], axis=1)
        for fused in (True, False):
            orig.fused_grouping = fused
            got = orig._prepare_features(pos, T(new_xyz), x, table)
            assert got.shape == (B, 3 + 3 + C, S, k) and np.array_equal(N(got), want), (s, fused)
        # RSConvSharedMSGDown / RSConvMSGDown: absolute and centred coordinates
        got, cen = shared._prepare_features(x, pos, T(new_xyz), table)
        assert np.array_equal(N(got), np.concatenate([oracle.group_fwd(xyz_t, otable), rel, gf], axis=1))
        assert np.array_equal(N(cen), centre)
    orig.fused_grouping = True
    msg = rsconv.RSConvMSGDown(npoint=S, radii=radii, nsample=nsample, down_conv_nn=[10, 16, C + 3],
                               channel_raising_nn=[C + 3, 48]).to(DEV)
    for module, width in ((orig, 64), (shared, 64), (msg, 48)):
        xin = x.clone().requires_grad_(True)
        out = module(NS(pos=pos, x=xin))
        assert out.x.shape == (B, width * len(radii), S) and torch.isfinite(out.x).all()
        assert np.array_equal(N(out.pos), new_xyz)
        out.x.square().mean().backward()
        assert xin.grad is not None and float(xin.grad.abs().sum()) > 0
        assert all(p.grad is not None for p in module.parameters())
    assert _cabi.launch_count() > before, "no kernel of libp3d_b200.so was launched"


def test_rsconv_message_passing_down_module():
    """RSConvDown (FPS -> radius edge list -> relation-shape convolution with a per-target max) on the device against
    the plain torch formulation of modules/RSConv/message_passing.py:29-52 on the same sampled points and edges."""
    import __graft_entry__
    __graft_entry__.build()
    before = _cabi.launch_count()
    torch.manual_seed(0)
    pos_h, batch_h = synthetic.indoor_spheres(num_spheres=3, raw_points=3000, radius=0.5, seed=9)
    pos, batch = T(pos_h), T(batch_h)
    C = 6
    x = torch.randn(pos.shape[0], C, device=DEV, requires_grad=True)
    down = rsconv.RSConvDown(ratio=0.25, radius=0.2, local_nn=[10, 16, C], down_conv_nn=[C, 32, 48]).to(DEV).eval()
    out = down(_Data(x=x, pos=pos, batch=batch))
    T_ = out.idx.shape[0]
    assert out.x.shape == (T_, 48) and torch.isfinite(out.x).all()
    assert torch.equal(out.pos, pos[out.idx]) and torch.equal(out.batch, batch[out.idx])
    src, dst = out.edge_index[0], out.edge_index[1]
    assert bool((dst[1:] >= dst[:-1]).all()) and bool((batch[src] == batch[out.idx][dst]).all())
    pos_i, pos_j = pos[out.idx][dst], pos[src]
    vij = pos_i - pos_j
    assert float(vij.norm(dim=1).max()) <= 0.2 * (1 + 1e-5)
    hij = torch.cat([vij.norm(dim=1, keepdim=True), vij, pos_i, pos_j], dim=1)
    msg = down._conv.local_nn(hij) * x[src]
    pooled = torch.zeros(T_, C, device=DEV).scatter_reduce(0, dst.unsqueeze(1).expand(-1, C), msg, "amax",
                                                          include_self=False)
    want = down._conv.global_nn(torch.relu(pooled))
    assert torch.allclose(out.x, want, rtol=1e-4, atol=1e-5)
    out.x.square().mean().backward()
    assert float(x.grad.abs().sum()) > 0 and all(p.grad is not None for p in down.parameters())
    assert _cabi.launch_count() > before, "no kernel of libp3d_b200.so was launched"


def test_rsconv_msn_network_step():
    """one training step of the whole RSConv_MSN network on the device (its arithmetic is pinned on the CPU against the
    reference's own network code, tests/test_reference_rsconv_net.py): shapes, finite logits, a gradient for every
    parameter, the fused and the op-by-op grouping paths agree."""
    import __graft_entry__
    __graft_entry__.build()
    before = _cabi.launch_count()
    torch.manual_seed(0)
    B, n = 2, 2048
    pos = T(synthetic.surface_clouds(B, n, seed=78))
    x = torch.randn(B, n, 3, device=DEV)
    y = torch.randint(0, 50, (B, n), device=DEV)
    y[:, ::31] = -1
    model = rsconv.RSConv_MSN(3, 50).to(DEV)
    logits = model(pos, x)
    assert logits.shape == (B * n, 50) and torch.isfinite(logits).all()
    loss = model.loss(logits, y)
    loss.backward()
    assert torch.isfinite(loss) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    with torch.no_grad():
        fused = model(pos, x)
        for m in model.down_modules:
            m.fused_grouping = False
        plain = model(pos, x)
    assert torch.allclose(fused, plain, rtol=1e-4, atol=1e-4 * float(plain.abs().max()))
    assert _cabi.launch_count() > before, "no kernel of libp3d_b200.so was launched"
