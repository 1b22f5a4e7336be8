"""Deformable KPConv through the public API against (1) golden vectors produced by the
reference's own KPConv_deform_ops + torch autograd (tests/golden/kpdeform_*.npz) and (2) the numpy oracle on larger
seeded shapes.  Bars: sq_distances bit-exact (fp32, reference operation order); outputs and gradients within 2e-5 of
max|reference| (fp32 sums in a different order; feature gradients are fp32 reductions at L2)."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-5
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("name", ["kpdeform_linear_mod", "kpdeform_linear", "kpdeform_gaussian_mod", "kpdeform_constant"])
def test_deform_against_reference_golden(name):
    from torch_points3d_b200 import ops
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    mod = "modulations" in g.files
    ext, infl = float(g["extent"]), str(g["influence"])
    x, W, off = (T(g[k]).requires_grad_(True) for k in ("x", "W", "offsets"))
    mods = T(g["modulations"]).requires_grad_(True) if mod else None
    out, sqd, deformed = ops.KPConv_deform_ops(T(g["query"]), T(g["support"]), T(g["neighbors"]), x, T(g["K_points"]),
                                               off, mods, W, ext, infl, "sum")
    assert relerr(N(out), g["out"]) <= TOL
    assert np.array_equal(N(sqd), g["sqd"]) and np.array_equal(N(deformed), g["deformed"])
    # the loss of tests/golden/make_golden.py: all three outputs, fitting term through torch.min over the neighbours
    kpmin = sqd.min(dim=1)[0]
    fitting = torch.mean(kpmin / (1.5 * ext) ** 2)
    assert abs(float(fitting.detach()) - float(g["fitting"])) <= 1e-5 * abs(float(g["fitting"]))
    loss = (out * T(g["dout"])).sum() + 3.0 * fitting + (sqd * T(g["dsq"])).sum() + (deformed * T(g["ddef"])).sum()
    loss.backward()
    assert relerr(N(x.grad), g["dx"]) <= TOL
    assert relerr(N(W.grad), g["dW"]) <= TOL
    assert relerr(N(off.grad), g["doffsets"]) <= TOL
    if mod:
        assert relerr(N(mods.grad), g["dmodulations"]) <= TOL

    # fused fitting path: kpmin + argmin instead of the [n,M,K] tensor, same gradients without the dense cotangent
    x2, W2, off2 = (T(g[k]).requires_grad_(True) for k in ("x", "W", "offsets"))
    mods2 = T(g["modulations"]).requires_grad_(True) if mod else None
    out2, kpmin2, deformed2 = ops.KPConv_deform_fitting(T(g["query"]), T(g["support"]), T(g["neighbors"]), x2,
                                                        T(g["K_points"]), off2, mods2, W2, ext, infl, "sum")
    assert torch.equal(out2, out) and torch.equal(kpmin2, kpmin)
    x3, W3, off3 = (T(g[k]).requires_grad_(True) for k in ("x", "W", "offsets"))
    mods3 = T(g["modulations"]).requires_grad_(True) if mod else None
    out3, sqd3, _ = ops.KPConv_deform_ops(T(g["query"]), T(g["support"]), T(g["neighbors"]), x3, T(g["K_points"]),
                                          off3, mods3, W3, ext, infl, "sum")
    ((out2 * T(g["dout"])).sum() + 3.0 * torch.mean(kpmin2 / (1.5 * ext) ** 2)).backward()
    ((out3 * T(g["dout"])).sum() + 3.0 * torch.mean(sqd3.min(dim=1)[0] / (1.5 * ext) ** 2)).backward()
    assert relerr(N(off2.grad), N(off3.grad)) <= 1e-6 and relerr(N(x2.grad), N(x3.grad)) <= TOL


@pytest.mark.parametrize("n,N0,M,Cin,Cout,infl,mod", [(3000, 5000, 34, 32, 64, "linear", True),
                                                     (700, 900, 40, 70, 33, "gaussian", False),
                                                     (1000, 1000, 7, 5, 128, "constant", True),
                                                     (257, 300, 100, 16, 16, "linear", False)])
def test_deform_vs_oracle(n, N0, M, Cin, Cout, infl, mod):
    from torch_points3d_b200 import kpconv, ops
    rng = np.random.default_rng(n)
    sup = (rng.random((N0, 3)) * 0.5).astype(np.float32)
    qry = sup[rng.permutation(N0)[:n]] + (0.002 * rng.standard_normal((n, 3))).astype(np.float32)
    d = ((qry[:, None] - sup[None]) ** 2).sum(-1)
    nb = np.argsort(d, 1)[:, :M].astype(np.int64)
    nb[rng.random((n, M)) < 0.15] = N0
    nb[1] = N0
    ext = 0.05
    Kp = kpconv.load_kernels(1.5 * ext, random_rotation=False).astype(np.float32)
    x = rng.standard_normal((N0, Cin)).astype(np.float32)
    W = (rng.standard_normal((15, Cin, Cout)) * 0.1).astype(np.float32)
    off = (rng.standard_normal((n, 15, 3)) * 0.3 * ext).astype(np.float32)
    mods = (2 / (1 + np.exp(-rng.standard_normal((n, 15))))).astype(np.float32) if mod else None
    dout = rng.standard_normal((n, Cout)).astype(np.float32)
    dsq = (rng.standard_normal((n, M, 15)) * (nb != N0)[:, :, None]).astype(np.float32)
    ddef = rng.standard_normal((n, 15, 3)).astype(np.float32)
    tx, tW, toff = T(x).requires_grad_(True), T(W).requires_grad_(True), T(off).requires_grad_(True)
    tm = T(mods).requires_grad_(True) if mod else None
    out, sqd, deformed = ops.KPConv_deform_ops(T(qry), T(sup), T(nb), tx, T(Kp), toff, tm, tW, ext, infl, "sum")
    want_out, want_sqd, want_def = oracle.kpconv_deform_fwd(qry, sup, nb, x, Kp, off, mods, W, ext, infl)
    assert relerr(N(out), want_out) <= TOL
    assert np.array_equal(N(sqd), want_sqd) and np.array_equal(N(deformed), want_def)
    ((out * T(dout)).sum() + (sqd * T(dsq)).sum() + (deformed * T(ddef)).sum()).backward()
    dx, dW, doff, dm = oracle.kpconv_deform_bwd(qry, sup, nb, x, Kp, off, mods, W, ext, infl, dout, dsq, ddef)
    assert relerr(N(tx.grad), dx) <= TOL and relerr(N(tW.grad), dW) <= TOL and relerr(N(toff.grad), doff) <= TOL
    if mod:
        assert relerr(N(tm.grad), dm) <= TOL
    assert float(out[1].abs().max()) == 0.0  # all-shadow row


def test_deformable_layer_and_errors():
    from torch_points3d_b200 import kpconv, ops, synthetic
    torch.manual_seed(0)
    pos_h, batch_h = synthetic.indoor_spheres(num_spheres=2, raw_points=6000, radius=0.5, seed=5)
    pos, batch = T(pos_h), T(batch_h)
    nb = ops.ball_query(0.1, 30, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch)[0]
    for modulated, loss_mode in ((False, "fitting"), (True, "fitting"), (True, "permissive")):
        layer = kpconv.KPConvDeformableLayer(16, 32, point_influence=0.04, modulated=modulated, loss_mode=loss_mode,
                                             random_rotation=False).to(DEV)
        x = torch.randn(pos.shape[0], 16, device=DEV, requires_grad=True)
        y = layer(pos, pos, nb, x)
        assert y.shape == (pos.shape[0], 32) and torch.isfinite(y).all()
        losses = layer.get_internal_losses()
        total = y.square().mean()
        for key in (("fitting_loss", "repulsion_loss") if loss_mode == "fitting" else ("permissive_loss",)):
            if torch.is_tensor(losses[key]) and torch.isfinite(losses[key]):
                total = total + losses[key]
        total.backward()
        for name, p in layer.named_parameters():
            if p.requires_grad:
                assert p.grad is not None and torch.isfinite(p.grad).all(), name
        assert float(layer.offset_weights.grad.abs().sum()) > 0 and float(x.grad.abs().sum()) > 0
    kp = layer.K_points
    off = torch.zeros(pos.shape[0], 15, 3, device=DEV)
    W = torch.zeros(15, 16, 8, device=DEV)
    x = torch.zeros(pos.shape[0], 16, device=DEV)
    with pytest.raises(TypeError):
        ops.KPConv_deform_ops(pos, pos, nb, x, kp, off, None, W, 0.04, "linear", "closest")
    with pytest.raises(ValueError):
        ops.KPConv_deform_ops(pos, pos, nb, x, kp, off, None, W, 0.04, "cubic", "sum")
    with pytest.raises(RuntimeError):
        ops.KPConv_deform_ops(pos.cpu(), pos.cpu(), nb.cpu(), x.cpu(), kp.cpu(), off.cpu(), None, W.cpu(), 0.04,
                              "linear", "sum")
    # zero offsets, no modulation == the rigid operator (linear influence)
    xr = torch.randn(pos.shape[0], 16, device=DEV)
    Wr = torch.randn(15, 16, 8, device=DEV) * 0.1
    rigid = ops.KPConv_ops(pos, pos, nb, xr, kp, Wr, 0.04, "linear", "sum", precision="fp32")
    deform = ops.KPConv_deform_ops(pos, pos, nb, xr, kp, off, None, Wr, 0.04, "linear", "sum")[0]
    assert relerr(N(deform), N(rigid)) <= 1e-5
