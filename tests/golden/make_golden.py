"""Generate golden vectors by running the REFERENCE's own code in this container.

Only the KPConv aggregation is importable offline: it is
loaded BY FILE PATH from /root/reference (nothing is copied), executed on CPU
fp32 tensors, and inputs + outputs + autograd gradients are stored as small
.npz fixtures next to this script.  /root/reference does not exist on the GPU
box, so tests only ever read the committed .npz files.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import struct
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_reference_kpconv_ops():
    # convolution_ops.py imports torch_points3d.core.common_modules.gathering; provide
    # exactly that module (also by file path) without importing the package __init__s.
    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    for pkg in ["torch_points3d", "torch_points3d.core", "torch_points3d.core.common_modules"]:
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    load("torch_points3d.core.common_modules.gathering",
         os.path.join(REF, "torch_points3d/core/common_modules/gathering.py"))
    return load("ref_convolution_ops", os.path.join(REF, "torch_points3d/modules/KPConv/convolution_ops.py"))


def read_kernel_ply(path):
    """k_015_center.ply: ASCII header, then 15 x 3 little-endian float64."""
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    body = raw[end:]
    n = len(body) // 24
    return np.array(struct.unpack("<%dd" % (3 * n), body[: 24 * n]), dtype=np.float64).reshape(n, 3)


def make_case(ops, seed, n, N0, M, Cin, Cout, influence, aggregation, extent=0.05, shadow_frac=0.2, K=15):
    g = torch.Generator().manual_seed(seed)
    kp = read_kernel_ply(os.path.join(REF, "torch_points3d/modules/KPConv/kernels/dispositions/k_015_center.ply"))
    K_points = torch.from_numpy((kp[:K] * 1.5 * extent).astype(np.float32))
    support = torch.rand((N0, 3), generator=g) * 0.3
    qsel = torch.randperm(N0, generator=g)[:n]
    query = support[qsel] + 0.002 * torch.randn((n, 3), generator=g)
    # neighbours: nearest ones (so that influences are non-trivial), then random shadows
    d = torch.cdist(query, support)
    nb = d.topk(M, largest=False).indices
    mask = torch.rand((n, M), generator=g) < shadow_frac
    nb = torch.where(mask, torch.full_like(nb, N0), nb)
    nb[0, :] = N0  # one all-shadow row -> exactly 0 (linear / gaussian)
    x = torch.randn((N0, Cin), generator=g)
    W = torch.randn((K, Cin, Cout), generator=g) * 0.1
    dout = torch.randn((n, Cout), generator=g)

    x.requires_grad_(True)
    W.requires_grad_(True)
    out = ops.KPConv_ops(query, support, nb.clone(), x, K_points, W, extent, influence, aggregation)
    out.backward(dout)
    return dict(
        query=query.numpy(), support=support.numpy(), neighbors=nb.numpy().astype(np.int64),
        x=x.detach().numpy(), K_points=K_points.numpy(), W=W.detach().numpy(), dout=dout.numpy(),
        extent=np.float32(extent), influence=influence, aggregation=aggregation,
        out=out.detach().numpy(), dx=x.grad.numpy(), dW=W.grad.numpy(),
    )


def make_deform_case(ops, seed, n, N0, M, Cin, Cout, influence, modulated, extent=0.05, shadow_frac=0.2, K=15):
    """KPConv_deform_ops (convolution_ops.py:110-235) forward + autograd gradients w.r.t. features, K_values,
    offsets and modulations, for a loss that uses all three outputs the way KPConvDeformableLayer does
    (kernels.py:228-246: new_feat, fitting_loss(sq_distances), repulsion over deformed_K_points) plus a dense
    cotangent on sq_distances restricted to the real (non-shadow) neighbours."""
    g = torch.Generator().manual_seed(seed)
    kp = read_kernel_ply(os.path.join(REF, "torch_points3d/modules/KPConv/kernels/dispositions/k_015_center.ply"))
    K_points = torch.from_numpy((kp[:K] * 1.5 * extent).astype(np.float32))
    support = torch.rand((N0, 3), generator=g) * 0.3
    qsel = torch.randperm(N0, generator=g)[:n]
    query = support[qsel] + 0.002 * torch.randn((n, 3), generator=g)
    d = torch.cdist(query, support)
    nb = d.topk(M, largest=False).indices
    mask = torch.rand((n, M), generator=g) < shadow_frac
    nb = torch.where(mask, torch.full_like(nb, N0), nb)
    nb[0, :] = N0
    x = torch.randn((N0, Cin), generator=g)
    W = torch.randn((K, Cin, Cout), generator=g) * 0.1
    offsets = torch.randn((n, K, 3), generator=g) * (0.3 * extent)
    mods = (2 * torch.sigmoid(torch.randn((n, K), generator=g))) if modulated else None
    dout = torch.randn((n, Cout), generator=g)
    dsq = torch.randn((n, M, K), generator=g) * (nb != N0).unsqueeze(-1).float()
    ddef = torch.randn((n, K, 3), generator=g)
    for t in (x, W, offsets) + ((mods,) if modulated else ()):
        t.requires_grad_(True)
    out, sqd, deformed = ops.KPConv_deform_ops(query, support, nb.clone(), x, K_points, offsets, mods, W, extent,
                                               influence, "sum")
    kpmin = sqd.min(dim=1)[0]
    fitting = torch.mean(kpmin / (1.5 * extent) ** 2)
    loss = (out * dout).sum() + 3.0 * fitting + (sqd * dsq).sum() + (deformed * ddef).sum()
    loss.backward()
    case = dict(
        query=query.numpy(), support=support.numpy(), neighbors=nb.numpy().astype(np.int64), x=x.detach().numpy(),
        K_points=K_points.numpy(), W=W.detach().numpy(), offsets=offsets.detach().numpy(), dout=dout.numpy(),
        dsq=dsq.numpy(), ddef=ddef.numpy(), extent=np.float32(extent), influence=influence,
        out=out.detach().numpy(), sqd=sqd.detach().numpy(), deformed=deformed.detach().numpy(),
        fitting=np.float32(fitting.item()), dx=x.grad.numpy(), dW=W.grad.numpy(), doffsets=offsets.grad.numpy(),
    )
    if modulated:
        case["modulations"] = mods.detach().numpy()
        case["dmodulations"] = mods.grad.numpy()
    return case


def main():
    ops = _load_reference_kpconv_ops()
    cases = [
        ("kpconv_linear_sum", dict(seed=11, n=96, N0=160, M=20, Cin=8, Cout=12, influence="linear", aggregation="sum")),
        ("kpconv_gaussian_sum", dict(seed=12, n=64, N0=128, M=16, Cin=5, Cout=7, influence="gaussian", aggregation="sum")),
        ("kpconv_constant_sum", dict(seed=13, n=64, N0=128, M=16, Cin=6, Cout=6, influence="constant", aggregation="sum")),
        ("kpconv_linear_closest", dict(seed=14, n=64, N0=128, M=18, Cin=8, Cout=8, influence="linear", aggregation="closest")),
        ("kpconv_linear_sum_c32", dict(seed=15, n=200, N0=300, M=34, Cin=32, Cout=32, influence="linear", aggregation="sum")),
    ]
    for name, kw in cases:
        case = make_case(ops, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **case)
        print(name, case["out"].shape, float(np.abs(case["out"]).mean()))
    deform_cases = [
        ("kpdeform_linear_mod", dict(seed=21, n=96, N0=160, M=20, Cin=8, Cout=12, influence="linear", modulated=True)),
        ("kpdeform_linear", dict(seed=22, n=120, N0=200, M=34, Cin=32, Cout=32, influence="linear", modulated=False)),
        ("kpdeform_gaussian_mod", dict(seed=23, n=64, N0=128, M=16, Cin=5, Cout=7, influence="gaussian", modulated=True)),
        ("kpdeform_constant", dict(seed=24, n=64, N0=128, M=16, Cin=6, Cout=6, influence="constant", modulated=False)),
    ]
    for name, kw in deform_cases:
        case = make_deform_case(ops, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **case)
        print(name, case["out"].shape, float(np.abs(case["out"]).mean()), float(np.abs(case["doffsets"]).mean()))
    # the deformable-KPConv regularisers (modules/KPConv/losses.py), values and gradients
    spec = importlib.util.spec_from_file_location("ref_kpconv_losses",
                                                  os.path.join(REF, "torch_points3d/modules/KPConv/losses.py"))
    losses = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(losses)
    g = torch.Generator().manual_seed(31)
    deformed = (torch.randn((40, 15, 3), generator=g) * 0.9).requires_grad_(True)
    sqd = (torch.rand((40, 20, 15), generator=g) * 0.01).requires_grad_(True)
    vals = {"fitting": losses.fitting_loss(sqd, 0.06), "repulsion": losses.repulsion_loss(deformed, 0.04),
            "permissive": losses.permissive_loss(deformed, 1.2)}
    grads = {"fitting": torch.autograd.grad(vals["fitting"], sqd)[0],
             "repulsion": torch.autograd.grad(vals["repulsion"], deformed)[0],
             "permissive": torch.autograd.grad(vals["permissive"], deformed)[0]}
    np.savez_compressed(os.path.join(HERE, "kpconv_losses.npz"), deformed=deformed.detach().numpy(),
                        sqd=sqd.detach().numpy(), **{k: np.float32(v.item()) for k, v in vals.items()},
                        **{"d" + k: v.numpy() for k, v in grads.items()})
    print("losses", {k: float(v) for k, v in vals.items()})
    kp = read_kernel_ply(os.path.join(REF, "torch_points3d/modules/KPConv/kernels/dispositions/k_015_center.ply"))
    # the 15 kernel-point disposition is DATA of the reference model (k_015_center.ply); kept as a fixture
    np.save(os.path.join(HERE, "k_015_center.npy"), kp)
    print("kernel points", kp.shape)


if __name__ == "__main__":
    main()
