"""Golden vectors for the two small modules every block of the path is built from, produced by running the
REFERENCE's own classes in this container: FastBatchNorm1d and MLP of
torch_points3d/core/common_modules/base_modules.py (:35-45, :128-153; the file needs only torch and numpy, so it is
loaded BY FILE PATH from /root/reference -- nothing is copied).  Stored: the state_dict key order of the reference
modules, seeded parameters, a seeded input, and the training-mode output, running statistics and input gradient.
/root/reference does not exist on the GPU box, so tests only read the committed .npz.

    python tests/golden/make_golden_modules.py
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/torch_points3d/core/common_modules/base_modules.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_base_modules", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(20)
    out = {}

    # MLP([7, 12, 5]) on a [N, C] table, two training steps (the second one sees updated running statistics)
    mlp = ref.MLP([7, 12, 5], bn_momentum=0.2, bias=False)
    keys = list(mlp.state_dict().keys())
    out["mlp_keys"] = np.array(keys)
    with torch.no_grad():
        for p in mlp.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.5)
    for k, v in mlp.state_dict().items():
        out["mlp_init/" + k] = v.numpy().copy()
    x = torch.randn((33, 7), generator=g, requires_grad=True)
    mlp.train()
    y1 = mlp(x)
    y2 = mlp(x * 1.5 + 0.25)
    gy = torch.randn(y2.shape, generator=g)
    (gx,) = torch.autograd.grad(y2, x, gy)
    out.update(mlp_x=x.detach().numpy(), mlp_y1=y1.detach().numpy(), mlp_y2=y2.detach().numpy(), mlp_gy=gy.numpy(),
               mlp_gx=gx.numpy())
    for k, v in mlp.state_dict().items():
        out["mlp_after/" + k] = v.numpy().copy()
    mlp.eval()
    out["mlp_eval"] = mlp(x).detach().numpy()

    # FastBatchNorm1d on the dense layout [B, N, C] (the `_forward_dense` branch)
    bn = ref.FastBatchNorm1d(6, momentum=0.05)
    out["bn_keys"] = np.array(list(bn.state_dict().keys()))
    with torch.no_grad():
        bn.batch_norm.weight.copy_(torch.rand(6, generator=g) + 0.5)
        bn.batch_norm.bias.copy_(torch.randn(6, generator=g))
    xd = torch.randn((3, 10, 6), generator=g)
    bn.train()
    out.update(bn_weight=bn.batch_norm.weight.detach().numpy().copy(), bn_bias=bn.batch_norm.bias.detach().numpy().copy(),
               bn_x=xd.numpy(), bn_y=bn(xd).detach().numpy(),
               bn_running_mean=bn.batch_norm.running_mean.numpy().copy(),
               bn_running_var=bn.batch_norm.running_var.numpy().copy())
    try:
        bn(torch.zeros(2, 2, 2, 2))
    except ValueError as e:
        out["bn_error"] = np.array(str(e))
    np.savez_compressed(os.path.join(HERE, "base_modules.npz"), **out)
    print("wrote base_modules.npz:", keys)


if __name__ == "__main__":
    main()
