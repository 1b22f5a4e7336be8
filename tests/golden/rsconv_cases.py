"""Case list shared by tests/golden/make_golden_rsconv.py (the REFERENCE's own modules/RSConv/dense.py, loaded by file
path with `torch_points_kernels` replaced by the oracle-backed stand-ins of dense_chain_cases) and
tests/test_reference_rsconv.py (torch_points3d_b200.rsconv with `ops.*` replaced by the same stand-ins): the three dense
RSConv down modules as first and as later layers, with and without input features -- outputs in training and eval
mode, input gradients, running statistics, parameter names."""
import numpy as np
import torch

from dense_chain_cases import seeded_parameters


def _np(t):
    return t.detach().clone().numpy()


def _run(out, tag, module, data_cls, pos, x, seed):
    seeded_parameters(module, seed)
    out[tag + ".names"] = np.asarray(",".join(module.state_dict().keys()))
    module.train()
    p = pos.clone().requires_grad_(True)
    f = None if x is None else x.clone().requires_grad_(True)
    res = module(data_cls(pos=p, x=f))
    out[tag + ".train.x"], out[tag + ".train.pos"] = _np(res.x), _np(res.pos)
    weight = torch.linspace(0.5, 1.5, res.x.numel()).reshape(res.x.shape)
    (res.x * weight).sum().backward()
    out[tag + ".dpos"] = _np(p.grad)
    if f is not None:
        out[tag + ".dx"] = _np(f.grad)
    for name, t in module.state_dict().items():
        if name.endswith("running_mean") or name.endswith("running_var"):
            out[tag + ".stat." + name] = _np(t)
    module.eval()
    with torch.no_grad():
        out[tag + ".eval.x"] = _np(module(data_cls(pos=pos, x=x)).x)


def run_cases(ns, data_cls):
    """ns: RSConvOriginalMSGDown, RSConvSharedMSGDown, RSConvMSGDown (+ `kwargs` for the first one)"""
    out = {}
    g = torch.Generator().manual_seed(11)
    B, N, C = 2, 96, 3
    pos = torch.rand((B, N, 3), generator=g)
    x = torch.randn((B, C, N), generator=g)
    wide = torch.randn((B, 5, N), generator=g)
    common = dict(npoint=24, radii=[0.2, 0.45], nsample=[4, 9])
    extra = getattr(ns, "kwargs", {})
    _run(out, "orig_first", ns.RSConvOriginalMSGDown(down_conv_nn=[[10, 8, 6], [C + 3, 6]], channel_raising_nn=[6, 12],
                                                     **common, **extra), data_cls, pos, x, 1)
    _run(out, "orig_later", ns.RSConvOriginalMSGDown(down_conv_nn=[10, 8, 5 + 3], channel_raising_nn=[5 + 3, 16],
                                                     **common, **extra), data_cls, pos, wide, 2)
    _run(out, "orig_pos_only", ns.RSConvOriginalMSGDown(down_conv_nn=[10, 4, 3], channel_raising_nn=[3, 8],
                                                        **common, **extra), data_cls, pos, None, 3)
    _run(out, "shared_first", ns.RSConvSharedMSGDown(down_conv_nn=[[10, 8, 6], [C + 3, 6]], channel_raising_nn=[6, 12],
                                                     **common), data_cls, pos, x, 4)
    _run(out, "shared_later", ns.RSConvSharedMSGDown(down_conv_nn=[10, 8, 5 + 3], channel_raising_nn=[5 + 3, 10],
                                                     **common), data_cls, pos, wide, 5)
    _run(out, "msg_later", ns.RSConvMSGDown(down_conv_nn=[10, 8, 5 + 3], channel_raising_nn=[5 + 3, 10],
                                            **common), data_cls, pos, wide, 6)
    _run(out, "msg_pos_only", ns.RSConvMSGDown(down_conv_nn=[10, 4, 3], channel_raising_nn=[3, 8], **common),
         data_cls, pos, None, 7)
    return out
