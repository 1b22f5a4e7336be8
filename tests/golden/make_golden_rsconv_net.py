"""Run the REFERENCE's own `RSConv_MSN` network forward + loss on a small input and store the outputs.

Executed from /root/reference by file path (nothing is copied): RSConvOriginalMSGDown (modules/RSConv/dense.py),
GlobalDenseBaseModule / DenseFPModule (core/base_conv/dense.py) built from conf/models/segmentation/rsconv.yaml
(`RSConv_MSN`) with the reference's resolver and `_fetch_arguments_from_list` / `_create_inner_modules` semantics
(models/base_architectures/unet.py:385-444); `set_input` and `forward` of models/segmentation/rsconv.py:51-124 compiled
from the file's text and bound to a bare nn.Module holding `down_modules`, `inner_modules`, `up_modules`, `FC_layer`
(assembled as :35-45 does from the reference's own Seq / Conv1D).  `torch_points_kernels` = oracle-backed stand-ins.

    python tests/golden/make_golden_rsconv_net.py      ->  tests/golden/rsconv_net.npz
"""
import ast
import os
import queue
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import dense_chain_cases  # noqa: E402
import make_golden_dense_chain as dense_loader  # noqa: E402
import make_golden_kpconv_arch as arch  # noqa: E402
import rsconv_net_cases as cases  # noqa: E402


class Data(dense_loader.Data):
    def to(self, device):
        return self


def main():
    ref = dense_loader.load_reference(dense_chain_cases.Kernels())
    dense_loader._load("torch_points3d.utils.colors", os.path.join(REF, "utils/colors.py"))
    dense_loader._module("torch_points3d.modules.RSConv")
    rs = dense_loader._load("torch_points3d.modules.RSConv.dense", os.path.join(REF, "modules/RSConv/dense.py"))
    base = sys.modules["torch_points3d.core.common_modules.base_modules"]
    dm = sys.modules["torch_points3d.core.common_modules.dense_modules"]
    classes = {"RSConvOriginalMSGDown": rs.RSConvOriginalMSGDown, "GlobalDenseBaseModule": ref.GlobalDenseBaseModule,
               "DenseFPModule": ref.DenseFPModule}
    resolver, fetch = arch.reference_resolver(), arch.reference_fetch()
    entry = arch.load_entry("/root/reference/conf/models/segmentation/rsconv.yaml", "RSConv_MSN", {"${data.use_category}": "False"})
    dataset = type("Dataset", (), {"feature_dimension": cases.FEAT, "num_classes": cases.NUM_CLASSES})()
    resolver.resolve_model(entry, dataset, "segmentation")

    path = os.path.join(REF, "models/segmentation/rsconv.py")
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RSConvLogicModel")
    wanted = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("set_input", "forward")]
    scope = {"torch": torch, "F": F, "queue": queue, "IGNORE_LABEL": -1}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), scope)
    Net = type("Net", (torch.nn.Module,), {"set_input": scope["set_input"], "forward": scope["forward"]})
    net = Net()
    net.down_modules, net.inner_modules, net.up_modules = torch.nn.ModuleList(), torch.nn.ModuleList(), torch.nn.ModuleList()
    for opt in entry.innermost:                                          # unet.py:385-398 (built before the others)
        args = dict(opt)
        net.inner_modules.append(classes[args.pop("module_name")](**args))
    for i in range(len(entry.down_conv.down_conv_nn)):
        args = fetch(None, entry.down_conv, i)
        net.down_modules.append(classes[args.pop("module_name")](index=i, **args))
    for i in range(len(entry.up_conv.up_conv_nn)):
        args = fetch(None, entry.up_conv, i)
        net.up_modules.append(classes[args.pop("module_name")](index=i, **args))
    opt = entry.mlp_cls
    net.FC_layer = base.Seq()
    for i in range(1, len(opt.nn)):
        net.FC_layer.append(dm.Conv1D(opt.nn[i - 1], opt.nn[i], bn=True, bias=False))
    if opt.dropout:
        net.FC_layer.append(torch.nn.Dropout(p=opt.dropout))
    net.FC_layer.append(dm.Conv1D(opt.nn[-1], cases.NUM_CLASSES, activation=None, bias=True, bn=False))
    net._num_classes, net._use_category, net._weight_classes = cases.NUM_CLASSES, False, None

    cases.seeded_parameters(net, 950)
    out = {"state_dict_keys": np.array(list(net.state_dict().keys()))}
    pos, x, y = cases.make_input()
    for tag, training in (("eval", False), ("train", True)):
        net.train(training)
        with torch.no_grad():
            net.set_input(Data(pos=pos.clone(), x=x.clone(), y=y.clone()), "cpu")
            logits = net.forward()
        cases.record(out, tag, logits, net.loss_seg)
    np.savez_compressed(os.path.join(HERE, "rsconv_net.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
