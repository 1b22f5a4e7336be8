"""Run the REFERENCE's own `pointnet2_charlesmsg` network forward + loss on a small input and store the outputs.

Executed from /root/reference by file path (nothing is copied):
  * the modules: PointNetMSGDown / GlobalDenseBaseModule / DenseFPModule (modules/pointnet2/dense.py over
    core/base_conv/dense.py), `torch_points_kernels` replaced by the oracle-backed stand-ins of dense_chain_cases;
  * the nesting: `BaseFactory`, `UnetSkipConnectionBlock` and UnetBasedModel's `_init_from_compact_format`,
    `_fetch_arguments_from_list`, `_fetch_arguments_up_and_down`, `_get_factory`, `_save_sampling_and_search`
    (models/base_architectures/unet.py), compiled from the file's text and bound to a bare nn.Module (the class
    around them needs the trainer's BaseModel), fed with conf/models/segmentation/pointnet2.yaml
    (`pointnet2_charlesmsg`) resolved by the reference's own resolver;
  * the network code: `set_input` and `forward` of models/segmentation/pointnet2.py:66-113, compiled the same way;
    FC_layer assembled as :52-63 does from the reference's own Seq / Conv1D.

    python tests/golden/make_golden_pointnet2_net.py      ->  tests/golden/pointnet2_net.npz
"""
import ast
import logging
import os
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
import pointnet2_net_cases as cases  # noqa: E402


class Data(dense_loader.Data):
    def to(self, device):
        return self


def _functions(tree, class_name, names):
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    return [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]


BUILDERS = ("_init_from_compact_format", "_fetch_arguments_from_list", "_fetch_arguments_up_and_down", "_get_factory",
            "_save_sampling_and_search")


def reference_unet_scope(identity_cls):
    """BaseFactory, UnetSkipConnectionBlock and UnetBasedModel's builder methods, compiled from unet.py's text"""
    path = os.path.join(REF, "models/base_architectures/unet.py")
    tree = ast.parse(open(path).read(), filename=path)
    keep = [n for n in tree.body if (isinstance(n, ast.Assign) and n.targets[0].id == "SPECIAL_NAMES")
            or (isinstance(n, ast.ClassDef) and n.name in ("BaseFactory", "UnetSkipConnectionBlock"))]
    scope = {"nn": torch.nn, "is_list": lambda e: isinstance(e, list), "Identity": identity_cls,
             "log": logging.getLogger("ref")}
    exec(compile(ast.Module(body=keep + _functions(tree, "UnetBasedModel", BUILDERS), type_ignores=[]), path, "exec"), scope)
    return scope


def main():
    modules_lib = dense_loader.load_reference(dense_chain_cases.Kernels())
    base = sys.modules["torch_points3d.core.common_modules.base_modules"]
    dm = sys.modules["torch_points3d.core.common_modules.dense_modules"]
    resolver = arch.reference_resolver()
    entry = arch.load_entry("/root/reference/conf/models/segmentation/pointnet2.yaml", "pointnet2_charlesmsg",
                            {"${data.use_category}": "False"})
    dataset = type("Dataset", (), {"feature_dimension": cases.FEAT, "num_classes": cases.NUM_CLASSES})()
    resolver.resolve_model(entry, dataset, "segmentation")

    builders, scope = BUILDERS, reference_unet_scope(base.Identity)
    path2 = os.path.join(REF, "models/segmentation/pointnet2.py")
    tree2 = ast.parse(open(path2).read(), filename=path2)
    scope2 = {"torch": torch, "F": F, "Data": Data, "IGNORE_LABEL": -1}
    exec(compile(ast.Module(body=_functions(tree2, "PointNet2_D", ("set_input", "forward")), type_ignores=[]), path2, "exec"), scope2)

    Net = type("Net", (torch.nn.Module,), {})
    for name in builders:
        setattr(Net, name, scope[name])
    for name in ("set_input", "forward"):
        setattr(Net, name, scope2[name])
    net = Net()
    net._spatial_ops_dict = {"neighbour_finder": [], "sampler": [], "upsample_op": []}
    net._init_from_compact_format(entry, "pointnet2", dataset, modules_lib)
    opt = entry.mlp_cls
    net.FC_layer = base.Seq()
    for i in range(1, len(opt.nn)):
        net.FC_layer.append(dm.Conv1D(opt.nn[i - 1], opt.nn[i], bn=True, bias=False))
    if opt.dropout:
        net.FC_layer.append(torch.nn.Dropout(p=opt.dropout))
    net.FC_layer.append(dm.Conv1D(opt.nn[-1], cases.NUM_CLASSES, activation=None, bias=True, bn=False))
    net._num_classes, net._use_category, net._weight_classes = cases.NUM_CLASSES, False, None

    cases.seeded_parameters(net, 800)
    out = {"state_dict_keys": np.array(list(net.state_dict().keys())),
           "state_dict_shapes": np.array(["%s" % (tuple(v.shape),) for v in net.state_dict().values()])}
    pos, x, y = cases.make_input()
    for tag, training in (("eval", False), ("train_mode_no_dropout", True)):
        net.train(training)
        for m in net.modules():                                # dropout is a random draw: compared with p applied = off
            if isinstance(m, torch.nn.Dropout):
                m.eval()
        with torch.no_grad():
            net.set_input(Data(pos=pos, x=x, y=y), "cpu")
            logits = net.forward()
        cases.record(out, tag, logits, net.loss_seg)
    np.savez_compressed(os.path.join(HERE, "pointnet2_net.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
