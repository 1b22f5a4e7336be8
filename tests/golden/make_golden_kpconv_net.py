"""Run the REFERENCE's own KPConv segmentation network forward + loss on a small input and store the outputs.

Executed from /root/reference by file path (nothing is copied):
  * the modules: KPDualBlock / FPModule_PD built from conf/models/segmentation/kpconv.yaml (`KPConvPaper`) exactly as
    make_golden_kpconv_arch.py does (reference resolver, reference argument extraction, reference classes; the
    reference's torch KPConv_ops runs on CPU; ball query and voxel means are the oracle-backed stand-ins of
    kpconv_blocks_cases, `knn_interpolate` the brute-force stand-in of mp_chain_cases);
  * the network code: the methods `set_input`, `forward` and `compute_loss` of models/segmentation/kpconv.py:82-153,
    compiled from the file's text and bound to a bare nn.Module that holds `down_modules`, `inner_modules`
    ([Identity]), `up_modules`, `FC_layer` (the class around them needs the trainer's BaseModel / model factory);
    FC_layer is assembled as kpconv.py:56-75 does for the resolved `mlp_cls` from the reference's own FastBatchNorm1d.
data.first_subsampling = 0.25 for the small input of kpconv_net_cases.py.

    python tests/golden/make_golden_kpconv_net.py      ->  tests/golden/kpconv_net.npz
"""
import ast
import os
import sys
from typing import Any

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import kpconv_net_cases as cases  # noqa: E402
import make_golden_kpconv_arch as arch  # noqa: E402
import make_golden_kpconv_blocks as loader  # noqa: E402
import make_golden_multiscale as pyg  # noqa: E402  (torch_geometric Data / Batch restated)
import mp_chain_cases  # noqa: E402
import multiscale_cases  # noqa: E402


class Data(pyg.Data):
    def to(self, device):
        return self


class Batch(pyg.Batch, Data):
    pass


def main():
    blocks = loader.load_reference()
    sys.modules["torch_geometric.data"].Data = Data
    sys.modules["torch_geometric.data"].Batch = Batch
    sys.modules["torch_geometric"].is_debug_enabled = lambda: False
    loader._load("torch_points3d.core.base_conv.base_conv", os.path.join(REF, "core/base_conv/base_conv.py"))
    pdm = loader._load("torch_points3d.core.base_conv.partial_dense", os.path.join(REF, "core/base_conv/partial_dense.py"))
    interp = sys.modules["torch_points3d.core.spatial_ops.interpolate"]
    interp.knn_interpolate, interp.knn, interp.scatter_add = mp_chain_cases.knn_interpolate, multiscale_cases.knn, pyg.scatter_add
    interp.Data = Data
    msd = loader._load("torch_points3d.datasets.multiscale_data", os.path.join(REF, "datasets/multiscale_data.py"))
    base = sys.modules["torch_points3d.core.common_modules.base_modules"]
    kernels = sys.modules["torch_points3d.modules.KPConv.kernels"]
    resolver, fetch = arch.reference_resolver(), arch.reference_fetch()
    entry = arch.load_entry("/root/reference/conf/models/segmentation/kpconv.yaml", "KPConvPaper",
                            {"${data.first_subsampling}": repr(cases.FIRST_SUBSAMPLING), "${data.use_category}": "False"})
    dataset = type("Dataset", (), {"feature_dimension": cases.FEAT, "num_classes": cases.NUM_CLASSES})()
    resolver.resolve_model(entry, dataset, "segmentation")

    np.random.seed(0)
    Net = type("Net", (torch.nn.Module,), {})
    net = Net()
    net.down_modules, net.inner_modules, net.up_modules = torch.nn.ModuleList(), torch.nn.ModuleList(), torch.nn.ModuleList()
    net.inner_modules.append(base.Identity())
    for i in range(len(entry.down_conv.down_conv_nn)):
        args = fetch(None, entry.down_conv, i)
        args.pop("module_name")
        net.down_modules.append(blocks.KPDualBlock(index=i, **args))
    for i in range(len(entry.up_conv.up_conv_nn)):
        args = fetch(None, entry.up_conv, i)
        args.pop("module_name")
        net.up_modules.append(pdm.FPModule_PD(index=i, **args))
    opt = entry.mlp_cls
    assert not opt.dropout
    net.FC_layer = torch.nn.Sequential()
    net.FC_layer.add_module("1", torch.nn.Sequential(torch.nn.Linear(opt.nn[0], opt.nn[1], bias=False),
                                                    base.FastBatchNorm1d(opt.nn[1], momentum=opt.bn_momentum),
                                                    torch.nn.LeakyReLU(0.2)))
    net.FC_layer.add_module("Class", torch.nn.Linear(opt.nn[1], cases.NUM_CLASSES, bias=False))
    net.FC_layer.add_module("Softmax", torch.nn.LogSoftmax(-1))
    net._use_category, net._weight_classes, net.lambda_reg, net.lambda_internal_losses = False, None, 0, None

    path = os.path.join(REF, "models/segmentation/kpconv.py")
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "KPConvPaper")
    wanted = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("set_input", "forward", "compute_loss")]
    scope = {"torch": torch, "F": F, "Any": Any, "add_ones": kernels.add_ones, "Identity": base.Identity,
             "MultiScaleBatch": msd.MultiScaleBatch, "IGNORE_LABEL": -1}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), scope)
    for name in ("set_input", "forward", "compute_loss"):
        setattr(Net, name, scope[name])

    cases.seeded_parameters(net, 700)
    out = {"state_dict_keys": np.array(list(net.state_dict().keys()))}
    pos, x, batch, y = cases.make_input()
    for tag, training in (("eval", False), ("train", True)):
        net.train(training)
        with torch.no_grad():
            net.set_input(Data(pos=pos, x=x, batch=batch, y=y), "cpu")
            log_probs = net.forward()
        cases.record(out, tag, log_probs, net.loss)
    out["train/head_running_mean"] = net.FC_layer[0][1].batch_norm.running_mean.numpy().copy()

    # ---- the same network fed with a pyramid made the reference's way: MultiScaleTransform per sample
    # (core/data_transform/transforms.py:579-651, compiled from the file's text) with the samplers / neighbour finders
    # / upsample operators the network's own modules hold (collected like unet.py:315-331), collated by
    # MultiScaleBatch.from_data_list; set_input hands the lists to the blocks through `precomputed=`
    import functools
    tpath = os.path.join(REF, "core/data_transform/transforms.py")
    ttree = ast.parse(open(tpath).read(), filename=tpath)
    tnode = next(n for n in ttree.body if isinstance(n, ast.ClassDef) and n.name == "MultiScaleTransform")
    tscope = {"partial": functools.partial, "Data": Data, "MultiScaleData": msd.MultiScaleData, "torch": torch}
    exec(compile(ast.Module(body=[tnode], type_ignores=[]), tpath, "exec"), tscope)
    strategies = {"sampler": [], "neighbour_finder": [], "upsample_op": []}
    for m in net.down_modules:
        strategies["sampler"] += m.sampler
        strategies["neighbour_finder"] += m.neighbour_finder
    for m in net.up_modules:
        strategies["upsample_op"].append(m.upsample_op)
    transform = tscope["MultiScaleTransform"](strategies)
    for tag, training in (("eval_precomputed", False), ("train_precomputed", True)):
        samples = []
        for b in range(int(batch.max()) + 1):
            sel = batch == b
            samples.append(transform(Data(pos=pos[sel].clone(), x=x[sel].clone(), y=y[sel].clone())))
        mb = msd.MultiScaleBatch.from_data_list(samples)
        mb.to = lambda device, mb=mb: mb
        out["pyramid/levels"] = np.array([int(l.pos.shape[0]) for l in mb.multiscale])
        out["pyramid/last_table"] = mb.multiscale[-1].idx_neighboors.numpy().copy()
        net.train(training)
        with torch.no_grad():
            net.set_input(mb, "cpu")
            log_probs = net.forward()
        cases.record(out, tag, log_probs, net.loss)
    np.savez_compressed(os.path.join(HERE, "kpconv_net.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
