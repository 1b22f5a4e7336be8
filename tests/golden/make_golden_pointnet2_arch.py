"""Build the modules of `pointnet2_charlesmsg` (BASELINE config 2) from the REFERENCE's own architecture file with the
reference's own classes, and store what came out (per-module parameter names and shapes, sampling / search settings).

Read from /root/reference (nothing is copied): conf/models/segmentation/pointnet2.yaml (entry `pointnet2_charlesmsg`)
resolved by the reference's own resolver; per-module arguments through the reference's own
`_fetch_arguments_from_list`; modules = the reference's own PointNetMSGDown / GlobalDenseBaseModule / DenseFPModule
(modules/pointnet2/dense.py over core/base_conv/dense.py); the head's Conv1D is the reference's
(core/common_modules/dense_modules.py), assembled as models/segmentation/pointnet2.py:52-63 does from the resolved
`mlp_cls`.  The reference nests these modules recursively (UnetBasedModel), so names are stored per module.
Dataset: ShapeNet part segmentation with normals -- 3 input features, 50 classes, data.use_category = False.

    python tests/golden/make_golden_pointnet2_arch.py      ->  tests/golden/pointnet2_arch.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import dense_chain_cases  # noqa: E402
import make_golden_dense_chain as dense_loader  # noqa: E402
import make_golden_kpconv_arch as arch  # noqa: E402

FEAT, NUM_CLASSES = 3, 50


def shapes(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def main():
    ref = dense_loader.load_reference(dense_chain_cases.Kernels())
    dm = sys.modules["torch_points3d.core.common_modules.dense_modules"]
    resolver, fetch = arch.reference_resolver(), arch.reference_fetch()
    entry = arch.load_entry("/root/reference/conf/models/segmentation/pointnet2.yaml", "pointnet2_charlesmsg",
                            {"${data.use_category}": "False"})
    dataset = type("Dataset", (), {"feature_dimension": FEAT, "num_classes": NUM_CLASSES})()
    resolver.resolve_model(entry, dataset, "segmentation")
    out = {"resolved": json.loads(json.dumps(entry)), "down": [], "up": []}
    for i in range(len(entry.down_conv.down_conv_nn)):
        args = fetch(None, entry.down_conv, i)
        assert args.pop("module_name") == "PointNetMSGDown"
        args["index"] = i
        m = ref.PointNetMSGDown(**args)
        out["down"].append({"state_dict": shapes(m), "npoint": int(m.sampler._num_to_sample),
                            "radii": [float(r) for r in m.neighbour_finder._radius],
                            "nsample": [int(k) for k in m.neighbour_finder._max_num_neighbors],
                            "use_xyz": bool(m.use_xyz)})
    inner = dict(entry.innermost)
    assert inner.pop("module_name") == "GlobalDenseBaseModule"
    out["inner"] = {"state_dict": shapes(ref.GlobalDenseBaseModule(**inner))}
    for i in range(len(entry.up_conv.up_conv_nn)):
        args = fetch(None, entry.up_conv, i)
        assert args.pop("module_name") == "DenseFPModule"
        args["index"] = i
        out["up"].append({"state_dict": shapes(ref.DenseFPModule(**args))})
    head, opt = torch.nn.Sequential(), entry.mlp_cls
    for i in range(1, len(opt.nn)):
        head.append(dm.Conv1D(opt.nn[i - 1], opt.nn[i], bn=True, bias=False))
    if opt.dropout:
        head.append(torch.nn.Dropout(p=opt.dropout))
    head.append(dm.Conv1D(opt.nn[-1], NUM_CLASSES, activation=None, bias=True, bn=False))
    out["head"] = {"state_dict": shapes(head), "children": [type(c).__name__ for c in head],
                   "dropout": float(opt.dropout)}
    total = sum(sum(int(torch.tensor(s).prod()) for k, s in part["state_dict"].items()
                    if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
                for part in out["down"] + [out["inner"]] + out["up"] + [out["head"]])
    out["num_parameters"] = total
    with open(os.path.join(HERE, "pointnet2_arch.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(total, "parameters")
    for part in out["down"]:
        print({k: v for k, v in part.items() if k != "state_dict"}, len(part["state_dict"]))
    print(out["head"])


if __name__ == "__main__":
    main()
