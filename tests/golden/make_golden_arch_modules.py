"""Build every down / innermost / up module of the reference's own architecture files for the model families on the
path, with the reference's own classes, and store per module its resolved constructor arguments and the names and
shapes of its parameters and buffers.  tests/test_reference_arch_modules.py then hands the SAME arguments to the
mirrored classes: a mirror that does not take the reference's configuration, or builds different tensors from it,
fails there.

Read from /root/reference (nothing is copied): conf/models/segmentation/{pointnet2,rsconv,randlanet,kpconv}.yaml, the
resolver and `_fetch_arguments_from_list` (see make_golden_kpconv_arch.py), and per family the module files the other
generators load by file path.  The three families need different stand-ins for the absent third-party packages, so
each runs in its own process (`--family`).  Hydra interpolations: data.use_category = False,
data.first_subsampling = 0.04; FEAT = 3 (4 for KPConv).

    python tests/golden/make_golden_arch_modules.py      ->  tests/golden/arch_modules.json
"""
import copy
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
CONF = "/root/reference/conf/models/segmentation"
SUBS = {"${data.use_category}": "False", "${data.first_subsampling}": "0.04"}
FAMILIES = {
    "dense": [("pointnet2.yaml", n) for n in ("pointnet2_largemsg", "pointnet2_charlesmsg", "pointnet2_charlesssg",
                                              "pointnet2_indoor")] +
             [("rsconv.yaml", n) for n in ("RSConv_MSN", "RSConv_MSN_S3DIS", "RSConv_Indoor")],
    "message_passing": [("pointnet2.yaml", "pointnet2"), ("pointnet2.yaml", "pointnet2ms"),
                        ("randlanet.yaml", "Randlanet_Res"), ("randlanet.yaml", "Randlanet_Conv"),
                        ("rsconv.yaml", "RSConv_2LD"), ("rsconv.yaml", "RSConv_4LD")],
    "kpconv": [("kpconv.yaml", n) for n in ("KPConvPaper", "KPDeformableConvPaper", "KPConvPaper_innermost",
                                            "KPConvPaper2LDMini")],
}


def plain(obj):
    return json.loads(json.dumps(obj))


def reference_classes(family):
    if family == "dense":
        import dense_chain_cases
        import make_golden_dense_chain as loader
        ref = loader.load_reference(dense_chain_cases.Kernels())
        loader._load("torch_points3d.utils.colors", os.path.join(loader.REF, "utils/colors.py"))
        loader._module("torch_points3d.modules.RSConv")
        rs = loader._load("torch_points3d.modules.RSConv.dense", os.path.join(loader.REF, "modules/RSConv/dense.py"))
        return {"PointNetMSGDown": ref.PointNetMSGDown, "DenseFPModule": ref.DenseFPModule,
                "GlobalDenseBaseModule": ref.GlobalDenseBaseModule, "RSConvOriginalMSGDown": rs.RSConvOriginalMSGDown}
    if family == "message_passing":
        import make_golden_mp_chain as mp_loader
        import make_golden_randla as loader
        import mp_chain_cases
        loader.load_reference_base(radius=mp_chain_cases.radius)
        sys.modules["torch_geometric.nn"].PointConv = mp_loader.PointConv
        base = sys.modules["torch_points3d.core.base_conv.message_passing"]
        sa = loader._load("ref_pointnet2_mp", os.path.join(loader.REF, "modules/pointnet2/message_passing.py"))
        rl = loader._load("ref_randla", os.path.join(loader.REF, "modules/RandLANet/modules.py"))
        rs = loader._load("ref_rsconv_mp", os.path.join(loader.REF, "modules/RSConv/message_passing.py"))
        return {"SAModule": sa.SAModule, "FPModule": base.FPModule, "GlobalBaseModule": base.GlobalBaseModule,
                "RandLANetRes": rl.RandLANetRes, "RandlaConv": rl.RandlaConv, "RSConvDown": rs.RSConvDown}
    import make_golden_kpconv_blocks as loader
    blocks = loader.load_reference()
    loader._load("torch_points3d.core.base_conv.base_conv", os.path.join(loader.REF, "core/base_conv/base_conv.py"))
    pdm = loader._load("torch_points3d.core.base_conv.partial_dense", os.path.join(loader.REF, "core/base_conv/partial_dense.py"))
    return {"KPDualBlock": blocks.KPDualBlock, "FPModule_PD": pdm.FPModule_PD}


def run_family(family):
    import numpy as np
    import make_golden_kpconv_arch as arch
    classes = reference_classes(family)
    resolver, fetch = arch.reference_resolver(), arch.reference_fetch()
    dataset = type("Dataset", (), {"feature_dimension": 4 if family == "kpconv" else 3, "num_classes": 13})()
    out = {}
    for yaml_name, name in FAMILIES[family]:
        entry = arch.load_entry(os.path.join(CONF, yaml_name), name, SUBS)
        resolver.resolve_model(entry, dataset, "segmentation")
        todo = []
        for i in range(len(entry.down_conv.down_conv_nn)):
            todo.append(("down", i, fetch(None, entry.down_conv, i)))
        inner = entry.get("innermost")
        if inner is not None:
            for i, opt in enumerate(inner if isinstance(inner, list) else [inner]):
                todo.append(("innermost", i, dict(opt)))
        if entry.get("up_conv") is not None:
            for i in range(len(entry.up_conv.up_conv_nn)):
                todo.append(("up", i, fetch(None, entry.up_conv, i)))
        modules = []
        for flow, i, args in todo:
            args = plain(args)
            module_name = args.pop("module_name")
            if flow != "innermost":
                args["index"] = i
            record = {"flow": flow, "position": i, "module_name": module_name, "args": copy.deepcopy(args)}
            if module_name not in classes:
                record["not_built"] = "class belongs to another family's loader"
            else:
                np.random.seed(0)
                module = classes[module_name](**args)
                record["state_dict"] = {k: list(v.shape) for k, v in module.state_dict().items()}
                record["state_dict_order"] = list(module.state_dict().keys())
            modules.append(record)
        out[name] = {"conv_type": entry.conv_type, "yaml": yaml_name, "modules": modules,
                     "mlp_cls": plain(entry.get("mlp_cls"))}
    return out


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--family":
        print("JSON:" + json.dumps(run_family(sys.argv[2])))
        return
    merged = {}
    for family in FAMILIES:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--family", family], capture_output=True, text=True)
        if res.returncode != 0:
            sys.exit(res.stderr[-3000:])
        line = next(l for l in res.stdout.splitlines() if l.startswith("JSON:"))
        merged.update(json.loads(line[5:]))
    with open(os.path.join(HERE, "arch_modules.json"), "w") as f:
        json.dump(merged, f, indent=1, sort_keys=True)
    for name, e in merged.items():
        built = [m for m in e["modules"] if "state_dict" in m]
        print("%-24s %-16s %2d modules, %3d tensors%s" % (name, e["conv_type"], len(e["modules"]),
              sum(len(m["state_dict"]) for m in built),
              "" if len(built) == len(e["modules"]) else "  (%d not built)" % (len(e["modules"]) - len(built))))


if __name__ == "__main__":
    main()
