"""Build the KPConv segmentation network's modules from the REFERENCE's own architecture file with the reference's
own classes, and store what came out (parameter names and shapes, per-block search / sampling settings).

Read from /root/reference (nothing is copied): conf/models/segmentation/kpconv.yaml (entry `KPConvPaper`), resolved by
the reference's own utils/model_building_utils/model_definition_resolver.py (loaded by file path; omegaconf is absent,
DictConfig / ListConfig are dict / list subclasses with attribute access); the per-module argument extraction is the
reference's own `UnwrappedUnetBasedModel._fetch_arguments_from_list` (models/base_architectures/unet.py:456-476,
that one function compiled from the file's text -- the module around it needs the trainer's BaseModel); the modules
are the reference's own KPDualBlock (modules/KPConv/blocks.py) and FPModule_PD (core/base_conv/partial_dense.py), the
classifier head follows models/segmentation/kpconv.py:52-71 via the same resolved `mlp_cls` values.
The two Hydra interpolations of the entry are given the values the benchmark uses: data.first_subsampling = 0.04
(conf/data/segmentation/s3dis1x1.yaml), data.use_category = False; dataset: 4 input features, 13 classes.

    python tests/golden/make_golden_kpconv_arch.py      ->  tests/golden/kpconv_arch.json
"""
import ast
import json
import os
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/torch_points3d"
sys.path.insert(0, HERE)
import make_golden_kpconv_blocks as loader  # noqa: E402

FEAT, NUM_CLASSES, FIRST_SUBSAMPLING = 4, 13, 0.04


def _supported(value):
    """omegaconf refuses values that are not primitives or containers (UnsupportedValueType, a ValueError): this is
    what leaves strings such as `max` alone when the resolver's eval turns them into a builtin"""
    if not isinstance(value, (type(None), bool, int, float, str, dict, list)):
        raise ValueError("Value '%s' is not a supported primitive type" % type(value).__name__)
    return value


class DictConfig(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        dict.__setitem__(self, k, _supported(v))


class ListConfig(list):
    def __setitem__(self, k, v):
        list.__setitem__(self, k, _supported(v))


def wrap(obj):
    if isinstance(obj, dict):
        return DictConfig({k: wrap(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return ListConfig([wrap(v) for v in obj])
    return obj


def reference_resolver():
    """utils/model_building_utils/model_definition_resolver.py by file path, over the omegaconf stand-ins"""
    loader._module("omegaconf")
    loader._module("omegaconf.dictconfig", DictConfig=DictConfig)
    loader._module("omegaconf.listconfig", ListConfig=ListConfig)
    return loader._load("ref_resolver", os.path.join(REF, "utils/model_building_utils/model_definition_resolver.py"))


def reference_fetch():
    """`_fetch_arguments_from_list` of models/base_architectures/unet.py:456-476 (identical in UnetBasedModel and
    UnwrappedUnetBasedModel), compiled from the file's text together with SPECIAL_NAMES"""
    path = os.path.join(REF, "models/base_architectures/unet.py")
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "UnwrappedUnetBasedModel")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_fetch_arguments_from_list")
    special = next(n for n in tree.body if isinstance(n, ast.Assign) and n.targets[0].id == "SPECIAL_NAMES")
    scope = {"is_list": lambda e: isinstance(e, list)}
    exec(compile(ast.Module(body=[special, fn], type_ignores=[]), path, "exec"), scope)
    return scope["_fetch_arguments_from_list"]


def load_entry(yaml_path, name, substitutions):
    text = open(yaml_path).read()
    for old, new in substitutions.items():
        text = text.replace(old, new)
    tree = yaml.safe_load(text)
    return wrap(tree["models"][name] if "models" in tree else tree[name])


def describe_block(block):
    kp = block.kp_conv if type(block).__name__ == "SimpleBlock" else block.kp_conv.kp_conv
    simple = block if type(block).__name__ == "SimpleBlock" else block.kp_conv
    sampler = simple.sampler
    return {"class": type(block).__name__,
            "sampler_grid": None if sampler is None else float(sampler._grid_size),
            "radius": float(simple.neighbour_finder._radius),
            "max_num_neighbors": int(simple.neighbour_finder._max_num_neighbors),
            "kernel_radius": float(kp.kernel_radius), "point_influence": float(kp.point_influence),
            "n_kernel_points": int(kp.K_points.shape[0]),
            "bn_momentum": float(simple.bn.batch_norm.momentum),
            "has_bottleneck": bool(getattr(block, "has_bottleneck", False))}


def main():
    blocks = loader.load_reference()
    loader._load("torch_points3d.core.base_conv.base_conv", os.path.join(REF, "core/base_conv/base_conv.py"))
    pdm = loader._load("torch_points3d.core.base_conv.partial_dense", os.path.join(REF, "core/base_conv/partial_dense.py"))
    resolver, fetch = reference_resolver(), reference_fetch()

    entry = load_entry("/root/reference/conf/models/segmentation/kpconv.yaml", "KPConvPaper",
                       {"${data.first_subsampling}": repr(FIRST_SUBSAMPLING), "${data.use_category}": "False"})
    dataset = type("Dataset", (), {"feature_dimension": FEAT, "num_classes": NUM_CLASSES})()
    resolver.resolve_model(entry, dataset, "segmentation")

    import torch
    np.random.seed(0)
    down, up = torch.nn.ModuleList(), torch.nn.ModuleList()
    out = {"resolved": json.loads(json.dumps(entry)), "down_blocks": [], "up": []}
    for i in range(len(entry.down_conv.down_conv_nn)):
        args = fetch(None, entry.down_conv, i)
        assert args.pop("module_name") == "KPDualBlock"
        args["index"] = i
        module = blocks.KPDualBlock(**args)
        down.append(module)
        out["down_blocks"].append([describe_block(b) for b in module.blocks])
    for i in range(len(entry.up_conv.up_conv_nn)):
        args = fetch(None, entry.up_conv, i)
        assert args.pop("module_name") == "FPModule_PD"
        args["index"] = i
        module = pdm.FPModule_PD(**args)
        up.append(module)
        out["up"].append({"k": int(module.upsample_op.k), "bn_momentum": float(module.nn[0][1].batch_norm.momentum)})
    net = torch.nn.Module()
    net.down_modules, net.up_modules = down, up
    out["state_dict"] = {k: list(v.shape) for k, v in net.state_dict().items()}
    out["num_parameters"] = int(sum(p.numel() for p in net.parameters()))
    with open(os.path.join(HERE, "kpconv_arch.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(out["state_dict"]), "tensors,", out["num_parameters"], "parameters")
    for lv in out["down_blocks"]:
        print(lv)
    print(out["up"], out["resolved"]["mlp_cls"])


if __name__ == "__main__":
    main()
