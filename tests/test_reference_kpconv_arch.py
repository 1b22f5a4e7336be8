"""kpconv.KPConvPaper (the C5 network, its architecture written out in code) against the REFERENCE's own architecture
file built with the reference's own classes.

tests/golden/kpconv_arch.json was made by tests/golden/make_golden_kpconv_arch.py: conf/models/segmentation/kpconv.yaml
(`KPConvPaper`) resolved by the reference's resolver, one KPDualBlock / FPModule_PD per entry through the reference's
`_fetch_arguments_from_list`, then every parameter / buffer name with its shape and, per block, what it searches and
samples with.  The mirror must produce the same table.  No device needed."""
import json
import os

import pytest

from torch_points3d_b200 import kpconv

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kpconv_arch.json")


def _describe(block):
    simple = block if type(block).__name__ == "SimpleBlock" else block.kp_conv
    kp = simple.kp_conv
    nf = simple.neighbour_finder
    return {"class": type(block).__name__,
            "sampler_grid": None if simple.sampler is None else float(simple.sampler._size),
            "radius": float(nf._radius), "max_num_neighbors": int(nf._max_num_neighbors),
            "kernel_radius": float(kp.kernel_radius), "point_influence": float(kp.point_influence),
            "n_kernel_points": int(kp.K_points.shape[0]), "bn_momentum": float(simple.bn.batch_norm.momentum),
            "has_bottleneck": bool(getattr(block, "has_bottleneck", False))}


def test_kpconv_paper_is_the_reference_architecture_file():
    want = json.load(open(GOLDEN))
    model = kpconv.KPConvPaper(4, 13, in_feat=64, first_subsampling=0.04, bn_momentum=0.02)
    sd = {k: list(v.shape) for k, v in model.state_dict().items()}
    body = {k: v for k, v in sd.items() if not k.startswith("FC_layer.")}
    assert sorted(body) == sorted(want["state_dict"])                      # same names, nothing extra, nothing missing
    for key, shape in want["state_dict"].items():
        assert body[key] == shape, key
    body_params = sum(p.numel() for n, p in model.named_parameters() if not n.startswith("FC_layer."))
    assert body_params == want["num_parameters"]
    # what every block searches and samples with
    assert len(model.down_modules) == len(want["down_blocks"])
    for lv, (module, ref_blocks) in enumerate(zip(model.down_modules, want["down_blocks"])):
        got = [_describe(b) for b in module.blocks]
        assert len(got) == len(ref_blocks)
        for g, r in zip(got, ref_blocks):
            for key, value in r.items():
                if isinstance(value, float):
                    assert g[key] == pytest.approx(value, rel=1e-12), (lv, key)
                else:
                    assert g[key] == value, (lv, key)
    assert [{"k": m.upsample_op.k, "bn_momentum": m.nn[0][1].batch_norm.momentum} for m in model.up_modules] == want["up"]
    # the head of models/segmentation/kpconv.py:52-71 for the resolved mlp_cls entry
    head = want["resolved"]["mlp_cls"]
    assert head["dropout"] == 0 and [n for n, _ in model.FC_layer.named_children()] == ["1", "Class", "Softmax"]
    assert sd["FC_layer.1.0.weight"] == [head["nn"][1], head["nn"][0]] and sd["FC_layer.Class.weight"] == [13, head["nn"][1]]
    assert model.FC_layer[0][1].batch_norm.momentum == head["bn_momentum"]
    # the yaml's neighbour caps are the class constant the benchmarks read
    assert kpconv.KPConvPaper.MAX_NN == want["resolved"]["down_conv"]["max_num_neighbors"]
