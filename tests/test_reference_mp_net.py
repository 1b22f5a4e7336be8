"""The MESSAGE_PASSING PointNet++ network (`pointnet2` of the reference's architecture file: two SAModule, the global
module, three FPModule) composed from the MIRRORED modules, against the data path of the reference's own nested U-Net.

tests/golden/mp_net.npz was made by tests/golden/make_golden_mp_net.py (reference resolver, reference nesting code of
models/base_architectures/unet.py, reference modules; third-party functions stood in for).  Here the mirrors are
built from the same resolved arguments (tests/golden/arch_modules.json), composed the way UnetSkipConnectionBlock.forward
does -- each up module receives (output of the level below, input of its own level) -- and run on the same one-cloud
input with `ops.*` replaced by the same stand-ins.  No device needed."""
import copy
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import mp_net_cases as cases  # noqa: E402

from torch_points3d_b200 import message_passing as mp, ops, partial_dense as pd  # noqa: E402


def test_message_passing_pointnet2_equals_the_reference_network(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "mp_net.npz"))
    entry = json.load(open(os.path.join(GOLDEN, "arch_modules.json")))["pointnet2"]

    def edge_message(x, pos_src, pos_dst, src, dst):
        rel = pos_src[src] - pos_dst[dst]
        return rel if x is None else torch.cat([x[src], rel], dim=1)

    monkeypatch.setattr(ops, "fps", cases.fps)
    monkeypatch.setattr(ops, "radius", cases.radius)
    monkeypatch.setattr(ops, "knn_interpolate", cases.knn_interpolate)
    monkeypatch.setattr(ops, "edge_message", edge_message)
    monkeypatch.setattr(ops, "segment_max", cases.segment_max)
    monkeypatch.setattr(ops, "segment_sum", lambda msg, dst, T: torch.zeros((T, msg.shape[1])).index_add_(0, dst, msg))
    classes = {"SAModule": mp.SAModule, "GlobalBaseModule": mp.GlobalBaseModule, "FPModule": mp.FPModule}
    built = {"down": [], "innermost": [], "up": []}
    for record in entry["modules"]:
        built[record["flow"]].append(classes[record["module_name"]](**copy.deepcopy(record["args"])))
    net = torch.nn.Module()
    # registration order of the reference's nesting: a level's down module, everything below it, then its up module
    net.down0, net.down1, net.inner, net.up0, net.up1, net.up2 = (built["down"][0], built["down"][1], built["innermost"][0],
                                                                built["up"][0], built["up"][1], built["up"][2])
    assert ["%s" % (tuple(v.shape),) for v in net.state_dict().values()] == [str(s) for s in want["state_dict_shapes"]]
    cases.seeded_parameters(net, 900)
    net.eval()
    pos, x, batch = cases.make_input()
    with torch.no_grad():
        d0 = pd._Data(pos=pos, x=x, batch=batch)
        d1 = net.down0(d0)
        d2 = net.down1(d1)
        u = net.up0((net.inner(d2), d2))
        u = net.up1((u, d1))
        u = net.up2((u, d0))
    assert torch.equal(u.pos, torch.from_numpy(want["pos"])) and torch.equal(u.batch, torch.from_numpy(want["batch"]))
    ref = want["x"]
    assert u.x.shape == ref.shape
    assert float(np.abs(u.x.numpy() - ref).max()) <= 1e-5 * float(np.abs(ref).max())
