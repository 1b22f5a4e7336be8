"""Every down / innermost / up module of the reference's architecture files for the four model families on the path
(17 entries of conf/models/segmentation/{pointnet2,rsconv,randlanet,kpconv}.yaml), built from the reference's resolved
constructor arguments with the MIRRORED classes, must hold the tensors the reference's own classes hold.

tests/golden/arch_modules.json was made by tests/golden/make_golden_arch_modules.py (reference resolver, reference
argument extraction, reference classes).  Names, shapes and state_dict order are compared module by module: this is
the "does the reference's configuration drive the mirror" check; the arithmetic of the modules is pinned by the
test_reference_*chain* / *blocks* tests.  No device needed."""
import copy
import json
import os

import numpy as np
import pytest

from torch_points3d_b200 import kpconv, message_passing as mp, partial_dense as pd, pointnet2, rsconv

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "arch_modules.json")))
MIRRORS = {
    "DENSE": {"PointNetMSGDown": pointnet2.PointNetMSGDown, "DenseFPModule": pointnet2.DenseFPModule,
              "GlobalDenseBaseModule": pointnet2.GlobalDenseBaseModule,
              "RSConvOriginalMSGDown": rsconv.RSConvOriginalMSGDown},
    "MESSAGE_PASSING": {"SAModule": mp.SAModule, "FPModule": mp.FPModule, "GlobalBaseModule": mp.GlobalBaseModule,
                        "RandLANetRes": mp.RandLANetRes, "RandlaConv": mp.RandlaConv, "RSConvDown": rsconv.RSConvDown},
    "PARTIAL_DENSE": {"KPDualBlock": kpconv.KPDualBlock, "FPModule_PD": pd.FPModule_PD,
                      "GlobalBaseModule": mp.GlobalBaseModule},
}


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_mirrors_take_the_reference_configuration(name):
    entry = GOLDEN[name]
    classes = MIRRORS[entry["conv_type"]]
    checked = 0
    for record in entry["modules"]:
        where = "%s %s[%d] %s" % (name, record["flow"], record["position"], record["module_name"])
        assert record["module_name"] in classes, where
        np.random.seed(0)
        module = classes[record["module_name"]](**copy.deepcopy(record["args"]))
        if "state_dict" not in record:
            continue                                        # built by the mirror; the generator had no reference class
        got = {k: list(v.shape) for k, v in module.state_dict().items()}
        assert list(got) == record["state_dict_order"], where
        assert got == record["state_dict"], where
        checked += 1
    assert checked >= len(entry["modules"]) - 1
