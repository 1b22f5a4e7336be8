"""pointnet2.PointNet2_D (BASELINE config 2's network, its architecture written out in code) against the REFERENCE's
own architecture file built with the reference's own classes.

tests/golden/pointnet2_arch.json was made by tests/golden/make_golden_pointnet2_arch.py:
conf/models/segmentation/pointnet2.yaml (`pointnet2_charlesmsg`) resolved by the reference's resolver, one
PointNetMSGDown / GlobalDenseBaseModule / DenseFPModule per entry, the head of models/segmentation/pointnet2.py:52-63;
per module every parameter / buffer name with its shape, and what each level samples and searches with.
No device needed."""
import json
import os

from torch_points3d_b200 import pointnet2

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pointnet2_arch.json")


def _shapes(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def test_pointnet2_d_is_the_reference_architecture_file():
    want = json.load(open(GOLDEN))
    model = pointnet2.PointNet2_D(feat_dim=3, num_classes=50)
    assert len(model.down) == len(want["down"]) and len(model.up) == len(want["up"])
    for m, ref in zip(model.down, want["down"]):
        assert _shapes(m) == ref["state_dict"]
        assert int(m.sampler._num_to_sample) == ref["npoint"] and bool(m.use_xyz) == ref["use_xyz"]
        assert [float(r) for r in m.neighbour_finder._radius] == ref["radii"]
        assert [int(k) for k in m.neighbour_finder._max_num_neighbors] == ref["nsample"]
    assert _shapes(model.inner) == want["inner"]["state_dict"]
    for m, ref in zip(model.up, want["up"]):
        assert _shapes(m) == ref["state_dict"]
    assert _shapes(model.FC_layer) == want["head"]["state_dict"]
    assert want["head"]["children"] == ["Conv1D", "Dropout", "Conv1D"]  # the mirror's Conv1D groups are Sequentials
    assert len(model.FC_layer) == 3 and type(model.FC_layer[1]).__name__ == "Dropout"
    assert model.FC_layer[1].p == want["head"]["dropout"]
    assert sum(p.numel() for p in model.parameters()) == want["num_parameters"]
