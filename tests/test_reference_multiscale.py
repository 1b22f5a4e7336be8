"""The pyramid wire format of torch_points3d_b200.multiscale against the REFERENCE's own code.

tests/golden/multiscale.npz holds outputs of the reference's MultiScaleTransform (core/data_transform/transforms.py:
579-651), MultiScaleBatch.from_data_list (datasets/multiscale_data.py:56-104) and from_data_list_token (:107-165), run by
tests/golden/make_golden_multiscale.py with torch_geometric's Data / Batch stood in for (see that file).  The same
cases run here on the mirrors: per-sample tables, the upsample list innermost first, the collation's index shifting
(shadow indices shifted with the rest, negative entries left alone, x_idx / y_idx by query / support counts), bool /
`*index*` / python-number attributes, error types and messages.  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import multiscale_cases as cases  # noqa: E402

from torch_points3d_b200 import multiscale, ops, partial_dense as pd, spatial_ops  # noqa: E402


def test_pyramid_format_equals_the_reference(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "multiscale.npz"))
    monkeypatch.setattr(ops, "knn_edges", cases.knn)
    ns = SimpleNamespace(Data=pd._Data, MultiScaleData=multiscale.MultiScaleData,
                         MultiScaleBatch=multiscale.MultiScaleBatch, MultiScaleTransform=multiscale.MultiScaleTransform,
                         from_data_list_token=multiscale.from_data_list_token,
                         KNNInterpolate=spatial_ops.KNNInterpolate)
    got = cases.run_cases(ns)
    assert sorted(got.keys()) == sorted(want.files)
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "USbi":
            assert value.shape == ref.shape and np.array_equal(value, ref), key
        else:
            assert value.shape == ref.shape, key
            assert np.allclose(value, ref, rtol=1e-6, atol=1e-7), (key, float(np.abs(value - ref).max()))
