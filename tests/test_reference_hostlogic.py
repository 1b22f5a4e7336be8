"""The host logic of torch_points3d_b200.spatial_ops against the REFERENCE's own classes: tests/golden/
spatial_ops_hostlogic.json was produced by running tests/golden/hostlogic_cases.py on the reference's
core/spatial_ops/{sampling,neighbour_finder,interpolate}.py (file-path import, third-party kernels replaced by
recording stand-ins: tests/golden/make_golden_hostlogic.py).  Here the same cases run on the mirrors with `ops.*`
replaced by the same stand-ins: constructor rules, attributes, exception types and messages, the kernel each class
calls and the scalars it passes, and the interpolation arithmetic must all agree.  No device needed."""
import json
import os
import sys

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import hostlogic_cases  # noqa: E402

from torch_points3d_b200 import ops, spatial_ops  # noqa: E402


def _close(a, b):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=1e-6, atol=1e-7)


def test_spatial_ops_host_logic_equals_the_reference(monkeypatch):
    want = json.load(open(os.path.join(GOLDEN, "spatial_ops_hostlogic.json")))
    k = hostlogic_cases.Kernels()
    monkeypatch.setattr(ops, "fps", k.fps)
    monkeypatch.setattr(ops, "furthest_point_sample", k.furthest_point_sample)
    monkeypatch.setattr(ops, "radius", k.radius)
    monkeypatch.setattr(ops, "ball_query", k.ball_query)
    monkeypatch.setattr(ops, "knn_edges", k.knn)
    got = json.loads(json.dumps(hostlogic_cases.run_cases(spatial_ops, k)))
    for section in ("sampler_constructor", "sampler_counts", "sampler_calls", "sampler_kernel_log",
                    "multiscale_constructor", "finder_calls", "finder_kernel_log"):
        assert got[section] == want[section], section
    gi, wi = got["interpolate"], want["interpolate"]
    assert (gi["num_nodes"], gi["x_idx"], gi["y_idx"]) == (wi["num_nodes"], wi["x_idx"], wi["y_idx"])
    assert _close(gi["weights"], wi["weights"]) and _close(gi["normalisation"], wi["normalisation"])
    assert _close(gi["upsampled"], wi["upsampled"])
    assert gi["wrong_size"] == wi["wrong_size"]
