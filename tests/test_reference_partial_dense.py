"""The PARTIAL_DENSE host chain of torch_points3d_b200.partial_dense against the REFERENCE's own classes.

tests/golden/partial_dense_chain.npz holds outputs of the reference's BasePartialDenseConvolutionDown.forward,
GlobalPartialDenseBaseModule and FPModule_PD (core/base_conv/partial_dense.py:35-146, file-path import with the absent
third-party functions replaced by the torch stand-ins of tests/golden/partial_dense_cases.py:
tests/golden/make_golden_partial_dense.py).  The same cases run here on the mirrors with `ops.*` replaced by the same
stand-ins: shadow padding and centring of the gathers, what `conv` receives, which attributes are carried over, the
pooled outputs, interpolation with / without precomputed tables, the `up_idx` counter, the innermost broadcast and the
parameter names must agree.  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import partial_dense_cases as cases  # noqa: E402

from torch_points3d_b200 import ops, partial_dense as pd  # noqa: E402


def test_partial_dense_chain_equals_the_reference_classes(monkeypatch):
    want = np.load(os.path.join(GOLDEN, "partial_dense_chain.npz"))

    def segment_max(x, batch, n):
        return cases.global_max_pool(x, batch, n)

    def segment_sum(x, batch, n):
        return torch.zeros((n, x.shape[1]), dtype=x.dtype).index_add_(0, batch, x)

    monkeypatch.setattr(ops, "knn_edges", cases.knn)
    monkeypatch.setattr(ops, "knn_interpolate", cases.knn_interpolate)
    monkeypatch.setattr(ops, "segment_max", segment_max)
    monkeypatch.setattr(ops, "segment_sum", segment_sum)
    ns = SimpleNamespace(BasePartialDenseConvolutionDown=pd.BasePartialDenseConvolutionDown,
                         GlobalPartialDenseBaseModule=pd.GlobalPartialDenseBaseModule, FPModule_PD=pd.FPModule_PD)
    got = cases.run_cases(ns, pd._Data)
    assert sorted(got.keys()) == sorted(want.files)
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "US" or value.dtype == bool or value.dtype.kind == "i":
            assert np.array_equal(value, ref), key
        else:
            assert value.shape == ref.shape, key
            assert np.allclose(value, ref, rtol=1e-5, atol=1e-6), (key, float(np.abs(value - ref).max()))
