"""The dense RSConv down modules of torch_points3d_b200.rsconv against the REFERENCE's own modules/RSConv/dense.py.

tests/golden/rsconv_chain.npz holds outputs of the reference's RSConvOriginalMSGDown / RSConvSharedMSGDown /
RSConvMSGDown (file-path import, `torch_points_kernels` replaced by the oracle-backed stand-ins of
tests/golden/dense_chain_cases.py: tests/golden/make_golden_rsconv.py).  The same cases run here on the mirrors with
`ops.*` replaced by the same stand-ins (the fused grouping op by its documented contract): relation vectors (incl.
the in-place centring of RSConvOriginalMSGDown), shared convolutions / per-scale norms, the lift of the first layer,
max over the neighbours, channel raising -- outputs in training and eval mode, gradients w.r.t. positions and
features, running statistics, parameter names in state_dict order.  No device needed."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
sys.path.insert(0, os.path.dirname(__file__))
import dense_chain_cases  # noqa: E402
import rsconv_cases  # noqa: E402
from test_reference_dense_chain import _patch  # noqa: E402

from torch_points3d_b200 import rsconv  # noqa: E402


@pytest.mark.parametrize("fused", [False, True])
def test_rsconv_dense_equals_the_reference_modules(monkeypatch, fused):
    want = np.load(os.path.join(GOLDEN, "rsconv_chain.npz"))
    _patch(monkeypatch, dense_chain_cases.Kernels())
    ns = SimpleNamespace(RSConvOriginalMSGDown=rsconv.RSConvOriginalMSGDown,
                         RSConvSharedMSGDown=rsconv.RSConvSharedMSGDown, RSConvMSGDown=rsconv.RSConvMSGDown,
                         kwargs={"fused_grouping": fused})
    got = rsconv_cases.run_cases(ns, SimpleNamespace)
    assert sorted(got.keys()) == sorted(want.files)
    for key, value in got.items():
        ref = want[key]
        if value.dtype.kind in "US":
            assert str(value) == str(ref), key
        else:
            assert value.shape == ref.shape, key
            assert np.allclose(value, ref, rtol=1e-5, atol=1e-6), (key, float(np.abs(value - ref).max()))
