"""Run tests/golden/rsconv_cases.py on the REFERENCE's own dense RSConv modules and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): modules/RSConv/dense.py over core/base_conv/dense.py,
core/common_modules/dense_modules.py and core/spatial_ops/*.py (the loader of make_golden_dense_chain.py);
`torch_points_kernels` is replaced by the oracle-backed stand-ins of dense_chain_cases, utils/colors.py is the
reference's own file.  /root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_rsconv.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import dense_chain_cases  # noqa: E402
import make_golden_dense_chain as dense_loader  # noqa: E402
import rsconv_cases  # noqa: E402


def main():
    dense_loader.load_reference(dense_chain_cases.Kernels())
    dense_loader._load("torch_points3d.utils.colors", os.path.join(dense_loader.REF, "utils/colors.py"))
    dense_loader._module("torch_points3d.modules.RSConv")
    ref = dense_loader._load("torch_points3d.modules.RSConv.dense", os.path.join(dense_loader.REF, "modules/RSConv/dense.py"))
    ns = types.SimpleNamespace(RSConvOriginalMSGDown=ref.RSConvOriginalMSGDown,
                               RSConvSharedMSGDown=ref.RSConvSharedMSGDown, RSConvMSGDown=ref.RSConvMSGDown)
    out = rsconv_cases.run_cases(ns, dense_loader.Data)
    np.savez_compressed(os.path.join(HERE, "rsconv_chain.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else str(v)[:150])


if __name__ == "__main__":
    main()
