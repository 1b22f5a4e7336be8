"""Run tests/golden/rsconv_mp_cases.py on the REFERENCE's own MESSAGE_PASSING RSConv modules and store the outputs.

Loaded BY FILE PATH from /root/reference (nothing is copied): modules/RSConv/message_passing.py over the base classes
make_golden_randla.py loads (core/base_conv/{base_conv,message_passing}.py, core/common_modules/base_modules.py,
core/spatial_ops/*.py).  torch_geometric is absent: `radius` is the oracle-backed stand-in of mp_chain_cases, `knn`
brute force, `MessagePassing` the restatement in make_golden_randla.py (aggr="max": 0 for a target without edges).
/root/reference does not exist on the GPU box: the test only reads the committed .npz.

    python tests/golden/make_golden_rsconv_mp.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden_randla as base  # noqa: E402
import rsconv_mp_cases as cases  # noqa: E402


def main():
    base.load_reference_base(radius=cases.radius)
    ref = base._load("ref_rsconv_mp", os.path.join(base.REF, "modules/RSConv/message_passing.py"))
    ns = types.SimpleNamespace(Convolution=ref.Convolution, RSConvDown=ref.RSConvDown)
    out = cases.run_cases(ns, base.Bag)
    np.savez_compressed(os.path.join(HERE, "rsconv_mp_chain.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
