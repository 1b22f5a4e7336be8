"""Kernel-point disposition of the KPConv layers, read and processed by the REFERENCE's own code.

Loaded BY FILE PATH from /root/reference (nothing is copied): modules/KPConv/{plyutils,kernel_utils}.py.  Stored:
the 15-point `center` disposition as the reference's own `read_ply` returns it from
modules/KPConv/kernels/dispositions/k_015_center.ply (this package ships the same table as k_015_center.npy), and what
the reference's `load_kernels(radius, 15, 1, 3, "center")` makes of it (scaled, randomly rotated, 1 % jitter) for one
numpy seed -- compared through rotation-invariant quantities, the draw itself is not reproducible across generators.

    python tests/golden/make_golden_kernel_points.py      ->  tests/golden/kernel_points.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_kpconv_blocks as loader  # noqa: E402

RADIUS = 0.06


def main():
    loader._module("matplotlib")
    loader._module("matplotlib.pyplot")
    for pkg in ("torch_points3d", "torch_points3d.modules", "torch_points3d.modules.KPConv"):
        loader._module(pkg)
    kp = "torch_points3d.modules.KPConv."
    ply = loader._load(kp + "plyutils", os.path.join(loader.REF, "modules/KPConv/plyutils.py"))
    ku = loader._load(kp + "kernel_utils", os.path.join(loader.REF, "modules/KPConv/kernel_utils.py"))
    data = ply.read_ply(os.path.join(loader.REF, "modules/KPConv/kernels/dispositions/k_015_center.ply"))
    table = np.vstack((data["x"], data["y"], data["z"])).T
    np.random.seed(3)
    made = np.asarray(ku.load_kernels(RADIUS, 15, 1, 3, "center"))
    out = {"disposition": table, "radius": np.asarray(RADIUS), "loaded": made.reshape(15, 3),
           "loaded_shape": np.asarray(made.shape), "loaded_dtype": np.asarray(str(made.dtype))}
    np.savez_compressed(os.path.join(HERE, "kernel_points.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if v.ndim else v)


if __name__ == "__main__":
    main()
