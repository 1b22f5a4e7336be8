"""Build recipe for the CPU oracle (test infrastructure only).

Compiles oracle/p3d_oracle.c -> oracle/_build/libp3d_oracle.so with gcc.
-ffp-contract=off is REQUIRED: the fp32 squared distance must not be fused
into FMAs, otherwise ball-query membership can flip
at 1 ulp against the CUDA kernels (which use __fmul_rn/__fadd_rn).

There is no `oracle/_ref`: the reference (leo-stan/torch-points3d) is pure
Python and ships no C/C++ source to compile; its
arithmetic for this path lives in third-party wheels that are not vendored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "p3d_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libp3d_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force) and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = [
        "gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-fopenmp",
        "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
        SRC, "-o", OUT + ".tmp", "-lm",
    ]
    subprocess.run(cmd, check=True)
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
