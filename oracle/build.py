"""Compile oracle/papc_oracle.c -> oracle/libpapc_oracle.so (gcc, no GPU needed).

TEST INFRASTRUCTURE.  Called by __graft_entry__.build(); the .so is git-ignored but
travels to the GPU box with the gpurun snapshot.  The reference itself is pure Python
(+numba), so there is nothing under /root/reference to compile into oracle/_ref/: the
reference's numba voxeliser is instead *imported* in the build container by
oracle/ref_voxel.py to generate tests/golden/voxel_*.npz.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "papc_oracle.c")
SRCS = [SRC, os.path.join(HERE, "nms_oracle.c")]
OUT = os.path.join(HERE, "libpapc_oracle.so")


def build(force=False, verbose=False):
    if (not force and os.path.exists(OUT)
            and all(os.path.getmtime(OUT) >= os.path.getmtime(f) for f in SRCS)):
        return OUT
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
           "-fno-fast-math", "-Wall", "-o", OUT, *SRCS, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
