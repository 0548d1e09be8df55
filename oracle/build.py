"""Compile oracle/papc_oracle.c -> oracle/libpapc_oracle.so (gcc, no GPU needed), and stage the reference's
own runnable files under oracle/_ref/.

TEST INFRASTRUCTURE.  Called by __graft_entry__.build(); the .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  The reference itself is pure Python (+numba): nothing under /root/reference compiles.
``build_ref()`` is the committed recipe that copies the few reference files the checker and bench.py's CPU
baseline EXECUTE UNMODIFIED (the numba voxeliser, the PointNet++ layers file, the two model files, the CPU NMS)
from where they lie under /root/reference into oracle/_ref/ -- git-ignored (never in history), not
gpurun-ignored (so it travels to the GPU box, where /root/reference does not exist).  Nothing in papc_b200/
imports from it.
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


REF_ROOT = "/root/reference/PAPC/models/"
REF_DIR = os.path.join(HERE, "_ref")
REF_FILES = {   # destination under oracle/_ref/ -> source under /root/reference/PAPC/models/
    "point_cloud_ops.py": "detect/pointpillars/libs/ops/point_cloud/point_cloud_ops.py",
    "bev_ops.py": "detect/pointpillars/libs/ops/point_cloud/bev_ops.py",
    "pointnet2_basic_layers.py": "layers/pointnet2_basic_layers.py",
    "classify_pointnet2.py": "classify/pointnet2/pointnet2.py",
    "segment_pointnet2.py": "segment/pointnet2/pointnet2.py",
    "nms_gpu.py": "detect/pointpillars/libs/ops/non_max_suppression/nms_gpu.py",
}


def build_ref(verbose=False):
    """Stage the reference files (see the module docstring).  No-op where /root/reference is absent (the GPU box
    uses what the snapshot brought).  Returns the directory, or None if nothing is there."""
    import shutil
    if os.path.isdir(REF_ROOT):
        os.makedirs(REF_DIR, exist_ok=True)
        for dst, src in REF_FILES.items():
            shutil.copyfile(os.path.join(REF_ROOT, src), os.path.join(REF_DIR, dst))
            if verbose:
                print("staged", src, "->", os.path.join("oracle/_ref", dst))
    return REF_DIR if all(os.path.exists(os.path.join(REF_DIR, f)) for f in REF_FILES) else None


def ref_file(name):
    """Path of a staged reference file, or None."""
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


if __name__ == "__main__":
    build_ref(verbose=True)
    print(build(force="--force" in sys.argv, verbose=True))
