"""Build papc_b200/lib/libpapc_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

The library is a plain C-ABI shared object (include/papc_b200.h): no torch, no pybind11.  It
is built IN-TREE so it travels to the GPU box with the gpurun snapshot (git-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_DIR = os.path.join(os.path.dirname(HERE), "lib")
OUT = os.path.join(OUT_DIR, "libpapc_b200.so")
SOURCES = ["capi.cu", "fps.cu", "ball_query.cu", "sa_mlp.cu", "sa_mlp_tc.cu", "sa_mlp_tt.cu", "sa_chain.cu", "pillars.cu", "pillar_batch.cu", "feature_prop.cu", "nms.cu"]
# per-file flags.  fps.cu: the distance is pinned as separately rounded multiplies and adds.
EXTRA = {"fps.cu": ["-fmad=false"], "nms.cu": ["-fmad=false"]}   # nms.cu: IoU arithmetic pinned like oracle/nms_oracle.c
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", HERE]
FLAGS += os.environ.get("PAPC_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DPAPC_TT_PROD_WARPS=16


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(HERE, "common.cuh"), os.path.join(ROOT, "include", "papc_b200.h"),
            os.path.abspath(__file__)]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    obj_dir = os.path.join(OUT_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        obj = os.path.join(obj_dir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, src):
            cmd = [NVCC, *FLAGS, *EXTRA.get(s, []), "-c", src, "-o", obj]
            if ptxas_info:
                cmd += ["-Xptxas", "-v"]
            if verbose:
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd)))
    failed = [s for s, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for " + ", ".join(failed))
    stale_link = not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs)
    if procs or force or stale_link:
        cmd = [NVCC, "-shared", "-o", OUT, *objs, "-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
