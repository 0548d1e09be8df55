"""Generate tests/golden/nms_ref.npz by EXECUTING THE REFERENCE'S OWN numba.cuda kernels
(PAPC/models/detect/pointpillars/libs/ops/non_max_suppression/nms_gpu.py: nms_gpu :133-164, rotate_nms_gpu
:453-488, rotate_iou_gpu :518-553, rotate_iou_gpu_eval :603-653 and every device function they call) under numba's
CUDA simulator (NUMBA_ENABLE_CUDASIM=1 -- there is no GPU in the build container).  The module's pybind11 build of
nms.so (:8-19) is cut out with ``ast``; nothing else is modified.

What this pins: the reference's logic -- score order, the 64x64 mask tiling, `iou > thresh`, the suppress scan,
the polygon clipping / vertex sort / area of the rotated IoU.  The simulator evaluates scalar expressions in
Python (float64 intermediates, float32 stores), the real kernels in fp32 with NVVM's FMA contraction, so IoU
values are references to ~1e-6; keep lists are compared exactly (the cases are generated with no IoU within
1e-4 of the threshold).  Build-container only:  python tests/golden/make_golden_nms.py
"""
import ast
import os
import sys

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/PAPC/models/detect/pointpillars/libs/ops/non_max_suppression/nms_gpu.py"


def load_reference():
    tree = ast.parse(open(SRC).read())
    body = []
    for n in tree.body:
        if isinstance(n, ast.Try):
            continue                                     # the nms.so build / import
        if isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("libs."):
            continue
        if isinstance(n, ast.FunctionDef) and n.name == "nms_gpu_cc":
            continue                                     # wrapper of the pybind11 module
        body.append(n)
    ns = {}
    exec(compile(ast.Module(body=body, type_ignores=[]), SRC, "exec"), ns)
    return ns


def boxes_2d(rng, n, extent=60.0):
    c = rng.uniform(0, extent, (n, 2))
    wh = rng.uniform(2.0, 12.0, (n, 2))
    s = rng.uniform(0.05, 1.0, (n, 1))
    return np.concatenate([c - wh / 2, c + wh / 2, s], 1).astype(np.float32)


def rboxes(rng, n, extent=40.0):
    c = rng.uniform(0, extent, (n, 2))
    wh = rng.uniform(1.5, 6.0, (n, 2))
    a = rng.uniform(-np.pi, np.pi, (n, 1))
    return np.concatenate([c, wh, a], 1).astype(np.float32)


if __name__ == "__main__":
    R = load_reference()
    rng = np.random.default_rng(2024)
    out = {}
    # axis-aligned NMS: 150 boxes (3 column blocks of 64), two thresholds
    d = boxes_2d(rng, 150)
    out["nms_dets"] = d
    for t in (0.3, 0.6):
        out[f"nms_keep_{t}"] = np.asarray(R["nms_gpu"](d, np.float32(t)), np.int32)
    # rotated NMS: 100 boxes + score
    rb = rboxes(rng, 100)
    rd = np.concatenate([rb, rng.uniform(0.05, 1.0, (100, 1)).astype(np.float32)], 1)
    out["rnms_dets"] = rd
    for t in (0.1, 0.4):
        out[f"rnms_keep_{t}"] = np.asarray(R["rotate_nms_gpu"](rd, np.float32(t)), np.int32)
    # rotated IoU matrix 70 x 45 (crosses the 64-wide tile edge), all four criteria; plus special pairs
    a, b = rboxes(rng, 70), rboxes(rng, 45)
    b[:5] = a[:5]                                        # identical boxes -> IoU 1
    b[5, :2] = a[5, :2]; b[5, 4] = a[5, 4] + np.float32(np.pi / 2)   # same centre, turned by 90 degrees
    a[6] = [10, 10, 4, 2, 0]; b[6] = [10, 10, 2, 1, 0]   # contained box
    a[7] = [0, 0, 2, 2, 0]; b[7] = [100, 100, 2, 2, 0.3]  # disjoint
    out["riou_boxes"], out["riou_query"] = a, b
    out["riou"] = R["rotate_iou_gpu"](a, b)
    for crit in (-1, 0, 1, 2):
        out[f"riou_eval_{crit}"] = R["rotate_iou_gpu_eval"](a, b, crit)
    np.savez_compressed(os.path.join(HERE, "nms_ref.npz"), **out)
    for k, v in out.items():
        print(f"{k:16s} {str(np.asarray(v).dtype):8s} {np.asarray(v).shape}")
