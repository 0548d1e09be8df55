"""Compare papc_b200.nms with the reference's OWN numba.cuda kernels running on this GPU (oracle/_ref/nms_gpu.py,
unmodified): rotated IoU matrices (max |diff|, entries that differ bitwise) and, for a rotated NMS whose keep lists
differ, the first decision that differs and how far its IoU is from the threshold.
usage (GPU box): python tools/nms_vs_reference.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import build as ob, ref_nms
from papc_b200 import nms as pnms

ref = ref_nms.load(ob.ref_file("nms_gpu.py"))
rng = np.random.default_rng(5)
def rdets(n):
    c = rng.uniform(0, 70.0, (n, 2)); wh = rng.uniform(1.5, 4.5, (n, 2)); a = rng.uniform(-np.pi, np.pi, (n, 1)); s = rng.uniform(0.05, 1.0, (n, 1))
    return np.concatenate([c, wh, a, s], 1).astype(np.float32)
d1 = rdets(1000); d4 = rdets(4096)          # the same draws as bench.py's N3 block
for d in (d1, d4):
    n = d.shape[0]
    r = ref["rotate_iou_gpu"](d[:, :5].copy(), d[:, :5].copy())
    m = pnms.rotate_iou_gpu(d[:, :5].copy(), d[:, :5].copy())
    dg = np.arange(n)
    print(f"n={n}: IoU of a box with itself: reference {np.unique(np.round(r[dg, dg], 6))[:6]}, ours {np.unique(np.round(m[dg, dg], 6))[:6]}")
    r[dg, dg] = 0; m[dg, dg] = 0      # the identical-box quirk is looked at separately (line above)
    diff = np.abs(r - m)
    nz = (r != m)
    print(f"n={n}: IoU matrix max |diff| {diff.max():.3e}; {nz.sum()} of {n*n} entries differ bitwise ({(r>0).sum()} non-zero); "
          f"largest relative {np.max(diff[nz] / np.maximum(np.abs(r[nz]), 1e-30)) if nz.any() else 0:.3e}")
    kr = list(map(int, ref["rotate_nms_gpu"](d, np.float32(0.5))))
    km = list(map(int, pnms.rotate_nms_gpu(d, 0.5)))
    if kr == km:
        print(f"n={n}: keep lists identical ({len(kr)} kept)")
        continue
    i = next(i for i, (a, b) in enumerate(zip(kr, km)) if a != b)
    print(f"n={n}: keep lists differ from position {i}: reference keeps {kr[i]}, ours {km[i]}; lengths {len(kr)} / {len(km)}; "
          f"same SET: {set(kr) == set(km)}; scores {d[kr[i], 5]!r} vs {d[km[i], 5]!r}; boxes with a duplicated score: {n - len(np.unique(d[:, 5]))}")
    # the box one side dropped: IoU against the boxes kept before it
    for name, cand, iou in (("reference", kr[i], r), ("ours", km[i], m)):
        other = km if name == "reference" else kr
        if cand not in other[:i + 5]:
            prev = kr[:i]
            v_ref, v_our = r[cand, prev], m[cand, prev]
            j = int(np.argmax(np.maximum(v_ref, v_our)))
            print(f"   box {cand} (kept by {name} only): its largest IoU with an earlier kept box ({prev[j]}): reference {v_ref[j]:.9f}, ours {v_our[j]:.9f}, threshold 0.5")
