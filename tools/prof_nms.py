"""Time the N3 block of the bench (rotated / axis-aligned NMS, device resident) on its own.
usage (GPU box): python tools/prof_nms.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)


def timeit(fn, reps=20, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = bench.nms_block(torch, dev, timeit)
for k, v in out.items():
    if isinstance(v, dict):
        print(k, "ms", round(v["ms"], 4), "kept", v["kept"], "reference", (v.get("reference") or {}).get("ms"),
              "same_keep_list", (v.get("reference") or {}).get("same_keep_list"))
    else:
        print(k, v)
