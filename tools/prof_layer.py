"""Run selected grouped-MLP layer launches of the bench workload (for ncu captures).
usage: python tools/prof_layer.py [names substring ...]   e.g.  sa2.l2 sa1.l3"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from papc_b200 import _lib, layers, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
want = sys.argv[1:]
under_ncu = os.environ.get("PAPC_PROF_NCU") == "1"   # one warm-up + one launch per selected layer
res = bench.layer_roofline(torch, _lib.lib(), layers, synth, dev, bench.B_PER_GPU, only=want or None,
                           reps=1 if under_ncu else 5, warm=1 if under_ncu else 3)
for r in res["layers"]:
    if True:
        print(f"{r['name']:40s} {r['ms'] * 1e3:9.1f} us  {r['tflops']:7.1f} TFLOP/s")
