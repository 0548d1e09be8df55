"""Run a few passes of the bench step (sa1 -> sa2 -> sa3, B=32 x 1024 points) and nothing else:
the short command ncu captures are taken from.  usage: python tools/prof_step.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from papc_b200 import sa_stack, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
B = bench.B_PER_GPU
xyz = torch.from_numpy(synth.clouds(B, bench.N_POINTS, seed=0)).to(dev)
st1 = torch.from_numpy(synth.fps_start(B, bench.N_POINTS, seed=1)).to(dev)
st2 = torch.zeros(B, dtype=torch.int64, device=dev)
model = sa_stack.SSGSetAbstractionStack().to(dev)
for i, sa in enumerate(model.layers_()):
    sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(bench.SA_CFG[i][3], bench.SA_CFG[i][4], seed=2 + i))
for _ in range(steps):
    _, l3 = model(xyz, None, start_idx=(st1, st2))
torch.cuda.synchronize()
print("ok", tuple(l3.shape), float(l3.abs().mean()))
