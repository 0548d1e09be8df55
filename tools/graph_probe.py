"""Does capturing the whole SSG SetAbstraction step in a CUDA graph pay?  Eager vs graph replay."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from papc_b200 import sa_stack, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
B = bench.B_PER_GPU
xyz = torch.from_numpy(synth.clouds(B, bench.N_POINTS, seed=0)).to(dev)
st1 = torch.from_numpy(synth.fps_start(B, bench.N_POINTS, seed=1)).to(dev)
st2 = torch.zeros(B, dtype=torch.int64, device=dev)
model = sa_stack.SSGSetAbstractionStack().to(dev)
for i, sa in enumerate(model.layers_()):
    sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(bench.SA_CFG[i][3], bench.SA_CFG[i][4], seed=2 + i))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    return model(xyz, None, start_idx=(st1, st2))[1]


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


ref = step().clone()
print("eager  ms/step", timed(step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        out = step()
    g.replay()
    torch.cuda.synchronize()
    print("graph output equal:", bool(torch.equal(out, ref)))
    print("graph  ms/step", timed(g.replay))
except Exception as e:
    print("capture failed:", repr(e)[:300])
