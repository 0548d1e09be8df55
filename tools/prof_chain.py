"""prof_chain.py -- a few forward passes of the SSG sa1 / sa2 SetAbstraction layers (B = 32) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from papc_b200 import layers, synth
DEV = "cuda:0"
which = sys.argv[1:] or ["sa1", "sa2"]
rng = np.random.default_rng(7)
B = 32
if "sa1" in which:
    g = layers.PointNetSetAbstraction(512, 0.2, 32, 3, [64, 64, 128], False).to(DEV)
    xyz = torch.from_numpy(synth.clouds(B, 1024, seed=3)).to(DEV)
    st = torch.from_numpy(synth.fps_start(B, 1024, seed=4)).to(DEV)
    for _ in range(3):
        g(xyz, None, start_idx=st)
if "sa2" in which:
    g = layers.PointNetSetAbstraction(128, 0.4, 64, 131, [128, 128, 256], False).to(DEV)
    xyz = torch.from_numpy(synth.clouds(B, 512, seed=3)).to(DEV)
    feats = torch.from_numpy(np.maximum(rng.standard_normal((B, 128, 512)), 0).astype(np.float32)).to(DEV)
    st = torch.from_numpy(synth.fps_start(B, 512, seed=4)).to(DEV)
    for _ in range(3):
        g(xyz, feats, start_idx=st)
torch.cuda.synchronize()
