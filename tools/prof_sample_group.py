"""Time papc_sample_group_f32 against papc_fps_f32 / ball query / point_moments alone (CUDA events, L2 flushed).
usage: python tools/prof_sample_group.py   (PAPC_LIB=... for the triage build; PAPC_SG_DBG=1|2 there)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from papc_b200 import layers, synth
DEV = "cuda:0"
B, N, S, K, r = 32, 1024, 512, 32, 0.2
xyz = torch.from_numpy(np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1))).to(DEV)
st = torch.from_numpy(synth.fps_start(B, N, seed=1)).to(DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(t))

print("fps alone          %.1f us" % timeit(lambda: layers.farthest_point_sample_idx(xyz, S, st, return_xyz=True)))
_, nx = layers.farthest_point_sample_idx(xyz, S, st, return_xyz=True)
print("ball query alone   %.1f us" % timeit(lambda: layers._ball_query(r, K, xyz, nx, torch.int32)))
print("fused sample_group %.1f us (PAPC_SG_DBG=%s)" % (timeit(lambda: layers._sample_group_fused(xyz, S, st, r, K, True)), os.environ.get("PAPC_SG_DBG", "")))
