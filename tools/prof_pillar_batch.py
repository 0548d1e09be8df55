"""Per-kernel times of the batched pillar encode (voxelise batch -> PFN -> scatter), launch profiler + ncu-friendly.
usage: python tools/prof_pillar_batch.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from papc_b200 import _lib as L, pillars, synth, sa_stack
dev = torch.device("cuda:0")
NF = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = [torch.from_numpy(synth.lidar_frame(20000, seed=s)).to(dev) for s in range(NF)]
vs, rg, T, MV = synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, synth.KITTI_MAX_POINTS, synth.KITTI_MAX_VOXELS
pfn = pillars.PillarFeatureNet(num_input_features=4, use_norm=True, num_filters=(64,), with_distance=False, voxel_size=vs, pc_range=rg).to(dev)
scatter = pillars.PointPillarsScatter(output_shape=[1, 1, 496, 432], num_input_features=64)

def encode():
    v, c, n, fv, total = pillars.points_to_voxel_batch_device(frames, vs, rg, T, True, MV)
    feats = pfn(v, n, c, num_valid=total)
    return scatter(feats, c, NF, num_valid=total)

for _ in range(3): encode()
torch.cuda.synchronize()
lib = L.lib()
L.check(lib.papc_prof_reset(), "r"); L.check(lib.papc_prof_enable(1), "e")
for _ in range(5): encode()
torch.cuda.synchronize()
L.check(lib.papc_prof_enable(0), "e")
agg = {}
for r in L.prof_records():
    agg.setdefault(r["name"], []).append(r["ms"])
tot = 0
for k, v in agg.items():
    ms = float(np.mean(v)); tot += ms * len(v) / 5
    print(f"{k:28s} {1e3*ms:8.1f} us x{len(v)//5}")
print(f"sum {1e3*tot:.1f} us per call = {1e3*tot/NF:.1f} us/frame (eager, event-bracketed)")
g = sa_stack.GraphedForward(lambda *_: encode(), frames[0])
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1)/20*1e3:.1f} us per call = {e0.elapsed_time(e1)/20*1e3/NF:.1f} us/frame, total pillars {int(pillars.points_to_voxel_batch_device(frames, vs, rg, T, True, MV)[4].item())}")

# ---- per-stage graph replays (full clocks, no host gaps)
def stage_time(fn, name, reps=30):
    g = sa_stack.GraphedForward(lambda *_: fn(), frames[0])
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"  stage {name:10s} {e0.elapsed_time(e1)/reps*1e3:7.1f} us per call ({g.kernels_per_replay} library launches)")

v, c, n, fv, total = pillars.points_to_voxel_batch_device(frames, vs, rg, T, True, MV)
feats = pfn(v, n, c, num_valid=total)
stage_time(lambda: pillars.points_to_voxel_batch_device(frames, vs, rg, T, True, MV), "voxelise")
stage_time(lambda: pfn(v, n, c, num_valid=total), "pfn")
stage_time(lambda: scatter(feats, c, NF, num_valid=total), "scatter")
