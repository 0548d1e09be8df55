"""Forward times of the other BASELINE.json configurations (parity-test cases, not bench lines):
  C3  PointNet++ MSG segment SetAbstraction stack, B=16 x 2048 points (3-radius ball query)
  C4  PointPillars pillar encode: voxelise -> PillarFeatureNet -> scatter, 20k-point KITTI-shaped frames
CUDA events around 30 back-to-back passes after 5 warm-up passes.  usage: python tools/bench_configs.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from papc_b200 import pillars, sa_stack, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)


def timeit(fn, reps=30, warm=5):
    """ms per call, reps calls queued back to back between two events (keeps the GPU busy, so the SM
    clock does not fall back to idle between the small pillar kernels)."""
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


out = {}
# ---- C3: MSG segment stack (features = xyz, segment/pointnet2/pointnet2.py:82-86)
B, N = 16, 2048
xyz = torch.from_numpy(synth.clouds(B, N, seed=0)).to(dev)
st1 = torch.from_numpy(synth.fps_start(B, N, seed=1)).to(dev)
st2 = torch.zeros(B, dtype=torch.int64, device=dev)
msg = sa_stack.MSGSegSetAbstractionStack().to(dev)
ms = timeit(lambda: msg(xyz, xyz, start_idx=(st1, st2)))
out["C3_msg_seg_B16_N2048"] = {"ms_per_forward": ms, "points_per_s": B * N / (ms / 1e3), "gflop": 142.6,
                               "tflops": 142.6e9 / (ms / 1e3) / 1e12}

# ---- C3 + the decoder: sa1..sa3 -> fp3 -> fp2 -> fp1
seg = sa_stack.MSGSegEncoderDecoder().to(dev)
onehot = torch.zeros((B, 16, N), device=dev)
onehot[:, 3, :] = 1.0
ms2 = timeit(lambda: seg(xyz, onehot, start_idx=(st1, st2)))
out["C3_msg_seg_with_feature_propagation"] = {"ms_per_forward": ms2, "points_per_s": B * N / (ms2 / 1e3),
                                              "decoder_ms": ms2 - ms}

# ---- C4: pillar encode, 2 frames of 20 000 points (yaml geometry), device-resident chain
frames = [torch.from_numpy(synth.lidar_frame(20000, seed=s)).to(dev) for s in (0, 1)]
pfn = pillars.PillarFeatureNet(num_input_features=4, use_norm=True, num_filters=(64,), with_distance=False,
                               voxel_size=synth.KITTI_VOXEL_SIZE, pc_range=synth.KITTI_PC_RANGE).to(dev)
scatter = pillars.PointPillarsScatter(output_shape=[1, 1, 496, 432], num_input_features=64)


def pillar_step():
    res = []
    for f in frames:
        v, c, n, vn = pillars.points_to_voxel_device(f, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, 100, True, 12000)
        res.append((v, c, n, vn))
    return res


t_vox = timeit(pillar_step)
out["C4_voxelise_2x20k"] = {"ms": t_vox, "points_per_s": 40000 / (t_vox / 1e3),
                            "GBs_algorithmic": 2 * (0.32e6 + 19.4e6) / (t_vox / 1e3) / 1e9}
try:
    v, c, n, vn = pillar_step()[0]
    m = int(vn.item())
    coors4 = torch.cat([torch.zeros((m, 1), dtype=torch.int32, device=dev), c[:m]], 1).contiguous()
    feats = pfn(v[:m], n[:m], coors4)
    t_pfn = timeit(lambda: pfn(v[:m], n[:m], coors4))
    t_sc = timeit(lambda: scatter(feats, coors4, 1))
    out["C4_pfn_1frame"] = {"ms": t_pfn, "pillars": m, "GBs_algorithmic": (m * 1616 + m * 256) / (t_pfn / 1e3) / 1e9}
    out["C4_scatter_1frame"] = {"ms": t_sc, "GBs_algorithmic": 54.9e6 / (t_sc / 1e3) / 1e9}
except Exception as e:  # keep the voxeliser / MSG numbers even if the PFN wrapper signature differs
    out["C4_pfn_scatter_error"] = repr(e)
print(json.dumps(out, indent=1))
