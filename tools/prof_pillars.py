"""One KITTI-shaped frame through the pillar encode (voxelise -> PillarFeatureNet -> scatter), twice:
the short command ncu captures of the pillar kernels are taken from."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from papc_b200 import pillars, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
pts = torch.from_numpy(synth.lidar_frame(20000, seed=0)).to(dev)
pfn = pillars.PillarFeatureNet(num_input_features=4, use_norm=True, num_filters=(64,), with_distance=False,
                               voxel_size=synth.KITTI_VOXEL_SIZE, pc_range=synth.KITTI_PC_RANGE).to(dev)
scatter = pillars.PointPillarsScatter(output_shape=[1, 1, 496, 432], num_input_features=64)
for _ in range(2):
    v, c, n, vn = pillars.points_to_voxel_device(pts, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, 100, True, 12000)
    m = int(vn.item())
    coors4 = torch.cat([torch.zeros((m, 1), dtype=torch.int32, device=dev), c[:m]], 1).contiguous()
    feats = pfn(v[:m], n[:m], coors4)
    canvas = scatter(feats, coors4, 1)
torch.cuda.synchronize()
print("ok", m, tuple(canvas.shape), float(canvas.abs().sum()))
