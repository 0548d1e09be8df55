"""Host-side mirror of the PointPillars pillar encode over the sm_100a kernels.

Mirrors (paths under /root/reference/PAPC/models/detect/pointpillars/):
  libs/ops/point_cloud/point_cloud_ops.py:106-166   points_to_voxel
  core/voxel_generator.py:5-43                      VoxelGenerator
  models/bones/pillars.py:9-41, 43-108, 110-142     PFNLayer, PillarFeatureNet, PointPillarsScatter
Same names, argument order, shapes and dtypes.  There is no CPU path: NumPy inputs are copied to
the current CUDA device, processed by the CUDA kernels and copied back (the reference's
``points_to_voxel`` is a NumPy-in / NumPy-out function called from the data loader); torch CUDA
inputs stay on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .layers import _ws


def _geom(voxel_size, coors_range):
    vs = np.asarray(voxel_size, dtype=np.float32).reshape(3)
    cr = np.asarray(coors_range, dtype=np.float32).reshape(6)
    return (C.c_float * 3)(*vs.tolist()), (C.c_float * 6)(*cr.tolist()), vs, cr


def points_to_voxel_device(points, voxel_size, coors_range, max_points=35, reverse_index=True,
                           max_voxels=20000):
    """Device-resident form: points [N,F] CUDA -> (voxels [max_voxels,max_points,F], coors
    [max_voxels,3] int32, num_points [max_voxels] int32, voxel_num int32[1]) with no host sync;
    rows >= voxel_num are zero.  Feeds PillarFeatureNet / PointPillarsScatter through their
    ``num_valid`` argument."""
    L.require_cuda(points)
    points = L.f32c(points)
    if points.dim() != 2 or points.shape[1] < 3:
        raise ValueError("points must be [N, F>=3]")
    N, F = points.shape
    dev = points.device
    vs_c, cr_c, _, _ = _geom(voxel_size, coors_range)
    lib = L.lib()
    voxels = torch.empty((max_voxels, max_points, F), dtype=torch.float32, device=dev)
    coors = torch.empty((max_voxels, 3), dtype=torch.int32, device=dev)
    num = torch.empty((max_voxels,), dtype=torch.int32, device=dev)
    vnum = torch.empty((1,), dtype=torch.int32, device=dev)
    wsb = lib.papc_voxelize_workspace_bytes(N, vs_c, cr_c, max_voxels)
    if wsb == 0:
        raise ValueError("points_to_voxel: bad voxel_size / coors_range / max_voxels")
    ws = _ws(wsb, dev)
    L.check(lib.papc_voxelize_f32(L.ptr(points), N, F, vs_c, cr_c, int(max_points),
                                  1 if reverse_index else 0, int(max_voxels), L.ptr(voxels),
                                  L.ptr(coors), L.ptr(num), L.ptr(vnum), L.ptr(ws), wsb,
                                  L.stream_ptr(dev)), "points_to_voxel")
    return voxels, coors, num, vnum


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True,
                    max_voxels=20000):
    """point_cloud_ops.py:106-166.  NumPy in -> NumPy out (like the reference), or torch CUDA in ->
    torch CUDA out; returns (voxels [M,max_points,F], coordinates [M,3] int32, num_points [M] int32)."""
    is_np = isinstance(points, np.ndarray)
    if is_np:
        if not torch.cuda.is_available():
            raise L.PapcError("points_to_voxel needs a CUDA device (no CPU fallback)")
        pts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).cuda()
    else:
        pts = points
    voxels, coors, num, vnum = points_to_voxel_device(pts, voxel_size, coors_range, max_points,
                                                      reverse_index, max_voxels)
    m = int(vnum.item())  # the reference's return shapes depend on voxel_num (pc_ops.py:161-163)
    voxels, coors, num = voxels[:m], coors[:m], num[:m]
    if is_np:
        return voxels.cpu().numpy(), coors.cpu().numpy(), num.cpu().numpy()
    return voxels, coors, num


class VoxelGenerator:
    """core/voxel_generator.py:5-43."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = max_num_points
        self._max_voxels = max_voxels
        self._grid_size = grid_size

    def generate(self, points, max_voxels):
        return points_to_voxel(points, self._voxel_size, self._point_cloud_range,
                               self._max_num_points, True, max_voxels)

    def generate_device(self, points, max_voxels):
        return points_to_voxel_device(points, self._voxel_size, self._point_cloud_range,
                                      self._max_num_points, True, max_voxels)

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size


class PFNLayer(torch.nn.Module):
    """pillars.py:9-41.  Holds Linear(in, out, bias=not use_norm) + BatchNorm1D(eps 1e-3, momentum
    .01) parameters; the compute happens fused inside PillarFeatureNet.forward."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        self.name = "PFNLayer"
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        self.use_norm = use_norm
        # paddle.nn.Linear stores weight as [in_features, out_features]
        bound = 1.0 / np.sqrt(in_channels)
        self.weight = torch.nn.Parameter(torch.empty(in_channels, out_channels).uniform_(-bound, bound),
                                         requires_grad=False)
        self.bias = None if use_norm else torch.nn.Parameter(torch.zeros(out_channels), requires_grad=False)
        self.bn_weight = torch.nn.Parameter(torch.ones(out_channels), requires_grad=False)
        self.bn_bias = torch.nn.Parameter(torch.zeros(out_channels), requires_grad=False)
        self.register_buffer("_mean", torch.zeros(out_channels))
        self.register_buffer("_variance", torch.ones(out_channels))
        self.epsilon = 1e-3
        self.momentum = 0.01


class PillarFeatureNet(torch.nn.Module):
    """pillars.py:43-108.  ``forward(features [P,T,F], num_voxels [P], coors [P,4]) -> [P,C]``.

    Properly registered in the reference (``nn.LayerList``, :71), so ``train()`` uses batch
    statistics over all P*T rows (padding rows included) and ``eval()`` the running ones."""

    def __init__(self, num_input_features=4, use_norm=True, num_filters=(64, 128), with_distance=False,
                 voxel_size=(0.2, 0.2, 4), pc_range=(0, -40, -3, 70.4, 40, 1)):
        super().__init__()
        self.name = "PillarFeatureNet"
        assert len(num_filters) > 0
        num_input_features += 5
        if with_distance:
            num_input_features += 1
        self._with_distance = with_distance
        num_filters = [num_input_features] + list(num_filters)
        layers_ = []
        for i in range(len(num_filters) - 1):                                      # :62-70
            last_layer = not (i < len(num_filters) - 2)
            layers_.append(PFNLayer(num_filters[i], num_filters[i + 1], use_norm, last_layer=last_layer))
        self.pfn_layers = torch.nn.ModuleList(layers_)
        self.vx = voxel_size[0]
        self.vy = voxel_size[1]
        self.x_offset = self.vx / 2 + pc_range[0]
        self.y_offset = self.vy / 2 + pc_range[1]
        self.update_running_stats = False

    def _bn_mode(self, pfn):
        if not pfn.use_norm:
            return L.BN_NONE
        return L.BN_BATCH if self.training else L.BN_RUNNING

    def forward(self, features, num_voxels, coors, num_valid=None):
        L.require_cuda(features, num_voxels, coors)
        features = L.f32c(features)
        P, T, F = features.shape
        num_voxels = num_voxels.to(torch.int32).contiguous()
        coors = coors.to(torch.int32).contiguous()
        if coors.shape != (P, 4):
            raise ValueError("coors must be [P,4] (batch, z, y, x)")
        extra = 5 + (1 if self._with_distance else 0)
        if self.pfn_layers[0].weight.shape[0] != F + extra:
            raise ValueError(f"PFN expects {self.pfn_layers[0].weight.shape[0] - extra} point features, got {F}")
        dev = features.device
        lib = L.lib()
        f32 = np.float32
        geom = (float(f32(self.vx)), float(f32(self.vy)), float(f32(self.x_offset)), float(f32(self.y_offset)))
        if len(self.pfn_layers) == 1 and not self._with_distance:
            # the yaml shape (num_filters=[64]): decoration + Linear + BatchNorm + ReLU + max in one fused pass
            pfn = self.pfn_layers[0]
            cout = pfn.units
            out = torch.empty((P, cout), dtype=torch.float32, device=dev)
            mode = self._bn_mode(pfn)
            want = self.update_running_stats and mode == L.BN_BATCH
            bm = torch.empty((cout,), dtype=torch.float32, device=dev) if want else None
            bv = torch.empty((cout,), dtype=torch.float32, device=dev) if want else None
            wsb = lib.papc_pfn_workspace_bytes(P, cout)
            ws = _ws(wsb, dev)
            L.check(lib.papc_pfn_f32(L.ptr(features), L.ptr(num_voxels), L.ptr(coors), P, T, F, *geom,
                                     L.ptr(L.f32c(pfn.weight)),
                                     L.ptr(L.f32c(pfn.bias)) if pfn.bias is not None else None,
                                     L.ptr(L.f32c(pfn.bn_weight)), L.ptr(L.f32c(pfn.bn_bias)),
                                     L.ptr(pfn._mean), L.ptr(pfn._variance), mode, float(pfn.epsilon), cout,
                                     L.ptr(num_valid), L.ptr(out), L.ptr(bm), L.ptr(bv), L.ptr(ws), wsb,
                                     L.stream_ptr(dev)), "PillarFeatureNet")
            if want:
                pfn._mean.mul_(pfn.momentum).add_(bm, alpha=1 - pfn.momentum)
                pfn._variance.mul_(pfn.momentum).add_(bv, alpha=1 - pfn.momentum)
            return out.squeeze()                                                  # :108
        # general chain (the reference's default num_filters=(64,128), with_distance, ...): one generic
        # PFNLayer launch group per layer, the intermediate [P,T,2u] tensors materialised as in the reference
        x, cin, decorate = features, F, 1
        for pfn in self.pfn_layers:                                               # :105-106
            u = pfn.units
            last = 1 if pfn.last_vfe else 0
            out = torch.empty((P, u) if last else (P, T, 2 * u), dtype=torch.float32, device=dev)
            mode = self._bn_mode(pfn)
            want = self.update_running_stats and mode == L.BN_BATCH
            bm = torch.empty((u,), dtype=torch.float32, device=dev) if want else None
            bv = torch.empty((u,), dtype=torch.float32, device=dev) if want else None
            wsb = lib.papc_pfn_layer_workspace_bytes(P, T, u)
            ws = _ws(wsb, dev)
            L.check(lib.papc_pfn_layer_f32(L.ptr(x), decorate, 1 if self._with_distance else 0, L.ptr(num_voxels),
                                           L.ptr(coors), P, T, cin, *geom, L.ptr(L.f32c(pfn.weight)),
                                           L.ptr(L.f32c(pfn.bias)) if pfn.bias is not None else None,
                                           L.ptr(L.f32c(pfn.bn_weight)), L.ptr(L.f32c(pfn.bn_bias)),
                                           L.ptr(pfn._mean), L.ptr(pfn._variance), mode, float(pfn.epsilon), u, last,
                                           L.ptr(num_valid), L.ptr(out), L.ptr(bm), L.ptr(bv), L.ptr(ws), wsb,
                                           L.stream_ptr(dev)), "PFNLayer")
            if want:
                pfn._mean.mul_(pfn.momentum).add_(bm, alpha=1 - pfn.momentum)
                pfn._variance.mul_(pfn.momentum).add_(bv, alpha=1 - pfn.momentum)
            x, cin, decorate = out, 2 * u, 0
        return x.squeeze()                                                        # :108


class PointPillarsScatter(torch.nn.Module):
    """pillars.py:110-142.  ``forward(voxel_features [P,C], coords [P,4], batch_size) ->
    [batch_size, C, ny, nx]``."""

    def __init__(self, output_shape, num_input_features=4):
        super().__init__()
        self.name = "PointPillarsScatter"
        self.output_shape = output_shape
        self.ny = output_shape[2]
        self.nx = output_shape[3]
        self.nchannels = num_input_features

    def forward(self, voxel_features, coords, batch_size, num_valid=None):
        L.require_cuda(voxel_features, coords)
        voxel_features = L.f32c(voxel_features)
        if voxel_features.dim() == 1:
            voxel_features = voxel_features.reshape(1, -1)
        coords = coords.to(torch.int32).contiguous()
        P, Cc = voxel_features.shape
        if Cc != self.nchannels:
            raise ValueError(f"expected {self.nchannels} channels, got {Cc}")
        dev = voxel_features.device
        canvas = torch.empty((batch_size, Cc, self.ny, self.nx), dtype=torch.float32, device=dev)
        lib = L.lib()
        wsb = lib.papc_pillar_scatter_workspace_bytes(batch_size, self.ny, self.nx)
        ws = _ws(wsb, dev)
        L.check(lib.papc_pillar_scatter_f32(L.ptr(voxel_features), L.ptr(coords), P, Cc, batch_size,
                                            self.ny, self.nx, L.ptr(num_valid), L.ptr(canvas),
                                            L.ptr(ws), wsb, L.stream_ptr(dev)), "PointPillarsScatter")
        return canvas


def merge_coordinates(coors_list):
    """The 'coordinates' branch of merge_second_batch (pp/data/preprocess.py:30-38): prepend the
    sample index -> [P,4] (b,z,y,x).  Tiny glue on torch tensors."""
    out = []
    for i, c in enumerate(coors_list):
        pad = torch.full((c.shape[0], 1), i, dtype=c.dtype, device=c.device)
        out.append(torch.cat([pad, c], dim=1))
    return torch.cat(out, dim=0)
