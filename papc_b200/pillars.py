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


# ------------------------------------------------------------------ N4: the steps either side of the encode
def points_to_voxel_batch_device(points_list, voxel_size, coors_range, max_points=35, reverse_index=True,
                                 max_voxels=20000):
    """``points_to_voxel`` over a list of frames + the merged batch layout of ``merge_second_batch``
    (pp/data/preprocess.py:16-42), device resident, no host sync: returns
    ``(voxels [B*max_voxels,T,F], coordinates [B*max_voxels,4] int32 (b,z,y,x), num_points [B*max_voxels] int32,
    num_voxels [B] int32, total int32[1])``; rows >= total are zero.  Feeds ``PillarFeatureNet`` /
    ``PointPillarsScatter`` through ``num_valid=total``."""
    if len(points_list) < 1:
        raise ValueError("points_to_voxel_batch: empty batch")
    L.require_cuda(*points_list)
    pts = [L.f32c(p) for p in points_list]
    F = pts[0].shape[1]
    for p in pts:
        if p.dim() != 2 or p.shape[1] != F or F < 3:
            raise ValueError("every frame must be [N_i, F>=3] with the same F")
    dev = pts[0].device
    B = len(pts)
    offs = np.zeros(B + 1, np.int32)
    offs[1:] = np.cumsum([p.shape[0] for p in pts])
    allp = torch.cat(pts, dim=0) if B > 1 else pts[0]
    vs_c, cr_c, _, _ = _geom(voxel_size, coors_range)
    lib = L.lib()
    rows = B * max_voxels
    voxels = torch.empty((rows, max_points, F), dtype=torch.float32, device=dev)
    coors = torch.empty((rows, 4), dtype=torch.int32, device=dev)
    num = torch.empty((rows,), dtype=torch.int32, device=dev)
    fv = torch.empty((B,), dtype=torch.int32, device=dev)
    total = torch.empty((1,), dtype=torch.int32, device=dev)
    wsb = lib.papc_voxelize_batch_workspace_bytes(int(offs[-1]), B, vs_c, cr_c, max_voxels)
    if wsb == 0:
        raise ValueError("points_to_voxel_batch: bad voxel_size / coors_range / max_voxels / batch (<= 32)")
    ws = _ws(wsb, dev)
    offs_c = (C.c_int32 * (B + 1))(*offs.tolist())
    L.check(lib.papc_voxelize_batch_f32(L.ptr(allp), offs_c, B, F, vs_c, cr_c, int(max_points),
                                        1 if reverse_index else 0, int(max_voxels), L.ptr(voxels), L.ptr(coors),
                                        L.ptr(num), L.ptr(fv), L.ptr(total), L.ptr(ws), wsb, L.stream_ptr(dev)),
            "points_to_voxel_batch")
    return voxels, coors, num, fv, total


def merge_second_batch_voxels(points_list, voxel_size, coors_range, max_points=35, reverse_index=True,
                              max_voxels=20000):
    """The 'voxels' / 'num_points' / 'coordinates' entries of ``merge_second_batch`` (pp/data/preprocess.py:16-42;
    'num_voxels' is popped there, :20) for a list of point clouds: NumPy in -> (dict of NumPy arrays with the
    reference's shapes, per-frame voxel counts); this form reads the total back, as the reference's return
    shapes require."""
    is_np = isinstance(points_list[0], np.ndarray)
    pl = [torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32)).cuda() if is_np else p for p in points_list]
    v, c, n, fv, total = points_to_voxel_batch_device(pl, voxel_size, coors_range, max_points, reverse_index, max_voxels)
    m = int(total.item())
    out = {"voxels": v[:m], "num_points": n[:m], "coordinates": c[:m]}
    if is_np:
        return {k: t.cpu().numpy() for k, t in out.items()}, fv.cpu().numpy()
    return out, fv


def sparse_sum_for_anchors_mask(coors, shape, num_valid=None):
    """box_np_ops.py:772-777: ``ret[coors[i,1], coors[i,2]] += 1`` over (z,y,x) rows -> float32 [ny,nx].
    Also accepts merged (b,z,y,x) rows (the last two columns are used).  torch CUDA in / out, or NumPy in / out."""
    is_np = isinstance(coors, np.ndarray)
    c = torch.from_numpy(np.ascontiguousarray(coors, dtype=np.int32)).cuda() if is_np else coors
    L.require_cuda(c)
    c = c.to(torch.int32).contiguous()
    ny, nx = int(shape[0]), int(shape[1])
    dense = torch.empty((ny, nx), dtype=torch.float32, device=c.device)
    L.check(L.lib().papc_anchors_mask_f32(L.ptr(c), c.shape[0], c.shape[1], L.ptr(num_valid), ny, nx, None, 0, None, None,
                                          None, L.ptr(dense), None, L.stream_ptr(c.device)), "sparse_sum_for_anchors_mask")
    return dense.cpu().numpy() if is_np else dense


def anchors_area_from_coors(coors, shape, anchors_bv, stride, offset, grid_size, num_valid=None):
    """The anchors-mask chain of the reference's target assigner in one call:
    ``dense = sparse_sum_for_anchors_mask(coors, shape).cumsum(0).cumsum(1)`` then
    ``fused_get_anchors_area(dense, anchors_bv, stride, offset, grid_size)`` (box_np_ops.py:772-806).
    Returns (dense cumulative map [ny,nx], anchors_area [N])."""
    is_np = isinstance(coors, np.ndarray)
    c = torch.from_numpy(np.ascontiguousarray(coors, dtype=np.int32)).cuda() if is_np else coors
    a = torch.from_numpy(np.ascontiguousarray(anchors_bv, dtype=np.float32)).cuda() if isinstance(anchors_bv, np.ndarray) else anchors_bv
    L.require_cuda(c, a)
    c = c.to(torch.int32).contiguous()
    a = L.f32c(a)
    ny, nx = int(shape[0]), int(shape[1])
    dense = torch.empty((ny, nx), dtype=torch.float32, device=c.device)
    area = torch.empty((a.shape[0],), dtype=torch.float32, device=c.device)
    st = (C.c_float * 2)(float(stride[0]), float(stride[1]))
    of = (C.c_float * 2)(float(offset[0]), float(offset[1]))
    gs = (C.c_int32 * 2)(int(grid_size[0]), int(grid_size[1]))
    L.check(L.lib().papc_anchors_mask_f32(L.ptr(c), c.shape[0], c.shape[1], L.ptr(num_valid), ny, nx, L.ptr(a), a.shape[0],
                                          st, of, gs, L.ptr(dense), L.ptr(area), L.stream_ptr(c.device)),
            "fused_get_anchors_area")
    if is_np:
        return dense.cpu().numpy(), area.cpu().numpy()
    return dense, area


def points_to_bev(points, voxel_size, coors_range, with_reflectivity=False, density_norm_num=16, max_voxels=40000):
    """bev_ops.py:61-103: points [N,4] -> bev map [D+1(+1), H, W] (per-slice maximum normalised height,
    optional intensity map, point-count map).  NumPy in -> NumPy out or torch CUDA in -> out."""
    is_np = isinstance(points, np.ndarray)
    p = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).cuda() if is_np else points
    L.require_cuda(p)
    p = L.f32c(p)
    vs_c, cr_c, vs, cr = _geom(voxel_size, coors_range)
    shape = tuple(np.round((cr[3:] - cr[:3]) / vs).astype(np.int32).tolist())[::-1]   # DHW (:84-86)
    D, H, W = shape
    lowers = np.linspace(cr[2], cr[5], D, endpoint=False).astype(np.float32)            # (:89-90)
    hl = (C.c_float * D)(*lowers.tolist())
    nm = D + 1 + (1 if with_reflectivity else 0)
    bev = torch.empty((nm, H, W), dtype=torch.float32, device=p.device)
    lib = L.lib()
    wsb = lib.papc_points_to_bev_workspace_bytes(p.shape[0], vs_c, cr_c)
    ws = _ws(wsb, p.device)
    L.check(lib.papc_points_to_bev_f32(L.ptr(p), p.shape[0], p.shape[1], vs_c, cr_c, hl, 1 if with_reflectivity else 0,
                                       int(max_voxels), L.ptr(bev), L.ptr(ws), wsb, L.stream_ptr(p.device)), "points_to_bev")
    return bev.cpu().numpy() if is_np else bev
