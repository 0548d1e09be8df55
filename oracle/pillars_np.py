"""NumPy restatement of PAPC's PointPillars pillar encode -- TEST INFRASTRUCTURE, NOT PRODUCT.

Follows (paths under /root/reference/PAPC/models/detect/pointpillars/):
  libs/ops/point_cloud/point_cloud_ops.py:7-166   points_to_voxel (numba)
  core/voxel_generator.py:5-43                    VoxelGenerator
  libs/tools/__init__.py:26-35                    get_paddings_indicator
  models/bones/pillars.py:9-41, 43-108, 110-142   PFNLayer, PillarFeatureNet, PointPillarsScatter
  libs/functional.py:21-38                        mask_select, select_change
  data/preprocess.py:16-42                        merge_second_batch ('coordinates' branch)
  libs/ops/box_np_ops.py:772-806                  sparse_sum_for_anchors_mask, fused_get_anchors_area (N4)
  libs/ops/point_cloud/bev_ops.py:6-103           points_to_bev (N4)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.

Parity status: ``points_to_voxel`` is PINNED -- the reference's numba kernel imports and
runs in the build container (only needs numba + numpy), and tests/golden/voxel_*.npz hold
its outputs (tests/golden/make_golden.py).  PillarFeatureNet / PointPillarsScatter: their LOGIC
is pinned to vectors produced by executing the reference's own classes over a NumPy stand-in for
paddle (tests/golden/make_golden_pillars.py -> pillars_ref.npz; decoration and canvas bit-exact,
tests/test_oracle_vs_reference_source.py); the Linear / BatchNorm1D arithmetic remains "parity
unpinned" (it executes inside PaddlePaddle, absent here; no reference tests).  The N4 pieces are PINNED: the
reference's own functions (plain NumPy / numba) run as they are in tests/golden/make_golden_pillar_batch.py ->
pillar_batch_ref.npz.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True,
                    max_voxels=20000):
    """point_cloud_ops.py:106-166 (wrapper) + :7-53 / :55-103 (loop), in pure Python.

    Slow (Python loop); the C restatement ``oracle_voxelize_f32`` is the fast twin and the
    tests check both against the golden vectors.
    """
    if not isinstance(voxel_size, np.ndarray):
        voxel_size = np.array(voxel_size, dtype=points.dtype)
    if not isinstance(coors_range, np.ndarray):
        coors_range = np.array(coors_range, dtype=points.dtype)
    voxelmap_shape = (coors_range[3:] - coors_range[:3]) / voxel_size
    voxelmap_shape = tuple(np.round(voxelmap_shape).astype(np.int32).tolist())
    if reverse_index:
        voxelmap_shape = voxelmap_shape[::-1]
    num_points_per_voxel = np.zeros(shape=(max_voxels,), dtype=np.int32)
    coor_to_voxelidx = -np.ones(shape=voxelmap_shape, dtype=np.int32)
    voxels = np.zeros(shape=(max_voxels, max_points, points.shape[-1]), dtype=points.dtype)
    coors = np.zeros(shape=(max_voxels, 3), dtype=np.int32)

    N = points.shape[0]
    ndim = 3
    grid_size = (coors_range[3:] - coors_range[:3]) / voxel_size
    grid_size = np.round(grid_size).astype(np.int32)
    coor = np.zeros(shape=(3,), dtype=np.int32)
    voxel_num = 0
    for i in range(N):
        failed = False
        for j in range(ndim):
            c = np.floor((points[i, j] - coors_range[j]) / voxel_size[j])   # fp32 throughout
            if c < 0 or c >= grid_size[j]:
                failed = True
                break
            coor[(ndim - 1 - j) if reverse_index else j] = c
        if failed:
            continue
        voxelidx = coor_to_voxelidx[coor[0], coor[1], coor[2]]
        if voxelidx == -1:
            voxelidx = voxel_num
            if voxel_num >= max_voxels:
                break
            voxel_num += 1
            coor_to_voxelidx[coor[0], coor[1], coor[2]] = voxelidx
            coors[voxelidx] = coor
        num = num_points_per_voxel[voxelidx]
        if num < max_points:
            voxels[voxelidx, num] = points[i]
            num_points_per_voxel[voxelidx] += 1
    return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


def merge_coordinates(coors_list):
    """data/preprocess.py:30-38: prepend the sample index -> [P,4] (b,z,y,x)."""
    out = []
    for i, coor in enumerate(coors_list):
        out.append(np.pad(coor, ((0, 0), (1, 0)), mode="constant", constant_values=i))
    return np.concatenate(out, axis=0)


def get_paddings_indicator(actual_num, max_num, axis=0):
    """libs/tools/__init__.py:26-35."""
    actual_num = np.expand_dims(actual_num, axis + 1)
    max_num_shape = [1] * len(actual_num.shape)
    max_num_shape[axis + 1] = -1
    max_num = np.arange(max_num, dtype=np.int64).reshape(max_num_shape)
    return actual_num.astype(np.int64) > max_num


class PFNLayer:
    """pillars.py:9-41.  Linear(no bias) -> BatchNorm1D(eps 1e-3, momentum .01) -> ReLU -> max."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False, rng=None):
        rng = rng or np.random.default_rng(0)
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        self.use_norm = use_norm
        # paddle.nn.Linear weight is [in, out]
        self.weight = (rng.standard_normal((in_channels, out_channels)) / 3.0).astype(F32)
        self.bias = None if use_norm else np.zeros((out_channels,), dtype=F32)
        self.gamma = np.ones((out_channels,), dtype=F32)
        self.beta = np.zeros((out_channels,), dtype=F32)
        self._mean = np.zeros((out_channels,), dtype=F32)
        self._variance = np.ones((out_channels,), dtype=F32)
        self.eps = 1e-3
        self.momentum = 0.01
        self.training = True

    def forward(self, inputs, acc=np.float64):
        x = (inputs.astype(acc) @ self.weight.astype(acc))                 # :30
        if self.bias is not None:
            x = x + self.bias.astype(acc)
        x = x.astype(F32)
        if self.use_norm:                                                   # :31 BN over [P,C,T]
            xa = x.astype(acc)
            if self.training:
                mean = xa.mean(axis=(0, 1))
                var = xa.var(axis=(0, 1))
                self.last_mean, self.last_var = mean.astype(F32), var.astype(F32)
                self._mean = (self.momentum * self._mean + (1 - self.momentum) * mean).astype(F32)
                self._variance = (self.momentum * self._variance + (1 - self.momentum) * var).astype(F32)
            else:
                mean, var = self._mean.astype(acc), self._variance.astype(acc)
            x = ((xa - mean) / np.sqrt(var + acc(self.eps)) * self.gamma.astype(acc)
                 + self.beta.astype(acc)).astype(F32)
        x = np.maximum(x, F32(0))                                           # :32
        x_max = np.max(x, axis=1, keepdims=True)                            # :34
        if self.last_vfe:
            return x_max
        x_repeat = np.tile(x_max, (1, inputs.shape[1], 1))                  # :39
        return np.concatenate([x, x_repeat], axis=2)                        # :40

    __call__ = forward


class PillarFeatureNet:
    """pillars.py:43-108."""

    def __init__(self, num_input_features=4, use_norm=True, num_filters=(64, 128),
                 with_distance=False, voxel_size=(0.2, 0.2, 4),
                 pc_range=(0, -40, -3, 70.4, 40, 1), rng=None):
        assert len(num_filters) > 0
        num_input_features += 5
        if with_distance:
            num_input_features += 1
        self._with_distance = with_distance
        num_filters = [num_input_features] + list(num_filters)
        self.pfn_layers = []
        for i in range(len(num_filters) - 1):
            last_layer = not (i < len(num_filters) - 2)
            self.pfn_layers.append(PFNLayer(num_filters[i], num_filters[i + 1], use_norm,
                                            last_layer=last_layer, rng=rng))
        self.vx = voxel_size[0]
        self.vy = voxel_size[1]
        self.x_offset = self.vx / 2 + pc_range[0]
        self.y_offset = self.vy / 2 + pc_range[1]

    def decorate(self, features, num_voxels, coors):
        """pillars.py:81-102 -> [P,T,9(+1)] fp32, every op rounded to fp32 as Paddle does."""
        features = features.astype(F32)
        nv = num_voxels.astype(F32).reshape(-1, 1, 1)
        points_mean = features[:, :, :3].sum(axis=1, keepdims=True, dtype=F32) / nv     # :82
        f_cluster = features[:, :, :3] - points_mean                                     # :83
        f_center = np.zeros_like(features[:, :, :2])                                     # :86
        f_center[:, :, 0] = features[:, :, 0] - (
            coors[:, 3].astype(F32)[:, None] * F32(self.vx) + F32(self.x_offset))        # :87
        f_center[:, :, 1] = features[:, :, 1] - (
            coors[:, 2].astype(F32)[:, None] * F32(self.vy) + F32(self.y_offset))        # :88
        features_ls = [features, f_cluster, f_center]                                    # :91
        if self._with_distance:
            points_dist = np.sqrt((features[:, :, :3] ** 2).sum(axis=2, keepdims=True, dtype=F32))
            features_ls.append(points_dist)
        features = np.concatenate(features_ls, axis=-1)                                  # :95
        voxel_count = features.shape[1]
        mask = get_paddings_indicator(num_voxels, voxel_count, axis=0)                   # :100
        mask = np.expand_dims(mask, -1).astype(features.dtype)                           # :101
        features = features * mask                                                        # :102
        return features

    def forward(self, features, num_voxels, coors):
        features = self.decorate(features, num_voxels, coors)
        for pfn in self.pfn_layers:
            features = pfn(features)                                                     # :105-106
        return features.squeeze()                                                        # :108

    __call__ = forward


class PointPillarsScatter:
    """pillars.py:110-142 (+ mask_select / select_change, libs/functional.py:21-38)."""

    def __init__(self, output_shape, num_input_features=4):
        self.output_shape = output_shape
        self.ny = output_shape[2]
        self.nx = output_shape[3]
        self.nchannels = num_input_features

    def forward(self, voxel_features, coords, batch_size):
        batch_canvas = []
        for batch_itt in range(batch_size):                                              # :123
            canvas = np.zeros((self.nchannels, self.nx * self.ny), dtype=voxel_features.dtype)
            batch_mask = coords[:, 0] == batch_itt                                       # :126
            if batch_mask.any():
                this_coords = coords[batch_mask]                                         # :128
                indices = this_coords[:, 2] * self.nx + this_coords[:, 3]                # :129
                indices = indices.astype("int64")
                voxels = voxel_features[batch_mask].T                                    # :131-132
                canvas[:, indices] = voxels                                              # :134
            batch_canvas.append(canvas)
        batch_canvas = np.stack(batch_canvas, 0)                                         # :139
        return batch_canvas.reshape(batch_size, self.nchannels, self.ny, self.nx)        # :140

    __call__ = forward


# ------------------------------------------------------------------ N4: either side of the encode
def merge_second_batch_voxels(points_list, voxel_size, coors_range, max_points=35, reverse_index=True,
                              max_voxels=20000):
    """points_to_voxel per frame (what the dataset's prep function does) followed by the 'voxels' /
    'num_points' / 'coordinates' branches of merge_second_batch (data/preprocess.py:16-42; 'num_voxels' is popped
    there, :20 -- returned here beside the dict as the per-frame counts)."""
    vox, num, coors, nv = [], [], [], []
    for i, pts in enumerate(points_list):
        v, c, n = points_to_voxel(pts, voxel_size, coors_range, max_points, reverse_index, max_voxels)
        vox.append(v)
        num.append(n)
        coors.append(np.pad(c, ((0, 0), (1, 0)), mode="constant", constant_values=i))   # :33-36
        nv.append(v.shape[0])
    return {"voxels": np.concatenate(vox, axis=0), "num_points": np.concatenate(num, axis=0),
            "coordinates": np.concatenate(coors, axis=0)}, np.array(nv, np.int32)


def sparse_sum_for_anchors_mask(coors, shape):
    """box_np_ops.py:772-777."""
    ret = np.zeros(shape, dtype=np.float32)
    for i in range(coors.shape[0]):
        ret[coors[i, 1], coors[i, 2]] += 1
    return ret


def fused_get_anchors_area(dense_map, anchors_bv, stride, offset, grid_size):
    """box_np_ops.py:781-806 (dense_map is the 2-D inclusive prefix sum of the occupancy map); fp32 arithmetic,
    i.e. what the reference computes for float32 anchors / voxel_size / pc_range."""
    anchors_bv = np.asarray(anchors_bv, F32)
    stride = np.asarray(stride, F32)
    offset = np.asarray(offset, F32)
    gx, gy = int(grid_size[0]) - 1, int(grid_size[1]) - 1
    ret = np.zeros((anchors_bv.shape[0],), dtype=dense_map.dtype)
    for i in range(anchors_bv.shape[0]):
        c0 = int(np.floor((anchors_bv[i, 0] - offset[0]) / stride[0]))
        c1 = int(np.floor((anchors_bv[i, 1] - offset[1]) / stride[1]))
        c2 = int(np.floor((anchors_bv[i, 2] - offset[0]) / stride[0]))
        c3 = int(np.floor((anchors_bv[i, 3] - offset[1]) / stride[1]))
        c0, c1, c2, c3 = max(c0, 0), max(c1, 0), min(c2, gx), min(c3, gy)
        ret[i] = dense_map[c3, c2] - dense_map[c3, c0] - dense_map[c1, c2] + dense_map[c1, c0]
    return ret


def points_to_bev(points, voxel_size, coors_range, with_reflectivity=False, density_norm_num=16, max_voxels=40000):
    """bev_ops.py:6-103, the sequential loop restated (fp32 throughout, like the jitted kernel on float32 input)."""
    points = np.asarray(points, F32)
    voxel_size = np.asarray(voxel_size, dtype=F32)
    coors_range = np.asarray(coors_range, dtype=F32)
    shape = tuple(np.round((coors_range[3:] - coors_range[:3]) / voxel_size).astype(np.int32).tolist())[::-1]
    grid = shape[::-1]
    seen = -np.ones(shape, np.int32)
    D = shape[0]
    lowers = np.linspace(coors_range[2], coors_range[5], D, endpoint=False).astype(F32)
    bev = np.zeros((D + 1 + (1 if with_reflectivity else 0),) + shape[1:], F32)
    slice_h = voxel_size[2]
    voxel_num = 0
    for i in range(points.shape[0]):
        coor = [0, 0, 0]
        failed = False
        for j in range(3):
            c = np.floor((points[i, j] - coors_range[j]) / voxel_size[j])
            if c < 0 or c >= grid[j]:
                failed = True
                break
            coor[2 - j] = int(c)
        if failed:
            continue
        if seen[coor[0], coor[1], coor[2]] == -1:
            if voxel_num >= max_voxels:
                break
            seen[coor[0], coor[1], coor[2]] = voxel_num
            voxel_num += 1
        bev[-1, coor[1], coor[2]] += 1
        h = F32(F32(points[i, 2] - lowers[coor[0]]) / slice_h)
        if h > bev[coor[0], coor[1], coor[2]]:
            bev[coor[0], coor[1], coor[2]] = h
            if with_reflectivity:
                bev[-2, coor[1], coor[2]] = points[i, 3]
    return bev
