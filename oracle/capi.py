"""ctypes bindings of oracle/papc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpapc_oracle.so")
        if not os.path.exists(path) or any(os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, f))
                                           for f in ("papc_oracle.c", "nms_oracle.c")):
            from . import build as _b
            _b.build()
        _LIB = C.CDLL(path)
        _LIB.oracle_ball_query_f32.restype = C.c_int64
        _LIB.oracle_voxelize_f32.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def square_distance(src, dst):
    src, dst = _f32(src), _f32(dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().oracle_square_distance_f32(_p(src), _p(dst), C.c_int(B), C.c_int(N), C.c_int(M), _p(out))
    return out


def farthest_point_sample(xyz, npoint, start_idx, init_dist=1.0):
    """xyz [B,N,3] -> int64 [B,npoint]  (layers.py:65-95 with an explicit start index)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    start = np.ascontiguousarray(start_idx, dtype=np.int64)
    out = np.empty((B, npoint), np.int64)
    lib().oracle_fps_f32(_p(xyz), C.c_int(B), C.c_int(N), C.c_int(npoint), _p(start),
                         C.c_float(init_dist), _p(out))
    return out


def radius2_f32(radius):
    """layers.py:112: ``radius ** 2`` is a Python double, compared against fp32 tensor."""
    return np.float32(float(radius) ** 2)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """-> (int64 [B,S,nsample], number of empty balls)  (layers.py:98-126)."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), np.int64)
    empty = lib().oracle_ball_query_f32(_p(xyz), _p(new_xyz), C.c_int(B), C.c_int(N), C.c_int(S),
                                        C.c_float(radius2_f32(radius)), C.c_int(nsample), _p(out))
    return out, int(empty)


def three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    d = np.empty((B, N, 3), np.float32)
    i = np.empty((B, N, 3), np.int64)
    lib().oracle_three_nn_f32(_p(xyz1), _p(xyz2), C.c_int(B), C.c_int(N), C.c_int(S), _p(d), _p(i))
    return d, i


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True,
                    max_voxels=20000):
    """point_cloud_ops.py:106-166 via the C restatement."""
    points = _f32(points)
    vs = _f32(np.asarray(voxel_size))
    cr = _f32(np.asarray(coors_range))
    N, F = points.shape
    voxels = np.zeros((max_voxels, max_points, F), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    n = lib().oracle_voxelize_f32(_p(points), C.c_int(N), C.c_int(F), _p(vs), _p(cr),
                                  C.c_int(max_points), C.c_int(1 if reverse_index else 0),
                                  C.c_int(max_voxels), _p(voxels), _p(coors), _p(num))
    return voxels[:n], coors[:n], num[:n]


def nms(dets, thresh, rotated=False):
    """nms_gpu (nms_gpu.py:133-164) / rotate_nms_gpu (:453-488): kept ORIGINAL indices in kept order."""
    dets = _f32(dets)
    n = dets.shape[0]
    assert dets.shape[1] == (6 if rotated else 5)
    keep = np.zeros(max(n, 1), np.int32)
    lib().oracle_nms_f32.restype = C.c_int
    num = lib().oracle_nms_f32(_p(dets), C.c_int(n), C.c_float(thresh), C.c_int(1 if rotated else 0), _p(keep))
    return keep[:num].copy()


def rotate_iou(boxes, query_boxes, criterion=-1):
    """rotate_iou_gpu_eval (nms_gpu.py:603-653); criterion -1 is rotate_iou_gpu (:518-553)."""
    boxes, query_boxes = _f32(boxes), _f32(query_boxes)
    N, K = boxes.shape[0], query_boxes.shape[0]
    out = np.zeros((N, K), np.float32)
    if N and K:
        lib().oracle_rotate_iou_f32(_p(boxes), C.c_int(N), _p(query_boxes), C.c_int(K), C.c_int(criterion), _p(out))
    return out
