"""Seeded synthetic inputs for the hot path -- host side, NumPy only.

Shapes follow the reference's data contract: clouds are ``[B,3,N] float32`` exactly as
``pnloader`` yields them (PAPC/datasets/pnloader.py:43,96 ``datas[i].T.astype('float32')``)
and normalised like ``pc_normalize`` (pointnet2_basic_layers.py:17-23); lidar frames are
``[N,4] float32`` (x,y,z,intensity) like KITTI ``.bin`` files (pp/data/preprocess.py:306-363).
"""
from __future__ import annotations

import numpy as np

# yaml geometry of BASELINE config 4 (pp/params/configs/pointpillars_kitti_car_xy16.yaml:7-11)
KITTI_VOXEL_SIZE = (0.16, 0.16, 4.0)
KITTI_PC_RANGE = (0.0, -39.68, -3.0, 69.12, 39.68, 1.0)
KITTI_MAX_POINTS = 100
KITTI_MAX_VOXELS = 12000


def clouds(B, N, seed=0):
    """``xyz ~ U(-1,1)^3`` then per-cloud centroid removal and max-norm scaling -> [B,3,N] fp32."""
    rng = np.random.default_rng(seed)
    pc = rng.uniform(-1.0, 1.0, (B, N, 3)).astype(np.float32)
    pc = pc - pc.mean(axis=1, keepdims=True)
    m = np.sqrt((pc ** 2).sum(-1)).max(axis=1)
    pc = pc / np.where(m > 0, m, 1.0).astype(np.float32)[:, None, None]   # N = 1: the single point is the centroid
    return np.ascontiguousarray(pc.transpose(0, 2, 1)).astype(np.float32)


def fps_start(B, N, seed=1):
    """Seeded stand-in for ``paddle.randint(0, N, (B,))`` (layers.py:76)."""
    return np.random.default_rng(seed).integers(0, N, B).astype(np.int64)


def mlp_params(cin, mlp, seed=2):
    """Conv2D 1x1 weights N(0, sqrt(2/fan_in)), bias U(-.1,.1), gamma=1, beta=0 per layer."""
    rng = np.random.default_rng(seed)
    out = []
    last = cin
    for c in mlp:
        w = (rng.standard_normal((c, last)) * np.sqrt(2.0 / last)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, c).astype(np.float32)
        out.append(dict(weight=w, bias=b, gamma=np.ones(c, np.float32), beta=np.zeros(c, np.float32)))
        last = c
    return out


def lidar_frame(N=20000, seed=0, shuffle=False):
    """KITTI-shaped frame, the exact draw order of SURVEY 8c K5 -> [N,4] fp32."""
    rng = np.random.default_rng(seed)
    th = rng.uniform(-np.pi / 4, np.pi / 4, N)
    rho = np.exp(rng.uniform(np.log(2.0), np.log(70.0), N))
    z = rng.uniform(-2.5, 0.5, N)
    i = rng.uniform(0.0, 1.0, N)
    pts = np.stack([rho * np.cos(th), rho * np.sin(th), z, i], 1).astype(np.float32)
    if shuffle:
        pts = pts[np.random.default_rng(seed + 1000).permutation(N)]
    return np.ascontiguousarray(pts)


def lidar_uniform(N=20000, seed=0, pc_range=KITTI_PC_RANGE, margin=1.0):
    """Uniform-in-range stress cloud (hits the max_voxels cap early); some points outside."""
    rng = np.random.default_rng(seed)
    lo = np.array(pc_range[:3]) - margin
    hi = np.array(pc_range[3:]) + margin
    xyz = rng.uniform(lo, hi, (N, 3))
    i = rng.uniform(0, 1, (N, 1))
    return np.ascontiguousarray(np.concatenate([xyz, i], 1).astype(np.float32))


def pfn_weight(cin=9, cout=64, seed=3):
    """PFN Linear weight [cin,cout] ~ N(0, 1/3) (paddle.nn.Linear stores [in,out])."""
    return (np.random.default_rng(seed).standard_normal((cin, cout)) / 3.0).astype(np.float32)
