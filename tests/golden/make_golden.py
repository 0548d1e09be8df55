"""Generate tests/golden/voxel_*.npz by RUNNING THE REFERENCE'S OWN numba voxeliser.

Build-container only: imports /root/reference/.../point_cloud_ops.py by file path
(oracle/ref_voxel.py).  Run from the repo root:  python tests/golden/make_golden.py
The committed .npz files are what the oracle (and through it the CUDA path) is pinned to;
/root/reference does not exist on the GPU box.

Cases
  voxel_small_*   full output arrays (voxels, coors, num_points) on small grids, covering
                  reverse_index True/False, F=3/4/5, the max_voxels ``break``, max_points
                  overflow, points outside the range and exactly on cell / range edges.
  voxel_k5*       yaml geometry, N=20000 (unshuffled / shuffled / uniform
                  stress): coors + num_points in full, voxels as sha256 + per-voxel sums
                  (the 19 MB array itself is too large to commit).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_voxel  # noqa: E402
from papc_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def small_case(name, N, F, vs, rng_, max_points, max_voxels, reverse, seed, edge=False):
    rng = np.random.default_rng(seed)
    lo, hi = np.array(rng_[:3]), np.array(rng_[3:])
    pts = rng.uniform(lo - 0.5, hi + 0.5, (N, 3))
    if edge:
        # snap a third of the coordinates onto exact multiples of the voxel size / range edges
        snap = rng.random((N, 3)) < 0.33
        grid = np.round((pts - lo) / np.array(vs)) * np.array(vs) + lo
        pts = np.where(snap, grid, pts)
        pts[0, :] = lo
        pts[1, :] = hi
        pts[2, :] = np.nextafter(np.float32(hi), np.float32(-np.inf))
    extra = rng.uniform(0, 1, (N, F - 3))
    points = np.concatenate([pts, extra], 1).astype(np.float32)
    vs32 = np.array(vs, np.float32)
    cr32 = np.array(rng_, np.float32)
    v, c, n = ref_voxel.points_to_voxel(points, vs32, cr32, max_points, reverse, max_voxels)
    np.savez_compressed(os.path.join(OUT, f"voxel_small_{name}.npz"), points=points,
                        voxel_size=vs32, coors_range=cr32, max_points=max_points,
                        max_voxels=max_voxels, reverse_index=reverse, voxels=v, coors=c,
                        num_points=n)
    print(name, "voxels", v.shape, "sum(num)", int(n.sum()), "max(num)", int(n.max()) if len(n) else 0)


def k5_case(name, points):
    vs = np.array(synth.KITTI_VOXEL_SIZE, np.float32)
    cr = np.array(synth.KITTI_PC_RANGE, np.float32)
    v, c, n = ref_voxel.points_to_voxel(points, vs, cr, synth.KITTI_MAX_POINTS, True,
                                        synth.KITTI_MAX_VOXELS)
    sha = hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest()
    vsum = v.astype(np.float64).sum(axis=1)
    np.savez_compressed(os.path.join(OUT, f"voxel_{name}.npz"), coors=c, num_points=n,
                        voxels_sha256=sha, voxel_sums=vsum, n_points=points.shape[0],
                        points_sha256=hashlib.sha256(points.tobytes()).hexdigest())
    print(name, "voxels", v.shape, "sum(num)", int(n.sum()), "max(num)", int(n.max()), sha[:16])


if __name__ == "__main__":
    assert ref_voxel.available(), "needs /root/reference and numba"
    small_case("a", 3000, 4, (0.5, 0.5, 2.0), (0, -4, -1, 8, 4, 1), 5, 100000, True, 10)
    small_case("b_break", 3000, 4, (0.5, 0.5, 2.0), (0, -4, -1, 8, 4, 1), 5, 150, True, 11)
    small_case("c_norev", 2000, 3, (0.4, 0.25, 0.5), (-2, -2, -1, 2, 2, 1), 3, 500, False, 12)
    small_case("d_edge", 4000, 5, (0.25, 0.25, 0.25), (-1, -1, -1, 1, 1, 1), 4, 300, True, 13, edge=True)
    small_case("e_one", 500, 4, (8.0, 8.0, 2.0), (0, -4, -1, 8, 4, 1), 7, 10, True, 14)
    k5_case("k5", synth.lidar_frame(20000, 0, False))
    k5_case("k5_shuffled", synth.lidar_frame(20000, 0, True))
    k5_case("k5_uniform", synth.lidar_uniform(20000, 0))
