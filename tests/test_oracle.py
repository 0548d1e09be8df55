"""CPU tests pinning the oracle (no GPU).

* the voxeliser restatements (C and pure Python) against the golden vectors produced by the
  reference's own numba kernel (tests/golden/make_golden.py), and -- when /root/reference is
  present -- against the reference run live;
* FPS / ball query against the hand-derivable known-answer vectors K1-K4;
* the arithmetic-pinned C restatement against the line-by-line NumPy restatement.
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import capi, layers_np, pillars_np, ref_voxel
from papc_b200 import synth


# ------------------------------------------------------------------ voxeliser (pinned)
def _small_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "voxel_small_*.npz")))


def test_golden_files_present(golden_dir):
    assert len(_small_cases(golden_dir)) >= 5
    for n in ("k5", "k5_shuffled", "k5_uniform"):
        assert os.path.exists(os.path.join(golden_dir, f"voxel_{n}.npz"))


@pytest.mark.parametrize("impl", ["c", "python"])
def test_voxelize_small_golden(golden_dir, impl):
    fn = capi.points_to_voxel if impl == "c" else pillars_np.points_to_voxel
    for path in _small_cases(golden_dir):
        g = np.load(path)
        v, c, n = fn(g["points"], g["voxel_size"], g["coors_range"], int(g["max_points"]),
                     bool(g["reverse_index"]), int(g["max_voxels"]))
        assert v.shape == g["voxels"].shape, path
        np.testing.assert_array_equal(c, g["coors"], err_msg=path)
        np.testing.assert_array_equal(n, g["num_points"], err_msg=path)
        np.testing.assert_array_equal(v, g["voxels"], err_msg=path)


@pytest.mark.parametrize("name,points", [
    ("k5", lambda: synth.lidar_frame(20000, 0, False)),
    ("k5_shuffled", lambda: synth.lidar_frame(20000, 0, True)),
    ("k5_uniform", lambda: synth.lidar_uniform(20000, 0)),
])
def test_voxelize_k5_golden(golden_dir, name, points):
    g = np.load(os.path.join(golden_dir, f"voxel_{name}.npz"))
    pts = points()
    assert hashlib.sha256(pts.tobytes()).hexdigest() == str(g["points_sha256"]), \
        "synthetic generator drifted from the one the golden vectors were made with"
    v, c, n = capi.points_to_voxel(pts, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE,
                                   synth.KITTI_MAX_POINTS, True, synth.KITTI_MAX_VOXELS)
    np.testing.assert_array_equal(c, g["coors"])
    np.testing.assert_array_equal(n, g["num_points"])
    assert hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() == str(g["voxels_sha256"])
    if name == "k5":  # the three scalars quoted in SURVEY.md 8(c)
        assert v.shape == (12000, 100, 4) and int(n.sum()) == 19508 and int(n.max()) == 25


@pytest.mark.skipif(not ref_voxel.available(), reason="/root/reference (build container only)")
def test_voxelize_against_live_reference():
    rng = np.random.default_rng(77)
    for trial in range(6):
        N = int(rng.integers(1, 4000))
        F = int(rng.integers(3, 6))
        pts = rng.uniform(-3, 3, (N, F)).astype(np.float32)
        vs = np.array([0.3, 0.2, 0.7], np.float32)
        cr = np.array([-2, -2.2, -2.1, 2.2, 2, 2.1], np.float32)
        mp, mv, rev = int(rng.integers(1, 6)), int(rng.integers(1, 900)), bool(trial % 2)
        rv, rc, rn = ref_voxel.points_to_voxel(pts, vs, cr, mp, rev, mv)
        ov, oc, on = capi.points_to_voxel(pts, vs, cr, mp, rev, mv)
        np.testing.assert_array_equal(oc, rc)
        np.testing.assert_array_equal(on, rn)
        np.testing.assert_array_equal(ov, rv)


def test_voxelize_empty_and_all_outside():
    vs, cr = synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE
    v, c, n = capi.points_to_voxel(np.zeros((0, 4), np.float32), vs, cr, 5, True, 10)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and n.shape == (0,)
    pts = np.full((7, 4), 1000.0, np.float32)
    v, c, n = capi.points_to_voxel(pts, vs, cr, 5, True, 10)
    assert v.shape[0] == 0


# ------------------------------------------------------------------ FPS / ball query KATs
def _line(xs):
    return np.array([[[x, 0.0, 0.0] for x in xs]], np.float32)


@pytest.mark.parametrize("impl", ["c", "numpy"])
def test_fps_k1_k2(impl):
    def fps(xyz, npoint, start):
        if impl == "c":
            return capi.farthest_point_sample(xyz, npoint, start)
        return layers_np.farthest_point_sample(xyz, npoint, start_idx=start).astype(np.int64)
    # K1
    np.testing.assert_array_equal(fps(_line([0, .1, .5, .9]), 3, [0]), [[0, 3, 2]])
    # K2: distance is initialised to 1.0, so far points tie at the clamp -> first index
    np.testing.assert_array_equal(fps(_line([0, 2, 3]), 3, [0]), [[0, 1, 2]])


def test_fps_returns_float32_like_reference():
    out = layers_np.farthest_point_sample(_line([0, .1, .5, .9]), 3, start_idx=[0])
    assert out.dtype == np.float32  # layers.py:74 paddle.zeros default dtype


@pytest.mark.parametrize("impl", ["c", "numpy"])
def test_ball_query_k3_k4(impl):
    def bq(r, k, xyz, new_xyz):
        if impl == "c":
            return capi.query_ball_point(r, k, xyz, new_xyz)[0]
        return layers_np.query_ball_point(r, k, xyz, new_xyz)
    xyz = _line([0, .1, .15, .3, .19, .7])
    q = xyz[:, :1, :]
    np.testing.assert_array_equal(bq(0.2, 3, xyz, q), [[[0, 1, 2]]])            # K3
    np.testing.assert_array_equal(bq(0.2, 6, xyz, q), [[[0, 1, 2, 4, 0, 0]]])   # K3
    xyz4 = _line([0, .1, .15, .3, .2, .7])                                      # K4
    # d^2 = fp32(0.2)^2 = 0.040000003 > float32(0.2**2) = 0.04 -> excluded
    assert np.float32(0.2) * np.float32(0.2) > np.float32(0.2 ** 2)
    np.testing.assert_array_equal(bq(0.2, 6, xyz4, q), [[[0, 1, 2, 0, 0, 0]]])


def test_c_vs_numpy_restatement_sa_sized():
    """C (FMA-chain, pinned) vs NumPy (BLAS matmul, as the reference writes it) on SA-shaped data."""
    xyz = synth.clouds(4, 1024, seed=5).transpose(0, 2, 1).copy()
    start = synth.fps_start(4, 1024, seed=6)
    f_c = capi.farthest_point_sample(xyz, 128, start)
    f_n = layers_np.farthest_point_sample(xyz, 128, start_idx=start)
    np.testing.assert_array_equal(f_c, f_n.astype(np.int64))
    assert all(len(set(r)) == 128 for r in f_c.tolist())  # unique while distances > 0
    new_xyz = layers_np.index_points(xyz, f_c)
    for r, k in ((0.2, 32), (0.4, 64), (0.1, 16), (0.8, 128)):
        i_c, empty = capi.query_ball_point(r, k, xyz, new_xyz)
        i_n = layers_np.query_ball_point(r, k, xyz, new_xyz)
        assert empty == 0
        # BLAS's K=3 dot is not specified bit-for-bit; SURVEY found it identical to the FMA
        # chain.  Tolerate a vanishing fraction of threshold flips but report them.
        mism = int((i_c != i_n).any(axis=-1).sum())
        assert mism <= 1, f"r={r}: {mism} groups differ between C and NumPy restatements"
        assert (i_c >= 0).all() and (i_c < 1024).all()
        first = i_c[..., :1]
        asc = (np.diff(i_c, axis=-1) > 0) | (i_c[..., 1:] == first)
        assert asc.all()


def test_square_distance_c_vs_numpy():
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, (2, 50, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (2, 70, 3)).astype(np.float32)
    np.testing.assert_allclose(capi.square_distance(a, b), layers_np.square_distance(a, b),
                               rtol=0, atol=1e-6)


def test_ball_query_empty_ball_is_flagged():
    xyz = _line([0.0, 0.1])
    q = np.array([[[5.0, 0, 0]]], np.float32)
    idx, empty = capi.query_ball_point(0.2, 4, xyz, q)
    assert empty == 1 and (idx == 2).all()  # N, as the reference's sort would leave it


# ------------------------------------------------------------------ SA layer shapes / quirks
def test_sa_layer_shapes_and_concat_order():
    B, N = 2, 256
    xyz = synth.clouds(B, N, seed=1)
    feats = np.random.default_rng(2).standard_normal((B, 5, N)).astype(np.float32)
    start = synth.fps_start(B, N)
    sa = layers_np.PointNetSetAbstraction(64, 0.3, 16, 3 + 5, [16, 32], False)
    nx, npts = sa(xyz, feats, start_idx=start)
    assert nx.shape == (B, 3, 64) and npts.shape == (B, 32, 64)
    xt, ft = xyz.transpose(0, 2, 1), feats.transpose(0, 2, 1)
    _, grouped = layers_np.sample_and_group(64, 0.3, 16, xt, ft, start_idx=start)
    assert grouped.shape == (B, 64, 16, 8)
    assert np.abs(grouped[..., :3]).max() <= 0.3 + 1e-6       # xyz_norm first (:151)
    msg = layers_np.PointNetSetAbstractionMsg(64, [0.2, 0.4], [8, 16], 5, [[16, 16], [16, 24]])
    nx2, np2 = msg(xyz, feats, start_idx=start)
    assert nx2.shape == (B, 3, 64) and np2.shape == (B, 40, 64)
    np.testing.assert_array_equal(nx, nx2)
    ga = layers_np.PointNetSetAbstraction(None, None, None, 8, [16], True)
    gx, gp = ga(xyz, feats)
    assert gx.shape == (B, 3, 1) and (gx == 0).all() and gp.shape == (B, 16, 1)


def test_pfn_and_scatter_oracle_shapes():
    pts = synth.lidar_frame(3000, 1)
    v, c, n = capi.points_to_voxel(pts, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, 100, True, 12000)
    coors = pillars_np.merge_coordinates([c])
    pfn = pillars_np.PillarFeatureNet(4, True, (64,), False, (0.16, 0.16, 4), synth.KITTI_PC_RANGE)
    dec = pfn.decorate(v, n, coors)
    assert dec.shape == (v.shape[0], 100, 9)
    assert (dec[np.arange(100)[None, :] >= n[:, None]] == 0).all()
    out = pfn(v, n, coors)
    assert out.shape == (v.shape[0], 64) and (out >= 0).all()
    sc = pillars_np.PointPillarsScatter([1, 1, 496, 432], 64)
    canvas = sc(out, coors, 1)
    assert canvas.shape == (1, 64, 496, 432)
    p = 17
    np.testing.assert_array_equal(canvas[0, :, c[p, 1], c[p, 2]], out[p])
    assert int((np.abs(canvas).sum(1) > 0).sum()) <= v.shape[0]


def test_feature_propagation_restatement_quirk_and_shapes():
    """layers.py:316-324: idx is the argsort of the already sorted distances (identity), so the
    interpolation mixes the features of sampled points 0, 1, 2 with the 3 smallest-distance weights."""
    from oracle import layers_np as o
    rng = np.random.default_rng(5)
    B, N, S, D1, D2 = 2, 32, 8, 3, 5
    xyz1 = rng.standard_normal((B, 3, N)).astype(np.float32)
    xyz2 = np.ascontiguousarray(xyz1[:, :, :S])
    p1 = rng.standard_normal((B, D1, N)).astype(np.float32)
    p2 = rng.standard_normal((B, D2, S)).astype(np.float32)
    fp = o.PointNetFeaturePropagation(D1 + D2, [4, 6])
    z = fp.interpolate(xyz1, xyz2, p1, p2)
    assert z.shape == (B, N, D1 + D2)
    np.testing.assert_array_equal(z[:, :, :D1], p1.transpose(0, 2, 1))
    # hand computation for one point
    d = np.sort(o.square_distance(xyz1.transpose(0, 2, 1), xyz2.transpose(0, 2, 1)), -1)[0, 7, :3]
    w = (np.float32(1) / (d + np.float32(1e-8)))
    w = w / w.sum()
    want = (p2[0, :, :3] * w[None, :]).sum(1)
    np.testing.assert_allclose(z[0, 7, D1:], want, rtol=1e-6, atol=1e-6)
    # S == 1 tiles
    z1 = fp.interpolate(xyz1, xyz2[:, :, :1], None, p2[:, :, :1])
    np.testing.assert_array_equal(z1, np.tile(p2[:, :, :1].transpose(0, 2, 1), [1, N, 1]))
    assert fp(xyz1, xyz2, p1, p2).shape == (B, 6, N)


def test_pc_normalize_matches_reference_golden(golden_dir):
    """layers.py:17-23: outputs of the reference's own function (tests/golden/make_golden_pc_normalize.py);
    the synthetic cloud generator applies the same normalisation per cloud."""
    from papc_b200 import synth
    g = np.load(os.path.join(golden_dir, "pc_normalize.npz"))
    for x, y in ((g["x32"], g["y32"]), (g["x64"], g["y64"])):
        out = layers_np.pc_normalize(x)
        assert out.dtype == y.dtype
        np.testing.assert_array_equal(out, y)
    # synth.clouds(B, N, seed) == pc_normalize of the same raw draws, cloud by cloud
    B, N = 3, 1024
    raw = np.random.default_rng(0).uniform(-1.0, 1.0, (B, N, 3)).astype(np.float32)
    np.testing.assert_array_equal(raw[0], g["x32"])          # the golden input IS cloud 0 of seed 0
    got = synth.clouds(B, N, seed=0).transpose(0, 2, 1)
    np.testing.assert_allclose(got[0], g["y32"], rtol=0, atol=2e-7)
    for b in range(B):
        np.testing.assert_allclose(got[b], layers_np.pc_normalize(raw[b]), rtol=0, atol=2e-7)
