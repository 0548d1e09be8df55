"""GPU parity tests of the pillar path: voxeliser bit-exact vs the golden vectors produced by the
reference's own numba kernel (and vs the oracle on random cases); PFN / scatter vs the oracle."""
import glob
import hashlib
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import capi, pillars_np  # noqa: E402
from papc_b200 import pillars, synth  # noqa: E402

DEV = "cuda:0"
TOL = dict(rtol=1e-5, atol=1e-5)


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_voxelize_small_golden(golden_dir):
    paths = sorted(glob.glob(os.path.join(golden_dir, "voxel_small_*.npz")))
    assert len(paths) >= 5
    for path in paths:
        g = np.load(path)
        v, c, n = pillars.points_to_voxel(g["points"], g["voxel_size"], g["coors_range"],
                                          int(g["max_points"]), bool(g["reverse_index"]), int(g["max_voxels"]))
        assert v.shape == g["voxels"].shape and c.dtype == np.int32 and n.dtype == np.int32, path
        np.testing.assert_array_equal(c, g["coors"], err_msg=path)
        np.testing.assert_array_equal(n, g["num_points"], err_msg=path)
        np.testing.assert_array_equal(v, g["voxels"], err_msg=path)


@pytest.mark.parametrize("name,points", [
    ("k5", lambda: synth.lidar_frame(20000, 0, False)),
    ("k5_shuffled", lambda: synth.lidar_frame(20000, 0, True)),
    ("k5_uniform", lambda: synth.lidar_uniform(20000, 0)),
])
def test_voxelize_k5_golden(golden_dir, name, points):
    """BASELINE config 4 at full size against the reference's own output."""
    g = np.load(os.path.join(golden_dir, f"voxel_{name}.npz"))
    v, c, n = pillars.points_to_voxel(points(), synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE,
                                      synth.KITTI_MAX_POINTS, True, synth.KITTI_MAX_VOXELS)
    np.testing.assert_array_equal(c, g["coors"])
    np.testing.assert_array_equal(n, g["num_points"])
    assert hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() == str(g["voxels_sha256"])


@pytest.mark.parametrize("seed", range(6))
def test_voxelize_random_vs_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    N = int(rng.integers(1, 30000))
    F = int(rng.integers(3, 7))
    pts = rng.uniform(-3, 3, (N, F)).astype(np.float32)
    if seed % 2:
        pts[:, :3] = np.round(pts[:, :3] * 4) / 4  # many points exactly on cell edges, dense cells
    vs = np.array([0.3, 0.25, 0.7], np.float32)
    cr = np.array([-2, -2.25, -2.1, 2.2, 2, 2.1], np.float32)
    mp, mv, rev = int(rng.integers(1, 40)), int(rng.integers(1, 3000)), bool(seed % 3)
    ov, oc, on = capi.points_to_voxel(pts, vs, cr, mp, rev, mv)
    gv, gc, gn = pillars.points_to_voxel(pts, vs, cr, mp, rev, mv)
    np.testing.assert_array_equal(gc, oc)
    np.testing.assert_array_equal(gn, on)
    np.testing.assert_array_equal(gv, ov)


def test_voxelize_edge_cases():
    vs, cr = synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE
    v, c, n = pillars.points_to_voxel(np.zeros((0, 4), np.float32), vs, cr, 5, True, 10)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and n.shape == (0,)
    v, c, n = pillars.points_to_voxel(np.full((7, 4), 1000.0, np.float32), vs, cr, 5, True, 10)
    assert v.shape[0] == 0
    # everything in ONE cell, far more points than max_points: first max_points in input order
    pts = np.tile(np.array([[1.0, 1.0, 0.0, 0.0]], np.float32), (5000, 1))
    pts[:, 3] = np.arange(5000)
    v, c, n = pillars.points_to_voxel(pts, vs, cr, 100, True, 10)
    assert v.shape[0] == 1 and n.tolist() == [100]
    np.testing.assert_array_equal(v[0, :, 3], np.arange(100, dtype=np.float32))
    # device-resident form: padded outputs, rows >= voxel_num are zero
    dv, dc, dn, dnum = pillars.points_to_voxel_device(_cu(synth.lidar_frame(3000, 2)), vs, cr, 20, True, 4000)
    m = int(dnum.item())
    assert 0 < m < 4000
    assert dv[m:].abs().sum().item() == 0 and dc[m:].abs().sum().item() == 0 and dn[m:].sum().item() == 0


def _pfn_pair(seed, use_norm=True, F=4, cout=64, vsz=(0.16, 0.16, 4), pcr=synth.KITTI_PC_RANGE):
    rng = np.random.default_rng(seed)
    gpu = pillars.PillarFeatureNet(F, use_norm, (cout,), False, vsz, pcr)
    ref = pillars_np.PillarFeatureNet(F, use_norm, (cout,), False, vsz, pcr)
    w = synth.pfn_weight(F + 5, cout, seed)
    gamma = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    gamma[::5] *= -1
    beta = rng.uniform(-0.2, 0.2, cout).astype(np.float32)
    g, r = gpu.pfn_layers[0], ref.pfn_layers[0]
    g.weight.data = torch.from_numpy(w)
    g.bn_weight.data, g.bn_bias.data = torch.from_numpy(gamma), torch.from_numpy(beta)
    r.weight, r.gamma, r.beta = w, gamma, beta
    if not use_norm:
        b = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
        g.bias.data = torch.from_numpy(b)
        r.bias = b
    return gpu.to(DEV), ref


@pytest.mark.parametrize("mode", ["train", "eval", "nonorm"])
def test_pfn_vs_oracle(mode):
    pts = np.concatenate([synth.lidar_frame(6000, 1), synth.lidar_frame(5000, 2)])
    v0, c0, n0 = capi.points_to_voxel(pts[:6000], synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, 100, True, 12000)
    v1, c1, n1 = capi.points_to_voxel(pts[6000:], synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, 100, True, 12000)
    voxels = np.concatenate([v0, v1]); num = np.concatenate([n0, n1])
    coors = pillars_np.merge_coordinates([c0, c1])
    gpu, ref = _pfn_pair(3, use_norm=(mode != "nonorm"))
    if mode == "eval":
        rng = np.random.default_rng(8)
        m = rng.uniform(-1, 1, 64).astype(np.float32); var = rng.uniform(20, 60, 64).astype(np.float32)
        gpu.pfn_layers[0]._mean.copy_(torch.from_numpy(m)); gpu.pfn_layers[0]._variance.copy_(torch.from_numpy(var))
        ref.pfn_layers[0]._mean, ref.pfn_layers[0]._variance = m, var
        gpu.eval(); ref.pfn_layers[0].training = False
    out = gpu(_cu(voxels), _cu(num), _cu(coors))
    exp = ref(voxels, num, coors)
    assert tuple(out.shape) == exp.shape == (voxels.shape[0], 64)
    np.testing.assert_allclose(out.cpu().numpy(), exp, **TOL)


@pytest.mark.parametrize("tag,filters,dist", [("one", (64,), False), ("two", (32, 64), False), ("dist", (16,), True)])
def test_pfn_vs_reference_executed_classes(tag, filters, dist):
    """PillarFeatureNet against the outputs of the REFERENCE'S OWN classes (tests/golden/pillars_ref.npz, made by
    tests/golden/make_golden_pillars.py): single last layer, two layers (max-repeat-concat, pillars.py:36-41)
    and with_distance (:92-94)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pillars_ref.npz"))
    net = pillars.PillarFeatureNet(4, True, filters, dist, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE)
    for i, pfn in enumerate(net.pfn_layers):
        pfn.weight.data = torch.from_numpy(g[f"{tag}_w{i}"])
        pfn.bn_weight.data = torch.from_numpy(g[f"{tag}_gamma{i}"])
        pfn.bn_bias.data = torch.from_numpy(g[f"{tag}_beta{i}"])
    net.to(DEV).train()
    out = net(_cu(g["features"]), _cu(g["num_voxels"]), _cu(g["coors"]))
    assert tuple(out.shape) == g[f"{tag}_out"].shape
    np.testing.assert_allclose(out.cpu().numpy(), g[f"{tag}_out"], **TOL)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pfn_default_constructor_two_layers_vs_oracle(mode):
    """The reference's DEFAULT constructor num_filters=(64,128) (pillars.py:47) at a KITTI-sized frame: 9 -> 32
    (non-last, concat to 64) -> 128, batch statistics and running statistics, incl. the running-stat update."""
    pts = synth.lidar_frame(6000, 7)
    v, c, n = capi.points_to_voxel(pts, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, 100, True, 12000)
    coors = pillars_np.merge_coordinates([c])
    rng = np.random.default_rng(17)
    gpu = pillars.PillarFeatureNet(voxel_size=synth.KITTI_VOXEL_SIZE, pc_range=synth.KITTI_PC_RANGE)
    ref = pillars_np.PillarFeatureNet(voxel_size=synth.KITTI_VOXEL_SIZE, pc_range=synth.KITTI_PC_RANGE)
    assert [p.units for p in gpu.pfn_layers] == [32, 128] and not gpu.pfn_layers[0].last_vfe
    for gp, rp in zip(gpu.pfn_layers, ref.pfn_layers):
        w = (rng.standard_normal(tuple(gp.weight.shape)) / 3.0).astype(np.float32)
        gam = rng.uniform(0.5, 1.5, gp.units).astype(np.float32); gam[::7] *= -1
        bet = rng.uniform(-0.2, 0.2, gp.units).astype(np.float32)
        gp.weight.data, gp.bn_weight.data, gp.bn_bias.data = torch.from_numpy(w), torch.from_numpy(gam), torch.from_numpy(bet)
        rp.weight, rp.gamma, rp.beta = w, gam, bet
        if mode == "eval":
            m = rng.uniform(-1, 1, gp.units).astype(np.float32); var = rng.uniform(5, 40, gp.units).astype(np.float32)
            gp._mean.copy_(torch.from_numpy(m)); gp._variance.copy_(torch.from_numpy(var))
            rp._mean, rp._variance, rp.training = m, var, False
    gpu.to(DEV)
    gpu.train(mode == "train")
    gpu.update_running_stats = True
    out = gpu(_cu(v), _cu(n), _cu(coors))
    exp = ref(v, n, coors)
    assert tuple(out.shape) == exp.shape == (v.shape[0], 128)
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-5, atol=2e-5)
    if mode == "train":
        for gp, rp in zip(gpu.pfn_layers, ref.pfn_layers):
            np.testing.assert_allclose(gp._mean.cpu().numpy(), rp._mean, rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(gp._variance.cpu().numpy(), rp._variance, rtol=1e-4, atol=1e-5)


def test_scatter_bit_exact_and_fused_device_path():
    frames = [synth.lidar_frame(20000, 0), synth.lidar_frame(20000, 5, shuffle=True)]
    vox = [capi.points_to_voxel(f, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, 100, True, 12000) for f in frames]
    coors = pillars_np.merge_coordinates([v[1] for v in vox])
    P = coors.shape[0]
    feat = np.random.default_rng(1).standard_normal((P, 64)).astype(np.float32)
    ref = pillars_np.PointPillarsScatter([1, 1, 496, 432], 64)(feat, coors, 2)
    out = pillars.PointPillarsScatter([1, 1, 496, 432], 64)(_cu(feat), _cu(coors), 2)
    assert tuple(out.shape) == (2, 64, 496, 432)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    # an empty frame in the batch stays all-zero (pillars.py:127, 135-136)
    out3 = pillars.PointPillarsScatter([1, 1, 496, 432], 64)(_cu(feat), _cu(coors), 3)
    assert out3[2].abs().sum().item() == 0
    # end to end on device without host sync: voxelise -> PFN -> scatter via num_valid
    gpu, ref_pfn = _pfn_pair(4)
    dv, dc, dn, dnum = pillars.points_to_voxel_device(_cu(frames[0]), synth.KITTI_VOXEL_SIZE,
                                                      synth.KITTI_PC_RANGE, 100, True, 12000)
    dcoors = torch.cat([torch.zeros((dc.shape[0], 1), dtype=torch.int32, device=DEV), dc], 1)
    f = gpu(dv, dn, dcoors, num_valid=dnum)
    canvas = pillars.PointPillarsScatter([1, 1, 496, 432], 64)(f, dcoors, 1, num_valid=dnum)
    v, c, n = vox[0]
    cc = pillars_np.merge_coordinates([c])
    exp = pillars_np.PointPillarsScatter([1, 1, 496, 432], 64)(ref_pfn(v, n, cc), cc, 1)
    np.testing.assert_allclose(canvas.cpu().numpy(), exp, **TOL)


# ---- N4: batched voxelisation + merged layout, anchors mask, BEV map -- vs the reference's own outputs
def _n4(golden_dir):
    return np.load(os.path.join(golden_dir, "pillar_batch_ref.npz"))


@pytest.mark.parametrize("case", ["a", "b"])
def test_voxelize_batch_merged_golden(golden_dir, case):
    """points_to_voxel per frame + merge_second_batch (data/preprocess.py:16-42) in one device call: ragged frames,
    an empty frame, frames that hit the max_voxels break; bit-exact vs the reference's own output."""
    g = _n4(golden_dir)
    frames = [g[f"{case}_points{i}"] for i in range(len(g[f"{case}_sizes"]))]
    mp, mv = [int(x) for x in g[f"{case}_cfg"]]
    merged, fv = pillars.merge_second_batch_voxels(frames, g["voxel_size"], g["range"], mp, True, mv)
    np.testing.assert_array_equal(fv, g[f"{case}_frame_voxels"])
    for k in ("coordinates", "num_points", "voxels"):
        assert merged[k].dtype == g[f"{case}_{k}"].dtype, k
        np.testing.assert_array_equal(merged[k], g[f"{case}_{k}"], err_msg=k)
    # rows past the total are zero on the device form
    v, c, n, fvd, tot = pillars.points_to_voxel_batch_device([_cu(f) for f in frames], g["voxel_size"], g["range"], mp, True, mv)
    m = int(tot.item())
    assert m == int(fv.sum())
    assert not v[m:].any() and not c[m:].any() and not n[m:].any()


def test_voxelize_batch_k5_matches_single_frame():
    """BASELINE config 4 (20 000-point frames), batch of 3: every frame's rows equal the single-frame voxeliser's
    (itself pinned to the reference's numba kernel) and the merged coordinates carry the frame index."""
    frames = [synth.lidar_frame(20000, s, s == 1) for s in range(3)]
    v, c, n, fv, tot = pillars.points_to_voxel_batch_device([_cu(f) for f in frames], synth.KITTI_VOXEL_SIZE,
                                                            synth.KITTI_PC_RANGE, synth.KITTI_MAX_POINTS, True,
                                                            synth.KITTI_MAX_VOXELS)
    fv = fv.cpu().numpy()
    assert int(tot.item()) == int(fv.sum())
    r0 = 0
    for b, f in enumerate(frames):
        v1, c1, n1 = pillars.points_to_voxel(f, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE, synth.KITTI_MAX_POINTS,
                                             True, synth.KITTI_MAX_VOXELS)
        m = v1.shape[0]
        assert fv[b] == m
        np.testing.assert_array_equal(v[r0:r0 + m].cpu().numpy(), v1)
        np.testing.assert_array_equal(n[r0:r0 + m].cpu().numpy(), n1)
        cc = c[r0:r0 + m].cpu().numpy()
        assert (cc[:, 0] == b).all()
        np.testing.assert_array_equal(cc[:, 1:], c1)
        r0 += m


def test_anchors_mask_golden(golden_dir):
    g = _n4(golden_dir)
    grid = g["am_grid"]
    shape = tuple(int(x) for x in grid[::-1][1:])
    dense = pillars.sparse_sum_for_anchors_mask(g["am_coors"], shape)
    np.testing.assert_array_equal(dense, g["am_dense"])
    cum, area = pillars.anchors_area_from_coors(g["am_coors"], shape, g["am_anchors_bv"], g["voxel_size"], g["range"], grid)
    np.testing.assert_array_equal(cum, g["am_cum"])
    np.testing.assert_array_equal(area, g["am_area"])
    # merged (b,z,y,x) rows and a device-side count
    c4 = np.pad(g["am_coors"], ((0, 5), (1, 0)))
    nv = torch.tensor([g["am_coors"].shape[0]], dtype=torch.int32, device=DEV)
    d2 = pillars.sparse_sum_for_anchors_mask(_cu(c4), shape, num_valid=nv)
    np.testing.assert_array_equal(d2.cpu().numpy(), g["am_dense"])


@pytest.mark.parametrize("key,refl,mv", [("bev_plain", False, 40000), ("bev_refl", True, 40000), ("bev_refl_break", True, 700)])
def test_points_to_bev_golden(golden_dir, key, refl, mv):
    g = _n4(golden_dir)
    bev = pillars.points_to_bev(g["bev_points"], g["bev_voxel_size"], g["range"], refl, max_voxels=mv)
    assert bev.shape == g[key].shape
    np.testing.assert_array_equal(bev, g[key])


def test_points_to_bev_random_vs_oracle():
    rng = np.random.default_rng(3)
    pts = rng.uniform(-1, 1, (6000, 4)).astype(np.float32)
    pts[:, 0] = pts[:, 0] * 6 + 5
    pts[:, 1] *= 5
    pts[:, 2] = np.round(pts[:, 2] * 10) / 4 - 1
    vs, rg = np.array([0.25, 0.5, 0.5], np.float32), np.array([0, -4, -3, 10, 4, 1], np.float32)
    for refl in (False, True):
        for mv in (40000, 900):
            want = pillars_np.points_to_bev(pts, vs, rg, refl, max_voxels=mv)
            got = pillars.points_to_bev(pts, vs, rg, refl, max_voxels=mv)
            np.testing.assert_array_equal(got, want)
