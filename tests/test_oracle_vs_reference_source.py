"""The oracle against vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE
(pointnet2_basic_layers.py, unmodified) over a NumPy-backed stand-in for its paddle calls
(tests/golden/paddle_stub.py, tests/golden/make_golden_layers.py -> tests/golden/layers_ref.npz).

This pins the oracle's LOGIC -- control flow, operation and concat order, masks, the sort / pad of the
ball query, the FPS initial distance of 1.0, float32-encoded indices, the 3-NN interpolation quirk --
to the reference's code rather than to a hand restatement.  It does not pin arithmetic that happens
inside Paddle's kernels (see the stub's docstring)."""
import os

import numpy as np
import pytest

from oracle import capi, layers_np


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "layers_ref.npz"))


def test_known_answers_come_from_the_reference_code(g):
    """SURVEY.md 8(c) K1-K4, as the reference's functions return them."""
    np.testing.assert_array_equal(g["kat_k1_fps"], [[0, 3, 2]])
    np.testing.assert_array_equal(g["kat_k2_fps"], [[0, 1, 2]])            # running distance clamps at 1.0
    np.testing.assert_array_equal(g["kat_k3_ball3"], [[[0, 1, 2]]])
    np.testing.assert_array_equal(g["kat_k3_ball6"], [[[0, 1, 2, 4, 0, 0]]])
    np.testing.assert_array_equal(g["kat_k4_ball6"], [[[0, 1, 2, 0, 0, 0]]])  # fp32 0.2 is outside r = 0.2
    for k, npoint in (("k1", 3), ("k2", 3)):
        xyz = g[f"kat_{k}_xyz"]
        np.testing.assert_array_equal(layers_np.farthest_point_sample(xyz, npoint, start_idx=[0]), g[f"kat_{k}_fps"])
        np.testing.assert_array_equal(capi.farthest_point_sample(xyz, npoint, np.zeros(1, np.int64)),
                                      g[f"kat_{k}_fps"].astype(np.int64))
    for k, ns in (("k3", (3, 6)), ("k4", (6,))):
        xyz = g[f"kat_{k}_xyz"]
        for n in ns:
            want = g[f"kat_{k}_ball{n}"]
            np.testing.assert_array_equal(layers_np.query_ball_point(0.2, n, xyz, xyz[:, :1]), want)
            np.testing.assert_array_equal(capi.query_ball_point(0.2, n, xyz, xyz[:, :1])[0], want)


def test_primitives_match_the_reference_code(g):
    xyz, feats, start = g["prim_xyz"], g["prim_feats"], g["prim_start"]
    S = g["prim_fps"].shape[1]
    fps = layers_np.farthest_point_sample(xyz, S, start_idx=start)
    assert fps.dtype == np.float32                                          # layers.py:74
    np.testing.assert_array_equal(fps, g["prim_fps"])
    np.testing.assert_array_equal(capi.farthest_point_sample(xyz, S, start), g["prim_fps"].astype(np.int64))
    new_xyz = layers_np.index_points(xyz, fps)
    np.testing.assert_array_equal(new_xyz, g["prim_new_xyz"])
    np.testing.assert_array_equal(layers_np.square_distance(new_xyz, xyz), g["prim_sqdist"])
    for key in [k for k in g.files if k.startswith("prim_ball_")]:
        r, k = float(key.split("_r")[1].split("_k")[0]), int(key.split("_k")[1])
        np.testing.assert_array_equal(layers_np.query_ball_point(r, k, xyz, new_xyz), g[key], err_msg=key)
    a, b, c, d = layers_np.sample_and_group(S, 0.3, 16, xyz, feats, returnfps=True, start_idx=start)
    np.testing.assert_array_equal(a, g["prim_sg_new_xyz"])
    np.testing.assert_array_equal(b, g["prim_sg_new_points"])               # xyz first (:151)
    np.testing.assert_array_equal(c, g["prim_sg_grouped_xyz"])
    np.testing.assert_array_equal(d, g["prim_sg_fps"])
    _, b0 = layers_np.sample_and_group(S, 0.3, 16, xyz, None, start_idx=start)
    np.testing.assert_array_equal(b0, g["prim_sg0_new_points"])
    a, b = layers_np.sample_and_group_all(xyz, feats)
    np.testing.assert_array_equal(a, g["prim_sga_new_xyz"])
    np.testing.assert_array_equal(b, g["prim_sga_new_points"])


def test_c_oracle_ball_query_vs_the_reference_code(g):
    """The arithmetic-pinned C restatement (fma expansion form -- the arithmetic the CUDA kernels use) gives the
    reference code's indices; in general the two may differ where a distance lies within rounding of r^2
    (~1 in 1e7 comparisons) -- none of the committed cases has such a pair.  Same for the
    distance matrix itself."""
    xyz, new_xyz = g["prim_xyz"], g["prim_new_xyz"]
    for key in [k for k in g.files if k.startswith("prim_ball_")]:
        r, k = float(key.split("_r")[1].split("_k")[0]), int(key.split("_k")[1])
        got, _ = capi.query_ball_point(r, k, xyz, new_xyz)
        np.testing.assert_array_equal(got, g[key], err_msg=key)   # on these (committed) inputs: no borderline pair
    np.testing.assert_array_equal(capi.square_distance(new_xyz, xyz), g["prim_sqdist"])


def test_forward_passes_with_empty_mlps_match_the_reference_code(g):
    xyz_cf = np.ascontiguousarray(g["prim_xyz"].transpose(0, 2, 1))
    feats_cf = np.ascontiguousarray(g["prim_feats"].transpose(0, 2, 1))
    start, D = g["prim_start"], g["prim_feats"].shape[2]
    a, b = layers_np.PointNetSetAbstraction(32, 0.3, 16, 3 + D, [], False)(xyz_cf, feats_cf, start_idx=start)
    np.testing.assert_array_equal(a, g["sa_new_xyz"])
    np.testing.assert_array_equal(b, g["sa_new_points"])                    # max over K of [xyz_norm, feats]
    a, b = layers_np.PointNetSetAbstraction(None, None, None, 3 + D, [], True)(xyz_cf, feats_cf)
    np.testing.assert_array_equal(a, g["sa_all_new_xyz"])
    np.testing.assert_array_equal(b, g["sa_all_new_points"])
    a, b = layers_np.PointNetSetAbstractionMsg(32, [0.2, 0.4], [8, 16], D, [[], []])(xyz_cf, feats_cf, start_idx=start)
    np.testing.assert_array_equal(a, g["msg_new_xyz"])
    np.testing.assert_array_equal(b, g["msg_new_points"])                   # features first (:267), branches on C
    xyz2, p2 = g["fp_xyz2"], g["fp_points2"]
    fp = layers_np.PointNetFeaturePropagation(D + 9, [])
    np.testing.assert_array_equal(fp(xyz_cf, xyz2, feats_cf, p2), g["fp_out"])
    np.testing.assert_array_equal(layers_np.PointNetFeaturePropagation(9, [])(xyz_cf, xyz2, None, p2), g["fp_out_nop1"])
    np.testing.assert_array_equal(fp(xyz_cf, xyz2[:, :, :1], feats_cf, p2[:, :, :1]), g["fp_out_s1"])


# ---------------------------------------------------------------------------------- pillar path
@pytest.fixture(scope="module")
def gp(golden_dir):
    return np.load(os.path.join(golden_dir, "pillars_ref.npz"))


@pytest.mark.parametrize("tag,filters,dist", [("one", (64,), False), ("two", (32, 64), False), ("dist", (16,), True)])
def test_pillar_feature_net_matches_the_reference_classes(gp, tag, filters, dist):
    """tests/golden/make_golden_pillars.py: the reference's PillarFeatureNet / PFNLayer executed over the
    stub.  The decorated tensor (input of the first Linear) is an exact reference for pillars.py:81-102;
    the output additionally goes through the stub's Linear / BatchNorm1D arithmetic."""
    from oracle import pillars_np
    from papc_b200 import synth
    net = pillars_np.PillarFeatureNet(4, True, filters, dist, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE)
    for i, pfn in enumerate(net.pfn_layers):
        pfn.weight, pfn.gamma, pfn.beta = gp[f"{tag}_w{i}"], gp[f"{tag}_gamma{i}"], gp[f"{tag}_beta{i}"]
    feats, num, coors = gp["features"], gp["num_voxels"], gp["coors"]
    np.testing.assert_array_equal(net.decorate(feats, num, coors), gp[f"{tag}_decorated"])
    out = net(feats, num, coors)
    assert out.shape == gp[f"{tag}_out"].shape
    np.testing.assert_allclose(out, gp[f"{tag}_out"], rtol=1e-6, atol=1e-6)


def test_pillar_scatter_matches_the_reference_class(gp):
    from oracle import pillars_np
    nx, ny = int(gp["nx"]), int(gp["ny"])
    vf = gp["one_out"]
    sc = pillars_np.PointPillarsScatter([1, 1, ny, nx], num_input_features=vf.shape[1])
    np.testing.assert_array_equal(sc.forward(vf, gp["coors"], 2), gp["canvas_b2"])
    np.testing.assert_array_equal(sc.forward(vf, gp["coors_b3"], 3), gp["canvas_b3"])   # sample 1 is empty
    assert not gp["canvas_b3"][1].any()


# ---------------------------------------------------------------------------------- model definitions
MODEL_CASES = [("PointNet2_SSG_Clas", False, False), ("PointNet2_SSG_Clas", True, False),
               ("PointNet2_MSG_Clas", False, False), ("PointNet2_SSG_Seg", False, True),
               ("PointNet2_MSG_Seg", True, True)]


@pytest.mark.parametrize("name,normal_channel,seg", MODEL_CASES)
def test_model_definitions_match_the_reference_classes(golden_dir, name, normal_channel, seg):
    """tests/golden/make_golden_models.py: the reference's PointNet2_* classes executed over the stub.  Pins
    the oracle's model wiring, ``Categorical``, the heads and which BatchNorms follow eval(); the conv /
    linear / batch-norm arithmetic on both sides is an fp64-accumulated restatement (hence 1e-5, not 0)."""
    import sys
    sys.path.insert(0, golden_dir)
    import param_gen
    from oracle import models_np
    gm = np.load(os.path.join(golden_dir, "models_ref.npz"))
    tag = name + ("_nc" if normal_channel else "")
    model = getattr(models_np, name)(normal_channel=normal_channel)
    n = len(param_gen.install(model, tag))
    assert n == {"PointNet2_SSG_Clas": 23, "PointNet2_MSG_Clas": 47, "PointNet2_SSG_Seg": 35, "PointNet2_MSG_Seg": 51}[name]
    x = np.concatenate([gm["xyz"], gm["normals"]], 1) if normal_channel else gm["xyz"]
    inputs = (x, gm["labels"]) if seg else x
    start = (gm["start1"], gm["start2"])

    def check(y, key):
        if seg:
            np.testing.assert_allclose(y[:, ::8], gm[key + ":sub8"], rtol=1e-5, atol=1e-5)
            s = np.array([y.astype(np.float64).sum(), np.abs(y.astype(np.float64)).sum()])
            np.testing.assert_allclose(s, gm[key + ":sum"], rtol=1e-6)
        else:
            np.testing.assert_allclose(y, gm[key], rtol=1e-5, atol=1e-5)

    check(model.eval()(inputs, start_idx=start), f"{tag}:eval")
    check(model.train()(inputs, start_idx=start), f"{tag}:train")
    np.testing.assert_allclose(model.bn1._mean, gm[f"{tag}:bn1_mean_after_train"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(model.bn1._variance, gm[f"{tag}:bn1_var_after_train"], rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------- pillar glue (pure NumPy in the reference)
def test_voxel_generator_and_batch_merge_match_the_reference(golden_dir):
    """tests/golden/make_golden_pillar_glue.py: VoxelGenerator geometry (voxel_generator.py:5-43) and the batch
    layout of merge_second_batch (preprocess.py:16-42), produced by the reference's own code; checked on the
    oracle AND on the product's host-side mirrors (both run without a GPU)."""
    torch = pytest.importorskip("torch")
    from oracle import pillars_np
    from papc_b200 import pillars
    gg = np.load(os.path.join(golden_dir, "pillar_glue_ref.npz"))
    for tag in ("yaml", "default", "odd"):
        a = gg[f"{tag}_args"]
        gen = pillars.VoxelGenerator(tuple(a[:3]), tuple(a[3:]), 100, 12000)
        for got, key in ((gen.voxel_size, "voxel_size"), (gen.point_cloud_range, "range"), (gen.grid_size, "grid")):
            assert got.dtype == gg[f"{tag}_{key}"].dtype
            np.testing.assert_array_equal(got, gg[f"{tag}_{key}"])
        assert gen.max_num_points_per_voxel == 100
    coors = [gg[f"batch{i}_coordinates"] for i in range(3)]
    np.testing.assert_array_equal(pillars_np.merge_coordinates(coors), gg["merged_coordinates"])
    got = pillars.merge_coordinates([torch.from_numpy(c) for c in coors])
    assert got.dtype == torch.int32
    np.testing.assert_array_equal(got.numpy(), gg["merged_coordinates"])
    np.testing.assert_array_equal(np.concatenate([gg[f"batch{i}_voxels"] for i in range(3)]), gg["merged_voxels"])
    np.testing.assert_array_equal(np.concatenate([gg[f"batch{i}_num_points"] for i in range(3)]), gg["merged_num_points"])


def test_error_behaviour_of_the_reference_code(g):
    """SURVEY 8(b) error conventions, as the reference's own code behaves: an empty ball leaves N in every
    slot (no exception until the gather, which raises IndexError); nsample > N raises IndexError.  The oracle
    behaves the same; the product returns N-filled groups (``check_empty=True`` raises IndexError) and turns
    nsample > N into a ValueError before launch (tests/test_gpu_sa.py::test_ball_query_errors)."""
    assert str(g["err_empty_ball_query"]) == "none"
    assert (g["err_empty_ball_value"] == 6).all()
    assert str(g["err_empty_ball_gather"]) == "IndexError"
    assert str(g["err_nsample_gt_n"]) == "IndexError"
    k3 = g["kat_k3_xyz"]
    far = np.full((1, 1, 3), 50.0, np.float32)
    idx = layers_np.query_ball_point(0.2, 4, k3, far)
    np.testing.assert_array_equal(idx, g["err_empty_ball_value"])
    with pytest.raises(IndexError):
        layers_np.index_points(k3, idx)
    with pytest.raises(IndexError):
        layers_np.query_ball_point(0.2, 7, k3, k3[:, :1])


def test_c_oracle_indices_at_baseline_config_2(golden_dir):
    """tests/golden/make_golden_c2.py: FPS and ball-query indices of the bench workload (32 x 1024 points, sa1 and
    sa2 of PointNet++ SSG) from the reference's own code; the C oracle -- the arithmetic the kernels use --
    reproduces all 20 480 groups."""
    from papc_b200 import synth
    gc = np.load(os.path.join(golden_dir, "c2_indices_ref.npz"))
    B, N = 32, 1024
    xyz = np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1))
    f1 = capi.farthest_point_sample(xyz, 512, synth.fps_start(B, N, seed=1))
    np.testing.assert_array_equal(f1, gc["fps1"])
    l1 = layers_np.index_points(xyz, f1)
    np.testing.assert_array_equal(capi.query_ball_point(0.2, 32, xyz, l1)[0], gc["ball1"])
    f2 = capi.farthest_point_sample(l1, 128, np.zeros(B, np.int64))
    np.testing.assert_array_equal(f2, gc["fps2"])
    np.testing.assert_array_equal(capi.query_ball_point(0.4, 64, l1, layers_np.index_points(l1, f2))[0], gc["ball2"])


def _digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a).astype(np.int64)).tobytes()).hexdigest()


def test_c_oracle_indices_at_baseline_config_3(golden_dir):
    """tests/golden/make_golden_c3.py: PointNet++ MSG segment, 16 x 2048 points, three + two radii (2.2 M indices,
    committed as sha256 digests + cloud 0): the C oracle reproduces every array of the reference's code."""
    from papc_b200 import synth
    gc = np.load(os.path.join(golden_dir, "c3_indices_ref.npz"))
    B, N = 16, 2048
    xyz = np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1))
    f1 = capi.farthest_point_sample(xyz, 512, synth.fps_start(B, N, seed=1))
    assert _digest(f1) == str(gc["fps1:sha256"])
    np.testing.assert_array_equal(f1[0], gc["fps1:cloud0"])
    l1 = layers_np.index_points(xyz, f1)
    for r, k in ((0.1, 32), (0.2, 64), (0.4, 128)):
        b, empty = capi.query_ball_point(r, k, xyz, l1)
        assert empty == 0 and _digest(b) == str(gc[f"sa1_ball_r{r}_k{k}:sha256"]), (r, k)
        np.testing.assert_array_equal(b[0], gc[f"sa1_ball_r{r}_k{k}:cloud0"])
    f2 = capi.farthest_point_sample(l1, 128, np.zeros(B, np.int64))
    assert _digest(f2) == str(gc["fps2:sha256"])
    l2 = layers_np.index_points(l1, f2)
    for r, k in ((0.4, 64), (0.8, 128)):
        assert _digest(capi.query_ball_point(r, k, l1, l2)[0]) == str(gc[f"sa2_ball_r{r}_k{k}:sha256"]), (r, k)


# ---- N4: batched voxelisation layout, anchors mask, BEV map (tests/golden/make_golden_pillar_batch.py)
def _n4():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pillar_batch_ref.npz"))


@pytest.mark.parametrize("case", ["a", "b"])
def test_n4_merged_voxel_batch_oracle(case):
    from oracle import pillars_np
    g = _n4()
    frames = [g[f"{case}_points{i}"] for i in range(len(g[f"{case}_sizes"]))]
    mp, mv = [int(x) for x in g[f"{case}_cfg"]]
    merged, fv = pillars_np.merge_second_batch_voxels(frames, g["voxel_size"], g["range"], mp, True, mv)
    for k in ("voxels", "num_points", "coordinates"):
        np.testing.assert_array_equal(merged[k], g[f"{case}_{k}"], err_msg=k)
    np.testing.assert_array_equal(fv, g[f"{case}_frame_voxels"])


def test_n4_anchors_mask_oracle():
    from oracle import pillars_np
    g = _n4()
    grid = g["am_grid"]
    dense = pillars_np.sparse_sum_for_anchors_mask(g["am_coors"], tuple(grid[::-1][1:]))
    np.testing.assert_array_equal(dense, g["am_dense"])
    cum = dense.cumsum(0).cumsum(1)
    np.testing.assert_array_equal(cum, g["am_cum"])
    area = pillars_np.fused_get_anchors_area(cum, g["am_anchors_bv"], g["voxel_size"], g["range"], grid)
    np.testing.assert_array_equal(area, g["am_area"])


@pytest.mark.parametrize("key,refl,mv", [("bev_plain", False, 40000), ("bev_refl", True, 40000), ("bev_refl_break", True, 700)])
def test_n4_points_to_bev_oracle(key, refl, mv):
    from oracle import pillars_np
    g = _n4()
    bev = pillars_np.points_to_bev(g["bev_points"], g["bev_voxel_size"], g["range"], refl, max_voxels=mv)
    np.testing.assert_array_equal(bev, g[key])
