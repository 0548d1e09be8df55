"""CUDA (through the C ABI) against vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE over a
NumPy stand-in for its paddle calls (tests/golden/layers_ref.npz; see tests/golden/paddle_stub.py
and tests/test_oracle_vs_reference_source.py for what that pins).  Indices, gathers and the grouped
tensors are compared bit-exactly; the interpolated features within 1e-5."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from papc_b200 import layers  # noqa: E402

DEV = "cuda:0"


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "layers_ref.npz"))


def test_known_answers(g):
    for k in ("k1", "k2"):
        out = layers.farthest_point_sample(_cu(g[f"kat_{k}_xyz"]), 3, start_idx=[0])
        assert out.dtype == torch.float32                                   # layers.py:74
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"kat_{k}_fps"])
    for k, ns in (("k3", (3, 6)), ("k4", (6,))):
        xyz = g[f"kat_{k}_xyz"]
        for n in ns:
            out = layers.query_ball_point(0.2, n, _cu(xyz), _cu(xyz[:, :1]))
            np.testing.assert_array_equal(out.cpu().numpy(), g[f"kat_{k}_ball{n}"])


def test_primitives(g):
    xyz, feats, start = g["prim_xyz"], g["prim_feats"], g["prim_start"]
    S = g["prim_fps"].shape[1]
    idx, new_xyz = layers.farthest_point_sample_idx(_cu(xyz), S, _cu(start), return_xyz=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["prim_fps"].astype(np.int64))
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), g["prim_new_xyz"])
    fps_f32 = layers.farthest_point_sample(_cu(xyz), S, start_idx=_cu(start))
    np.testing.assert_array_equal(layers.index_points(_cu(xyz), fps_f32).cpu().numpy(), g["prim_new_xyz"])
    np.testing.assert_array_equal(layers.square_distance(_cu(g["prim_new_xyz"]), _cu(xyz)).cpu().numpy(),
                                  g["prim_sqdist"])
    for key in [k for k in g.files if k.startswith("prim_ball_")]:
        r, k = float(key.split("_r")[1].split("_k")[0]), int(key.split("_k")[1])
        out = layers.query_ball_point(r, k, _cu(xyz), _cu(g["prim_new_xyz"]))
        np.testing.assert_array_equal(out.cpu().numpy(), g[key], err_msg=key)


def test_sample_and_group(g):
    xyz, feats, start = g["prim_xyz"], g["prim_feats"], g["prim_start"]
    S = g["prim_sg_fps"].shape[1]
    a, b, c, d = layers.sample_and_group(S, 0.3, 16, _cu(xyz), _cu(feats), returnfps=True, start_idx=_cu(start))
    np.testing.assert_array_equal(a.cpu().numpy(), g["prim_sg_new_xyz"])
    np.testing.assert_array_equal(d.cpu().numpy(), g["prim_sg_fps"])
    np.testing.assert_array_equal(c.cpu().numpy(), g["prim_sg_grouped_xyz"])
    np.testing.assert_array_equal(b.cpu().numpy(), g["prim_sg_new_points"])
    _, b0 = layers.sample_and_group(S, 0.3, 16, _cu(xyz), None, start_idx=_cu(start))
    np.testing.assert_array_equal(b0.cpu().numpy(), g["prim_sg0_new_points"])
    a, b = layers.sample_and_group_all(_cu(xyz), _cu(feats))
    np.testing.assert_array_equal(a.cpu().numpy(), g["prim_sga_new_xyz"])
    np.testing.assert_array_equal(b.cpu().numpy(), g["prim_sga_new_points"])


def test_feature_interpolation(g):
    """layers.py:306-329 as the reference's forward (empty mlp) returns it, [B, D1+D2, N]."""
    xyz1 = np.ascontiguousarray(g["prim_xyz"])                              # [B,N,3]
    p1 = np.ascontiguousarray(g["prim_feats"])                              # [B,N,D1]
    xyz2 = np.ascontiguousarray(g["fp_xyz2"].transpose(0, 2, 1))            # [B,S,3]
    p2 = np.ascontiguousarray(g["fp_points2"].transpose(0, 2, 1))           # [B,S,D2]
    B, N, _ = xyz1.shape
    for key, a1, a2, b2 in (("fp_out", p1, xyz2, p2), ("fp_out_nop1", None, xyz2, p2),
                            ("fp_out_s1", p1, xyz2[:, :1], p2[:, :1])):
        rows, cin = layers.feature_interpolate(_cu(xyz1), _cu(a2), _cu(a1) if a1 is not None else None, _cu(b2))
        got = rows.cpu().numpy().reshape(B, N, -1)[:, :, :cin].transpose(0, 2, 1)
        np.testing.assert_allclose(got, g[key], rtol=1e-5, atol=1e-5, err_msg=key)


def test_indices_at_baseline_config_2(golden_dir):
    """The bench workload's sampling / grouping indices (32 x 1024 points; sa1: FPS 512, r 0.2, K 32; sa2: FPS 128,
    r 0.4, K 64) against the reference's own code (tests/golden/make_golden_c2.py): bit-exact."""
    from papc_b200 import synth
    gc = np.load(os.path.join(golden_dir, "c2_indices_ref.npz"))
    B, N = 32, 1024
    xyz = _cu(np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1)))
    f1, l1 = layers.farthest_point_sample_idx(xyz, 512, _cu(synth.fps_start(B, N, seed=1)), return_xyz=True)
    np.testing.assert_array_equal(f1.cpu().numpy(), gc["fps1"])
    np.testing.assert_array_equal(layers.query_ball_point(0.2, 32, xyz, l1).cpu().numpy(), gc["ball1"])
    f2, l2 = layers.farthest_point_sample_idx(l1, 128, _cu(np.zeros(B, np.int64)), return_xyz=True)
    np.testing.assert_array_equal(f2.cpu().numpy(), gc["fps2"])
    np.testing.assert_array_equal(layers.query_ball_point(0.4, 64, l1, l2).cpu().numpy(), gc["ball2"])


def test_indices_at_baseline_config_3(golden_dir):
    """PointNet++ MSG segment (16 x 2048 points; sa1: FPS 512, r .1/.2/.4, K 32/64/128; sa2: FPS 128, r .4/.8,
    K 64/128) against sha256 digests of the reference code's index arrays (tests/golden/make_golden_c3.py)."""
    import hashlib
    from papc_b200 import synth

    def digest(t):
        return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy().astype(np.int64)).tobytes()).hexdigest()

    gc = np.load(os.path.join(golden_dir, "c3_indices_ref.npz"))
    B, N = 16, 2048
    xyz = _cu(np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1)))
    f1, l1 = layers.farthest_point_sample_idx(xyz, 512, _cu(synth.fps_start(B, N, seed=1)), return_xyz=True)
    np.testing.assert_array_equal(f1[0].cpu().numpy(), gc["fps1:cloud0"])
    assert digest(f1) == str(gc["fps1:sha256"])
    for r, k in ((0.1, 32), (0.2, 64), (0.4, 128)):
        b = layers.query_ball_point(r, k, xyz, l1)
        np.testing.assert_array_equal(b[0].cpu().numpy(), gc[f"sa1_ball_r{r}_k{k}:cloud0"], err_msg=str((r, k)))
        assert digest(b) == str(gc[f"sa1_ball_r{r}_k{k}:sha256"]), (r, k)
    f2, l2 = layers.farthest_point_sample_idx(l1, 128, _cu(np.zeros(B, np.int64)), return_xyz=True)
    assert digest(f2) == str(gc["fps2:sha256"])
    for r, k in ((0.4, 64), (0.8, 128)):
        assert digest(layers.query_ball_point(r, k, l1, l2)) == str(gc[f"sa2_ball_r{r}_k{k}:sha256"]), (r, k)
