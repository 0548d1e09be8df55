"""Generate tests/golden/layers_ref.npz by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/PAPC/models/layers/pointnet2_basic_layers.py) over a NumPy-backed stand-in for the
paddle calls it makes (tests/golden/paddle_stub.py -- read its docstring for what this does and does
not pin).  Build-container only:  python tests/golden/make_golden_layers.py

Cases (all inputs are stored next to the outputs):
  kat_*   the hand-derived known answers K1-K4 of SURVEY.md 8(c), now produced by the reference code
  prim_*  square_distance / index_points / farthest_point_sample / query_ball_point /
          sample_and_group / sample_and_group_all on seeded normalised clouds
  sa_* / msg_* / fp_*  the forward passes of PointNetSetAbstraction, PointNetSetAbstractionMsg and
          PointNetFeaturePropagation built with EMPTY mlp lists: grouping, concat order, transposes,
          max-pool and the 3-NN interpolation exactly as the reference wires them
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import paddle_stub  # noqa: E402
from papc_b200 import synth  # noqa: E402

REF = "/root/reference/PAPC/models/layers/pointnet2_basic_layers.py"


def load_reference():
    P = paddle_stub.install()
    mod = types.ModuleType("ref_layers")
    mod.__file__ = REF
    exec(compile(open(REF).read(), REF, "exec"), mod.__dict__)
    return P, mod


if __name__ == "__main__":
    P, R = load_reference()
    T = P.to_tensor
    out = {}

    def fps(xyz, npoint, start):
        paddle_stub._next_randint.append(np.asarray(start, dtype=np.int64))
        return R.farthest_point_sample(T(xyz), npoint).numpy()

    # ---- K1-K4
    line = lambda xs: np.stack([np.array(xs, np.float32), np.zeros(len(xs), np.float32),
                                np.zeros(len(xs), np.float32)], 1)[None]
    out["kat_k1_xyz"] = line([0, .1, .5, .9]); out["kat_k1_fps"] = fps(out["kat_k1_xyz"], 3, [0])
    out["kat_k2_xyz"] = line([0, 2, 3]); out["kat_k2_fps"] = fps(out["kat_k2_xyz"], 3, [0])
    k3 = line([0, .1, .15, .3, .19, .7]); out["kat_k3_xyz"] = k3
    out["kat_k3_ball3"] = R.query_ball_point(0.2, 3, T(k3), T(k3[:, :1])).numpy()
    out["kat_k3_ball6"] = R.query_ball_point(0.2, 6, T(k3), T(k3[:, :1])).numpy()
    k4 = line([0, .1, .15, .3, .2, .7]); out["kat_k4_xyz"] = k4
    out["kat_k4_ball6"] = R.query_ball_point(0.2, 6, T(k4), T(k4[:, :1])).numpy()

    # ---- primitives on seeded clouds
    B, N, S, D = 3, 256, 64, 5
    rng = np.random.default_rng(42)
    xyz = np.ascontiguousarray(synth.clouds(B, N, seed=7).transpose(0, 2, 1))       # [B,N,3]
    feats = rng.standard_normal((B, N, D)).astype(np.float32)
    start = synth.fps_start(B, N, seed=8)
    out.update(prim_xyz=xyz, prim_feats=feats, prim_start=start)
    out["prim_fps"] = fps(xyz, S, start)                                              # float32-encoded
    new_xyz = R.index_points(T(xyz), T(out["prim_fps"])).numpy()
    out["prim_new_xyz"] = new_xyz
    out["prim_sqdist"] = R.square_distance(T(new_xyz), T(xyz)).numpy()
    for r, k in ((0.2, 8), (0.4, 16), (0.8, 32), (0.05, 4)):
        out[f"prim_ball_r{r}_k{k}"] = R.query_ball_point(r, k, T(xyz), T(new_xyz)).numpy()
    paddle_stub._next_randint.append(start)
    a, b, c, d = R.sample_and_group(S, 0.3, 16, T(xyz), T(feats), returnfps=True)
    out.update(prim_sg_new_xyz=a.numpy(), prim_sg_new_points=b.numpy(), prim_sg_grouped_xyz=c.numpy(),
               prim_sg_fps=d.numpy())
    paddle_stub._next_randint.append(start)
    a, b = R.sample_and_group(S, 0.3, 16, T(xyz), None)
    out.update(prim_sg0_new_points=b.numpy())
    a, b = R.sample_and_group_all(T(xyz), T(feats))
    out.update(prim_sga_new_xyz=a.numpy(), prim_sga_new_points=b.numpy())

    # ---- forward passes with empty mlp lists
    xyz_cf = np.ascontiguousarray(xyz.transpose(0, 2, 1))                             # [B,3,N]
    feats_cf = np.ascontiguousarray(feats.transpose(0, 2, 1))                         # [B,D,N]
    paddle_stub._next_randint.append(start)
    a, b = R.PointNetSetAbstraction(32, 0.3, 16, 3 + D, [], False)(T(xyz_cf), T(feats_cf))
    out.update(sa_new_xyz=a.numpy(), sa_new_points=b.numpy())
    a, b = R.PointNetSetAbstraction(None, None, None, 3 + D, [], True)(T(xyz_cf), T(feats_cf))
    out.update(sa_all_new_xyz=a.numpy(), sa_all_new_points=b.numpy())
    paddle_stub._next_randint.append(start)
    a, b = R.PointNetSetAbstractionMsg(32, [0.2, 0.4], [8, 16], D, [[], []])(T(xyz_cf), T(feats_cf))
    out.update(msg_new_xyz=a.numpy(), msg_new_points=b.numpy())
    sub = rng.permutation(N)[:48]
    xyz2 = np.ascontiguousarray(xyz_cf[:, :, sub]); p2 = rng.standard_normal((B, 9, 48)).astype(np.float32)
    out.update(fp_xyz2=xyz2, fp_points2=p2)
    out["fp_out"] = R.PointNetFeaturePropagation(D + 9, [])(T(xyz_cf), T(xyz2), T(feats_cf), T(p2)).numpy()
    out["fp_out_nop1"] = R.PointNetFeaturePropagation(9, [])(T(xyz_cf), T(xyz2), None, T(p2)).numpy()
    out["fp_out_s1"] = R.PointNetFeaturePropagation(D + 9, [])(T(xyz_cf), T(xyz2[:, :, :1]), T(feats_cf),
                                                               T(p2[:, :, :1])).numpy()
    # ---- error behaviour of the reference: what its own code does, recorded as exception names
    def raised(fn):
        try:
            fn()
            return "none"
        except Exception as e:      # noqa: BLE001 (the type is the datum)
            return type(e).__name__
    far = np.full((1, 1, 3), 50.0, np.float32)
    out["err_empty_ball_query"] = np.array(raised(lambda: R.query_ball_point(0.2, 4, T(k3), T(far))))
    out["err_empty_ball_value"] = R.query_ball_point(0.2, 4, T(k3), T(far)).numpy()      # N in every slot
    out["err_empty_ball_gather"] = np.array(raised(
        lambda: R.index_points(T(k3), R.query_ball_point(0.2, 4, T(k3), T(far)))))
    out["err_nsample_gt_n"] = np.array(raised(lambda: R.query_ball_point(0.2, 7, T(k3), T(k3[:, :1]))))
    assert not paddle_stub._next_randint
    np.savez_compressed(os.path.join(HERE, "layers_ref.npz"), **out)
    for k, v in out.items():
        print(f"{k:24s} {str(v.dtype):8s} {v.shape}")
