"""Generate tests/golden/c2_indices_ref.npz: the sampling / grouping INDICES of BASELINE config 2
(PointNet++ SSG, B = 32 clouds x 1024 points: sa1 = FPS 512 + ball query r 0.2 K 32, sa2 = FPS 128 + ball
query r 0.4 K 64) produced by EXECUTING THE REFERENCE'S OWN farthest_point_sample / index_points /
query_ball_point (pointnet2_basic_layers.py, unmodified) over the NumPy stand-in for paddle.  Inputs are
the bench's seeded clouds (synth.clouds(32, 1024, seed=0), start indices seed 1 / zeros), so only the
indices are stored (int16).  Build-container only:  python tests/golden/make_golden_c2.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import paddle_stub  # noqa: E402
from make_golden_layers import load_reference  # noqa: E402
from papc_b200 import synth  # noqa: E402

if __name__ == "__main__":
    P, R = load_reference()
    T = P.to_tensor
    B, N = 32, 1024
    xyz = np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1))
    st1, st2 = synth.fps_start(B, N, seed=1), np.zeros(B, np.int64)
    paddle_stub._next_randint.append(st1)
    fps1 = R.farthest_point_sample(T(xyz), 512)
    l1 = R.index_points(T(xyz), fps1)
    ball1 = R.query_ball_point(0.2, 32, T(xyz), l1).numpy()
    paddle_stub._next_randint.append(st2)
    fps2 = R.farthest_point_sample(l1, 128)
    l2 = R.index_points(l1, fps2)
    ball2 = R.query_ball_point(0.4, 64, l1, l2).numpy()
    out = dict(fps1=fps1.numpy().astype(np.int16), ball1=ball1.astype(np.int16),
               fps2=fps2.numpy().astype(np.int16), ball2=ball2.astype(np.int16))
    np.savez_compressed(os.path.join(HERE, "c2_indices_ref.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, int(v.min()), int(v.max()))
