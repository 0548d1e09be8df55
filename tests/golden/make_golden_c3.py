"""Generate tests/golden/c3_indices_ref.npz: the sampling / grouping indices of BASELINE config 3
(PointNet++ MSG segment, B = 16 clouds x 2048 points: sa1 = FPS 512 + ball queries r .1/.2/.4 K 32/64/128,
sa2 = FPS 128 + r .4/.8 K 64/128) from the reference's own farthest_point_sample / index_points /
query_ball_point executed over the NumPy stand-in for paddle.  The arrays (2.2 M indices) are stored as
sha256 digests of their int64 bytes plus cloud 0 in full (int16).  Inputs: synth.clouds(16, 2048, seed=0),
start indices seed 1 / zeros.  Build-container only:  python tests/golden/make_golden_c3.py"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import paddle_stub  # noqa: E402
from make_golden_layers import load_reference  # noqa: E402
from papc_b200 import synth  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a.astype(np.int64)).tobytes()).hexdigest()


if __name__ == "__main__":
    P, R = load_reference()
    T = P.to_tensor
    B, N = 16, 2048
    xyz = np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1))
    paddle_stub._next_randint.append(synth.fps_start(B, N, seed=1))
    fps1 = R.farthest_point_sample(T(xyz), 512)
    l1 = R.index_points(T(xyz), fps1)
    arrays = {"fps1": fps1.numpy()}
    for r, k in ((0.1, 32), (0.2, 64), (0.4, 128)):
        arrays[f"sa1_ball_r{r}_k{k}"] = R.query_ball_point(r, k, T(xyz), l1).numpy()
    paddle_stub._next_randint.append(np.zeros(B, np.int64))
    fps2 = R.farthest_point_sample(l1, 128)
    l2 = R.index_points(l1, fps2)
    arrays["fps2"] = fps2.numpy()
    for r, k in ((0.4, 64), (0.8, 128)):
        arrays[f"sa2_ball_r{r}_k{k}"] = R.query_ball_point(r, k, l1, l2).numpy()
    out = {}
    for k, v in arrays.items():
        out[k + ":sha256"] = np.array(digest(v))
        out[k + ":cloud0"] = v[0].astype(np.int16)
        print(k, v.shape, digest(v)[:16])
    np.savez_compressed(os.path.join(HERE, "c3_indices_ref.npz"), **out)
