"""Time the index-producing primitives of the bench workload (FPS, ball query) with CUDA events.
usage: python tools/prof_prims.py            (PAPC_FPS_WIDE=1 selects the wide FPS variant)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from papc_b200 import _lib as L  # noqa: E402
from papc_b200 import layers, synth  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("PAPC_PROF_B", "32"))


def timeit(fn, reps=20, warm=3):
    """Kernel-only time (the library's launch profiler: CUDA events around the launch itself)."""
    lib = L.lib()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    lib.papc_prof_reset()
    lib.papc_prof_enable(1)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    lib.papc_prof_enable(0)
    ts = [r["ms"] * 1e3 for r in L.prof_records()]
    lib.papc_prof_reset()
    return float(np.median(ts))


for (N, S, r, K) in [(1024, 512, 0.2, 32), (512, 128, 0.4, 64), (2048, 512, 0.2, 64)]:
    xyz = torch.from_numpy(np.ascontiguousarray(synth.clouds(B, N, seed=0).transpose(0, 2, 1))).to(dev)
    st = torch.zeros(B, dtype=torch.int64, device=dev)
    t_fps = timeit(lambda: layers.farthest_point_sample_idx(xyz, S, st, True))
    _, new_xyz = layers.farthest_point_sample_idx(xyz, S, st, True)
    t_bq = timeit(lambda: layers._ball_query(r, K, xyz, new_xyz, torch.int32))
    print(f"B={B} N={N} S={S}: fps {t_fps:8.1f} us ({t_fps * 1e3 / S:6.1f} ns/iter)   "
          f"ball_query(r={r},K={K}) {t_bq:7.1f} us", flush=True)
