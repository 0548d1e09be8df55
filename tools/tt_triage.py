"""Performance triage of the transposed tcgen05 layer kernel: time papc_mlp_layer_forward_f32 on
synthetic rows for several M and PAPC_TT_DBG masks (1 = producers idle, 2 = no MMA, 4 = epilogue
idle) to separate fixed cost from per-tile cost per pipeline role."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from papc_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()
st = L.stream_ptr(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(M, cin, cout, K, pool, want_y, act=True, reps=5):
    x = torch.randn((M, cin), device=dev)
    w = torch.randn((cout, cin), device=dev) * (2.0 / cin) ** 0.5
    bias = torch.zeros(cout, device=dev)
    sc = torch.ones(cin, device=dev) if act else None
    sh = torch.zeros(cin, device=dev) if act else None
    y = torch.empty((M, cout), device=dev) if want_y else None
    pmax = torch.empty((M // K, cout), device=dev) if pool else None
    pmin = torch.empty((M // K, cout), device=dev) if pool else None
    partial = torch.empty((lib.papc_mlp_stats_partial_rows(M), 2, cout), dtype=torch.float64, device=dev)
    ws = torch.empty(max(lib.papc_mlp_layer_workspace_bytes(cin, cout), 256), dtype=torch.uint8, device=dev)

    def launch():
        L.check(lib.papc_mlp_layer_forward_f32(None, L.ptr(x), L.ptr(sc), L.ptr(sh), M, cin, cout, K, L.ptr(w),
                                               L.ptr(bias), L.ptr(y), L.ptr(pmax), L.ptr(pmin), L.ptr(partial),
                                               L.ptr(ws), ws.numel(), st), "layer")
    for _ in range(2):
        launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


cfgs = [("K64->128 pool", 64, 128, 32, True, False), ("K128->128 y", 128, 128, 64, False, True)]
masks = [int(m) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,7,5,6,3".split(","))]
tiles = [int(t) for t in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,8,28".split(","))]
for name, cin, cout, K, pool, want_y in cfgs:
    for tiles_per_cta in tiles:
        M = 128 * 148 * tiles_per_cta
        row = []
        for m in masks:
            os.environ["PAPC_TT_DBG"] = str(m)
            row.append(f"dbg{m}={run(M, cin, cout, K, pool, want_y):7.1f}")
        print(f"{name:16s} tiles/CTA={tiles_per_cta:3d}  " + "  ".join(row) + "  us", flush=True)
