"""chain_check.py -- chained MLP kernels (sa_chain.cu) vs the layer-at-a-time path (PAPC_CHAIN=0) vs the
oracle, on the BASELINE config-2 SetAbstraction shapes.  Development aid; run on a GPU box:
    python tools/chain_check.py [B] [which: sa1 sa2 msg]
Prints the max abs difference per comparison and per-kernel times from the launch profiler."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import layers_np
from papc_b200 import _lib as L
from papc_b200 import layers, synth

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def set_params(g_convs, g_bns, r_convs, r_bns, params, rng):
    for l, p in enumerate(params):
        w = p["weight"].reshape(*p["weight"].shape, 1, 1)
        g_convs[l].weight, g_convs[l].bias = cu(w), cu(p["bias"])
        r_convs[l].weight, r_convs[l].bias = w, p["bias"]
        gam = rng.uniform(0.5, 1.5, p["bias"].shape).astype(np.float32)
        bet = rng.uniform(-0.2, 0.2, p["bias"].shape).astype(np.float32)
        g_bns[l].weight, g_bns[l].bias = cu(gam), cu(bet)
        r_bns[l].weight, r_bns[l].bias = gam, bet


def run_case(name, B, N, cfg, D, with_oracle=True):
    rng = np.random.default_rng(7)
    xyz = synth.clouds(B, N, seed=3)
    feats = rng.standard_normal((B, D, N)).astype(np.float32) if D else None
    if feats is not None:
        feats = np.maximum(feats, 0.0)  # like relu(bn(.)) outputs of the previous layer
    start = synth.fps_start(B, N, seed=4)
    g, r = layers.PointNetSetAbstraction(*cfg), layers_np.PointNetSetAbstraction(*cfg)
    set_params(g.mlp_convs, g.mlp_bns, r.mlp_convs, r.mlp_bns, synth.mlp_params(cfg[3], cfg[4], seed=5), rng)
    g.to(DEV)
    outs = {}
    for flag in ("1", "0"):
        os.environ["PAPC_CHAIN"] = flag
        try:
            gx, gp = g(cu(xyz), cu(feats) if D else None, start_idx=cu(start))
            torch.cuda.synchronize()
            outs[flag] = gp.cpu().numpy()
        except Exception as e:  # noqa: BLE001
            print(f"[{name}] PAPC_CHAIN={flag}: FAILED {type(e).__name__}: {e}")
            return False
    d = float(np.abs(outs["1"] - outs["0"]).max())
    print(f"[{name}] B={B} chain vs layer path: max abs diff {d:.3e}  (finite: {np.isfinite(outs['1']).all()})")
    ok = d < 2e-5
    if with_oracle:
        t0 = time.time()
        rx, rp = r(xyz, feats, start_idx=start)
        e1 = float(np.abs(outs["1"] - rp).max())
        e0 = float(np.abs(outs["0"] - rp).max())
        print(f"[{name}]   vs oracle: chain {e1:.3e}   layer path {e0:.3e}   (oracle {time.time() - t0:.1f}s)")
        ok = ok and e1 < 1e-5
        if e1 >= 1e-5:
            bad = np.argwhere(np.abs(outs["1"] - rp) >= 1e-5)
            print(f"[{name}]   {len(bad)} elements off; first {bad[:5].tolist()}; channels {sorted(set(bad[:, 1].tolist()))[:20]}")
    return ok


def profile(name, B, N, cfg, D, reps=5):
    rng = np.random.default_rng(7)
    xyz = cu(synth.clouds(B, N, seed=3))
    feats = cu(np.maximum(rng.standard_normal((B, D, N)), 0).astype(np.float32)) if D else None
    start = cu(synth.fps_start(B, N, seed=4))
    g = layers.PointNetSetAbstraction(*cfg)
    g.to(DEV)
    lib = L.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for flag in ("1", "0"):
        os.environ["PAPC_CHAIN"] = flag
        for _ in range(3):
            g(xyz, feats, start_idx=start)
        torch.cuda.synchronize()
        L.check(lib.papc_prof_reset(), "r")
        L.check(lib.papc_prof_enable(1), "e")
        for _ in range(reps):
            flush.zero_()
            g(xyz, feats, start_idx=start)
        torch.cuda.synchronize()
        L.check(lib.papc_prof_enable(0), "e")
        recs = L.prof_records()
        L.check(lib.papc_prof_reset(), "r")
        agg = {}
        for rr in recs:
            k = (rr["name"], rr["M"], rr["cin"], rr["cout"])
            agg.setdefault(k, []).append(rr["ms"])
        tot = 0.0
        print(f"[{name}] PAPC_CHAIN={flag} per-kernel (B={B}):")
        for k, v in agg.items():
            ms = float(np.mean(v)) * len(v) / reps
            tot += ms
            print(f"    {k[0]:44s} M={k[1]:8d} {k[2]:4d}->{k[3]:4d}  {1e3 * float(np.mean(v)):8.1f} us x{len(v) // reps}")
        print(f"    total {1e3 * tot:.1f} us")


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    which = sys.argv[2:] or ["sa1", "sa2"]
    SA1 = (512, 0.2, 32, 3, [64, 64, 128], False)
    SA2 = (128, 0.4, 64, 131, [128, 128, 256], False)
    ok = True
    if "sa1" in which:
        ok &= run_case("sa1", B, 1024, SA1, 0)
    if "sa2" in which:
        ok &= run_case("sa2", B, 512, SA2, 128)
    if "msg" in which:
        ok &= run_case("msg-b0", B, 1024, (256, 0.1, 32, 6, [32, 32, 64], False), 3)
        ok &= run_case("msg-b2", B, 1024, (256, 0.4, 128, 6, [64, 96, 128], False), 3)
    if "prof" in which:
        profile("sa1", 32, 1024, SA1, 0)
        profile("sa2", 32, 512, SA2, 128)
    print("CHAIN_CHECK", "OK" if ok else "FAILED")
