"""GPU parity tests of the chained grouped-MLP kernels (papc_b200/csrc/sa_chain.cu, opt-in through PAPC_CHAIN=1):
the SetAbstraction layers on the chained path vs the oracle (<= 1e-5) and vs the layer-at-a-time path."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import layers_np  # noqa: E402
from papc_b200 import _lib as L  # noqa: E402
from papc_b200 import layers, synth  # noqa: E402

DEV = "cuda:0"


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _pair(cfg, seed):
    rng = np.random.default_rng(seed)
    g, r = layers.PointNetSetAbstraction(*cfg), layers_np.PointNetSetAbstraction(*cfg)
    for l, p in enumerate(synth.mlp_params(cfg[3], cfg[4], seed=seed)):
        w = p["weight"].reshape(*p["weight"].shape, 1, 1)
        gam = rng.uniform(0.5, 1.5, p["bias"].shape).astype(np.float32)
        gam[::6] *= -1.0     # negative BatchNorm weights: the pooled minimum is the one that survives
        bet = rng.uniform(-0.2, 0.2, p["bias"].shape).astype(np.float32)
        g.mlp_convs[l].weight, g.mlp_convs[l].bias = _cu(w), _cu(p["bias"])
        g.mlp_bns[l].weight, g.mlp_bns[l].bias = _cu(gam), _cu(bet)
        r.mlp_convs[l].weight, r.mlp_convs[l].bias = w, p["bias"]
        r.mlp_bns[l].weight, r.mlp_bns[l].bias = gam, bet
    return g.to(DEV), r


def _run_both(g, xyz, feats, start, monkeypatch):
    outs, kernels = {}, {}
    lib = L.lib()
    for flag in ("1", "0"):
        monkeypatch.setenv("PAPC_CHAIN", flag)
        L.check(lib.papc_prof_reset(), "reset")
        L.check(lib.papc_prof_enable(1), "enable")
        _, gp = g(_cu(xyz), _cu(feats) if feats is not None else None, start_idx=_cu(start))
        torch.cuda.synchronize()
        L.check(lib.papc_prof_enable(0), "disable")
        kernels[flag] = {r["name"] for r in L.prof_records()}
        L.check(lib.papc_prof_reset(), "reset")
        outs[flag] = gp.cpu().numpy()
    return outs, kernels


CASES = [
    # name, B, N, (npoint, radius, nsample, in_channel, mlp, group_all), D
    ("sa1-c2", 4, 1024, (512, 0.2, 32, 3, [64, 64, 128], False), 0),        # folded first layer, two accumulators
    ("sa2-c2", 4, 512, (128, 0.4, 64, 131, [128, 128, 256], False), 128),   # gathered image, stored tiles, 2 n-tiles
    ("ragged-tail", 3, 300, (50, 0.5, 32, 3, [32, 64, 64], False), 0),      # M = 4800: a 64-row tail tile
    ("msg-like-k128", 2, 1024, (64, 0.6, 128, 6, [64, 96, 128], False), 3),  # D = 3, one group per tile
    ("gather-small-c", 2, 512, (64, 0.5, 32, 19, [32, 32, 64], False), 16),  # narrow layers through the gather path
]


@pytest.mark.parametrize("name,B,N,cfg,D", CASES, ids=[c[0] for c in CASES])
def test_chained_mlp_vs_oracle_and_layer_path(name, B, N, cfg, D, monkeypatch):
    rng = np.random.default_rng(5)
    xyz = synth.clouds(B, N, seed=3)
    feats = np.maximum(rng.standard_normal((B, D, N)), 0).astype(np.float32) if D else None
    start = synth.fps_start(B, N, seed=4)
    g, r = _pair(cfg, seed=6)
    outs, kernels = _run_both(g, xyz, feats, start, monkeypatch)
    assert any(k.startswith("mlp_chain<") for k in kernels["1"]), kernels["1"]     # the chained kernels really ran
    assert not any(k.startswith("mlp_chain<") for k in kernels["0"])
    _, rp = r(xyz, feats, start_idx=start)
    assert np.isfinite(outs["1"]).all()
    np.testing.assert_allclose(outs["1"], rp, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(outs["1"], outs["0"], rtol=2e-5, atol=2e-5)


def test_chained_mlp_large_activations_column_scale(monkeypatch):
    """Inputs and BatchNorm weights far outside the fp16 range: the image's data-dependent column scale and the
    exact power-of-two scales of the hidden activations must keep the result finite and within tolerance."""
    B, N, D = 2, 512, 32
    cfg = (64, 0.5, 32, 3 + D, [64, 64, 128], False)
    rng = np.random.default_rng(9)
    xyz = synth.clouds(B, N, seed=1)
    feats = (np.maximum(rng.standard_normal((B, D, N)), 0) * 3.0e5).astype(np.float32)   # >> 65504
    start = synth.fps_start(B, N, seed=2)
    g, r = _pair(cfg, seed=8)
    for l in range(2):
        gam = (r.mlp_bns[l].weight * 3000.0).astype(np.float32)
        r.mlp_bns[l].weight = gam
        g.mlp_bns[l].weight = _cu(gam)
    outs, kernels = _run_both(g, xyz, feats, start, monkeypatch)
    assert any(k.startswith("mlp_chain<") for k in kernels["1"])
    _, rp = r(xyz, feats, start_idx=start)
    assert np.isfinite(outs["1"]).all()
    np.testing.assert_allclose(outs["1"], rp, rtol=1e-5, atol=1e-5)


def test_chain_is_deterministic(monkeypatch):
    monkeypatch.setenv("PAPC_CHAIN", "1")
    B, N = 4, 1024
    g, _ = _pair((512, 0.2, 32, 3, [64, 64, 128], False), seed=3)
    xyz, start = _cu(synth.clouds(B, N, seed=0)), _cu(synth.fps_start(B, N, seed=1))
    a = g(xyz, None, start_idx=start)[1].clone()
    b = g(xyz, None, start_idx=start)[1].clone()
    assert torch.equal(a, b)
