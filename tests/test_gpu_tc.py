"""GPU tests of the tcgen05 (3xTF32) layer kernel through the step-wise C ABI
(papc_mlp_layer_forward_f32), against fp64 NumPy and against the fp32 SIMT kernel."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from papc_b200 import _lib as L  # noqa: E402
from papc_b200 import layers  # noqa: E402

DEV = "cuda:0"


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _run_layer(x, w, bias, scale, shift, K, want_y, want_pool, tc, src=None, M=None):
    lib = L.lib()
    # tc: "tt" (default: transposed kernel, weights in tensor memory), True / "tc" (shared/shared
    # tcgen05 kernel), False / "simt" (fp32 CUDA-core kernel)
    mode = {True: "tc", False: "simt"}.get(tc, tc)
    os.environ.pop("PAPC_MLP_TC", None)
    if mode != "tt":
        os.environ["PAPC_MLP_TC"] = "1" if mode == "tc" else "0"
    try:
        cout, cin = w.shape
        M = x.shape[0] if M is None else M
        y = torch.full((M, cout), float("nan"), device=DEV) if want_y else None
        G = M // K
        pmax = torch.full((G, cout), float("nan"), device=DEV) if want_pool else None
        pmin = torch.full((G, cout), float("nan"), device=DEV) if want_pool else None
        rows = lib.papc_mlp_stats_partial_rows(M)
        partial = torch.full((rows, 2, cout), float("nan"), dtype=torch.float64, device=DEV)
        wsb = lib.papc_mlp_layer_workspace_bytes(cin, cout)
        ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=DEV)
        wd, bd = _cu(w), (_cu(bias) if bias is not None else None)
        xd = _cu(x) if x is not None else None
        sc, sh = (_cu(scale), _cu(shift)) if scale is not None else (None, None)
        L.check(lib.papc_mlp_layer_forward_f32(C.byref(src) if src is not None else None, L.ptr(xd), L.ptr(sc),
                                               L.ptr(sh), M, cin, cout, K, L.ptr(wd), L.ptr(bd), L.ptr(y),
                                               L.ptr(pmax), L.ptr(pmin), L.ptr(partial), L.ptr(ws), wsb,
                                               L.stream_ptr(torch.device(DEV))), "layer")
        torch.cuda.synchronize()
        sums = partial.sum(0).cpu().numpy()
        return (y.cpu().numpy() if want_y else None, pmax.cpu().numpy() if want_pool else None,
                pmin.cpu().numpy() if want_pool else None, sums)
    finally:
        os.environ.pop("PAPC_MLP_TC", None)


def _describe(got, ref, name):
    err = np.abs(got - ref)
    bad = err > 1e-4 * (1 + np.abs(ref))
    msg = [f"{name}: max abs err {np.nanmax(err):.3e}, bad {bad.mean() * 100:.2f}% nan {np.isnan(got).mean() * 100:.2f}%"]
    if bad.any():
        r, c = np.nonzero(bad)
        msg.append(f"bad rows (first 16 distinct): {sorted(set(r.tolist()))[:16]} ... cols: {sorted(set(c.tolist()))[:16]}")
        msg.append(f"bad by row%8: {np.bincount(r % 8, minlength=8).tolist()}  by col%8: {np.bincount(c % 8, minlength=8).tolist()}")
        msg.append(f"bad by row//32 (first 8): {np.bincount(r // 32)[:8].tolist()}  by col//32: {np.bincount(c // 32).tolist()}")
        msg.append(f"got[0,:8]={got[0, :8]}  ref[0,:8]={ref[0, :8]}")
        msg.append(f"got[1,:8]={got[1, :8]}  ref[1,:8]={ref[1, :8]}")
    return "\n".join(msg)


def test_tc_identity_layout_probe():
    """W = I: y must reproduce act(x) exactly -- any descriptor / swizzle bug shows up as a permutation."""
    M, Cn = 256, 64
    x = (np.arange(M)[:, None] * 100.0 + np.arange(Cn)[None, :]).astype(np.float32) / 8.0
    w = np.eye(Cn, dtype=np.float32)
    for mode in ("tt", "tc"):
        y, _, _, _ = _run_layer(x, w, None, None, None, 32, True, False, tc=mode)
        assert np.array_equal(y, x), _describe(y, x, "identity " + mode)


@pytest.mark.parametrize("M,cin,cout,K", [(1024, 64, 64, 32), (4096, 64, 128, 32), (2048, 128, 128, 64),
                                          (4096, 128, 256, 64), (640, 96, 48, 32), (1000, 32, 200, 8),
                                          (128 * 300 + 77, 64, 64, 1), (128 * 700, 128, 128, 128),
                                          (96, 8, 16, 32), (128 * 149 + 32, 72, 300, 32),
                                          (1024, 256, 128, 64), (640, 512, 1024, 128), (2048, 192, 160, 32)])
def test_tc_plain_layer_vs_fp64(M, cin, cout, K):
    rng = np.random.default_rng(M + cin)
    x = rng.standard_normal((M, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin)) * np.sqrt(2.0 / cin)).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cin).astype(np.float32)
    shift = rng.uniform(-0.5, 0.5, cin).astype(np.float32)
    act = np.maximum(x.astype(np.float64) * scale + shift, 0.0).astype(np.float32).astype(np.float64)
    ref = act @ w.T.astype(np.float64) + bias
    pool = (M % K == 0) and K in (32, 64, 128)
    for tc in ("tt", "tc", "simt"):
        y, pmax, pmin, sums = _run_layer(x, w, bias, scale, shift, K, True, pool, tc=tc)
        name = tc
        assert np.allclose(y, ref, rtol=2e-6, atol=2e-6 * np.sqrt(cin)), _describe(y, ref, name)
        # the tensor-core fp32 accumulator truncates (a ~1e-7 relative bias toward zero per element),
        # so the statistic sums are checked against sum|y|, not against the (cancelling) sum itself
        np.testing.assert_allclose(sums[0], ref.sum(0), rtol=0,
                                   atol=2e-6 * max(1.0, cin / 128) * np.abs(ref).sum(0).max(), err_msg=name)
        np.testing.assert_allclose(sums[1], (ref ** 2).sum(0), rtol=4e-6 * max(1.0, cin / 128), atol=1e-3, err_msg=name)
        if pool:
            g = y.reshape(M // K, K, cout)  # pooled extrema must be exactly those of the stored y
            assert np.array_equal(pmax, g.max(1)) and np.array_equal(pmin, g.min(1)), name


@pytest.mark.parametrize("D,order", [(0, 0), (128, 0), (64, 1), (8, 1)])
def test_tc_gather_layer_vs_simt(D, order):
    B, N, S, K, cout = 4, 256, 64, 32, 128
    rng = np.random.default_rng(D)
    xyz = rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)
    new_xyz = xyz[:, :S].copy()
    feats = rng.standard_normal((B, N, D)).astype(np.float32) if D else None
    idx = rng.integers(0, N, (B, S, K)).astype(np.int32)
    w = (rng.standard_normal((cout, 3 + D)) * np.sqrt(2.0 / (3 + D))).astype(np.float32)
    keep = [_cu(xyz), _cu(new_xyz), _cu(feats) if D else None, _cu(idx)]
    src = layers._make_src(keep[0], keep[1], keep[2], keep[3], B, N, S, K, order)
    M = B * S * K
    outs = {}
    for tc in ("tt", "tc", "simt"):
        outs[tc] = _run_layer(None, w, None, None, None, K, True, True, tc=tc, src=src, M=M)
    # fp64 reference of the gathered rows
    g_xyz = np.stack([xyz[b][idx[b].reshape(-1)] for b in range(B)]).reshape(B, S, K, 3) - new_xyz[:, :, None]
    if D:
        g_f = np.stack([feats[b][idx[b].reshape(-1)] for b in range(B)]).reshape(B, S, K, D)
        rows = np.concatenate([g_xyz, g_f] if order == 0 else [g_f, g_xyz], -1)
    else:
        rows = g_xyz
    ref = rows.reshape(M, 3 + D).astype(np.float64) @ w.T.astype(np.float64)
    for tc in ("tt", "tc", "simt"):
        y = outs[tc][0]
        assert np.allclose(y, ref, rtol=2e-6, atol=2e-5), _describe(y, ref, tc)
        g = y.reshape(M // K, K, cout)
        assert np.array_equal(outs[tc][1], g.max(1)) and np.array_equal(outs[tc][2], g.min(1)), tc
