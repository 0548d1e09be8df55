"""CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU) and exports every
symbol include/papc_b200.h declares; no compute call is made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from papc_b200.csrc import build as cuda_build
    return cuda_build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "papc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(papc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for s in ("papc_fps_f32", "papc_ball_query_f32", "papc_gather_f32", "papc_group_gather_f32",
              "papc_sa_mlp_f32", "papc_voxelize_f32", "papc_pfn_f32", "papc_pillar_scatter_f32",
              "papc_square_distance_f32"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built)
    for s in _declared_symbols():
        assert hasattr(handle, s), f"{s} declared in include/papc_b200.h but not exported"


def test_python_binding_covers_every_declared_symbol(built):
    from papc_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    lib = _lib.lib()
    assert lib.papc_abi_version() == 1
    assert lib.papc_status_string(-2).decode().startswith("PAPC_EWORKSPACE")


def test_pure_host_queries(built):
    """Size queries and argument validation run on the host only."""
    from papc_b200 import _lib
    lib = _lib.lib()
    assert lib.papc_fps_workspace_bytes(32, 1024) == 0
    assert lib.papc_fps_workspace_bytes(2, 10000) == 2 * 10000 * 4
    assert lib.papc_mlp_stats_partial_rows(524288) == 296
    assert lib.papc_mlp_stats_partial_rows(4096) == 32
    vs = (ctypes.c_float * 3)(0.16, 0.16, 4.0)
    cr = (ctypes.c_float * 6)(0, -39.68, -3, 69.12, 39.68, 1)
    assert lib.papc_voxelize_workspace_bytes(20000, vs, cr, 12000) >= 432 * 496 * 4
    # invalid arguments are rejected before any launch (no GPU needed)
    assert lib.papc_fps_f32(None, 1, 0, 1, None, 1.0, None, None, None, 0, None) == -1
    assert lib.papc_ball_query_f32(None, None, 1, 8, 1, 0.04, 16, None, 64, None, None) == -1
    assert lib.papc_voxelize_f32(None, 5, 2, vs, cr, 5, 1, 10, None, None, None, None, None, 0, None) == -1


def test_sa_mlp_workspace_plan_on_the_host(built):
    """papc_sa_mlp_workspace_bytes is pure host arithmetic: it covers the two activation buffers, the pooled
    extrema, TWO partial-row buffers (layers alternate, a deferred BatchNorm finalisation reads the previous
    layer's) and ONE statistic slot per layer, and rejects a malformed MLP."""
    from papc_b200 import _lib
    lib = _lib.lib()

    def ws_bytes(B, S, K, D, couts, chain_free=True):
        src = _lib.GroupSource()
        src.B, src.N, src.S, src.K, src.D, src.order = B, 1024, S, K, D, _lib.XYZ_FIRST
        src.xyz = 256   # non-null placeholders: the size query never dereferences them
        src.idx = 256
        if D:
            src.feats = 256
        mlp = _lib.Mlp()
        mlp.num_layers, mlp.cin, mlp.bn_mode, mlp.eps = len(couts), 3 + D, _lib.BN_BATCH, 1e-5
        for l, c in enumerate(couts):
            mlp.layers[l].cout = c
            mlp.layers[l].weight = 256
        return lib.papc_sa_mlp_workspace_bytes(ctypes.byref(src), ctypes.byref(mlp))

    M = 32 * 128 * 64
    b3 = ws_bytes(32, 128, 64, 128, [128, 128, 256])
    rows = lib.papc_mlp_stats_partial_rows(M)
    lower = 2 * M * 128 * 4 + 2 * (32 * 128) * 256 * 4 + 2 * rows * 2 * 256 * 8 + 3 * (4 * 256 + 8) * 8
    assert lower <= b3 < lower + (8 << 20)
    # one more hidden layer of the same width: exactly one more statistic slot (nothing else grows)
    b4 = ws_bytes(32, 128, 64, 128, [128, 128, 128, 256])
    assert abs((b4 - b3) - (4 * 256 + 8) * 8) < 256   # (regions are 256-byte aligned)
    assert ws_bytes(32, 128, 64, 128, []) == 0 and ws_bytes(32, 128, 64, 128, [128, 0]) == 0


def test_no_cpu_fallback():
    import torch
    from papc_b200 import _lib, layers
    with pytest.raises(_lib.PapcError):
        layers.farthest_point_sample(torch.zeros(1, 8, 3), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "papc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
