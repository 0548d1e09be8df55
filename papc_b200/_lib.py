"""ctypes binding of the C ABI in include/papc_b200.h (papc_b200/lib/libpapc_b200.so).

This is the ONLY compute back end: there is no CPU or PyTorch fallback.  If the shared
library is missing the import of any op fails with a clear error; build it with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``python papc_b200/csrc/build.py``).
PyTorch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAPC_B200_LIB: development override (e.g. the triage build of tools/build_triage.sh)
LIB_PATH = os.environ.get("PAPC_B200_LIB") or os.path.join(_HERE, "lib", "libpapc_b200.so")

PAPC_OK = 0
XYZ_FIRST, FEATS_FIRST = 0, 1
BN_BATCH, BN_RUNNING, BN_NONE = 0, 1, 2
OUT_BSC, OUT_BCS = 0, 1
MAX_MLP_LAYERS = 8

_f32p = C.c_void_p
_vp = C.c_void_p


class GroupSource(C.Structure):
    _fields_ = [("grouped", _vp), ("xyz", _vp), ("new_xyz", _vp), ("feats", _vp), ("idx", _vp),
                ("B", C.c_int32), ("N", C.c_int32), ("S", C.c_int32), ("K", C.c_int32),
                ("D", C.c_int32), ("order", C.c_int32),
                ("xyz_moments", _vp), ("xyz_moment_rows", C.c_int32), ("reserved", C.c_int32),
                ("feats_colscale", _vp)]


class MlpLayer(C.Structure):
    _fields_ = [("weight", _vp), ("bias", _vp), ("gamma", _vp), ("beta", _vp),
                ("running_mean", _vp), ("running_var", _vp), ("batch_mean", _vp),
                ("batch_var", _vp), ("cout", C.c_int32), ("reserved", C.c_int32)]


class Mlp(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("cin", C.c_int32), ("bn_mode", C.c_int32),
                ("eps", C.c_float), ("layers", MlpLayer * MAX_MLP_LAYERS)]


class PapcError(RuntimeError):
    pass


_LIB = None

# name -> (restype, argtypes); mirrors include/papc_b200.h one to one
_I, _I64, _F, _D, _SZ = C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_SIGNATURES = {
    "papc_status_string": (C.c_char_p, [_I]),
    "papc_abi_version": (_I, []),
    "papc_last_cuda_error": (_I, []),
    "papc_launch_count": (C.c_uint64, []),
    "papc_prof_enable": (_I, [_I]),
    "papc_prof_reset": (_I, []),
    "papc_prof_count": (_I, []),
    "papc_prof_get": (_I, [_I, C.c_char_p, _I, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                           C.POINTER(C.c_int32), C.POINTER(_D), C.POINTER(_D), C.POINTER(_F)]),
    "papc_square_distance_f32": (_I, [_vp, _vp, _I, _I, _I, _vp, _vp]),
    "papc_gather_f32": (_I, [_vp, _vp, _I, _I, _I, _I, _vp, _vp]),
    "papc_fps_workspace_bytes": (_SZ, [_I, _I]),
    "papc_fps_f32": (_I, [_vp, _I, _I, _I, _vp, _F, _vp, _vp, _vp, _SZ, _vp]),
    "papc_sample_group_parts": (_I, [_I, _I, _I, _I]),
    "papc_sample_group_workspace_bytes": (_SZ, [_I, _I]),
    "papc_sample_group_f32": (_I, [_vp, _I, _I, _I, _vp, _F, _F, _I, _vp, _vp, _vp, _vp, _vp, _vp, _SZ, _vp]),
    "papc_ball_query_f32": (_I, [_vp, _vp, _I, _I, _I, _F, _I, _vp, _I, _vp, _vp]),
    "papc_ball_query_multi_f32": (_I, [_vp, _vp, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_void_p), _I, _vp, _vp]),
    "papc_knn_f32": (_I, [_vp, _vp, _I, _I, _I, _I, _vp, _I, _vp, _vp]),
    "papc_group_gather_f32": (_I, [_vp, _vp, _vp, _vp, _I, _I, _I, _I, _I, _I, _vp, _vp]),
    "papc_sa_mlp_workspace_bytes": (_SZ, [C.POINTER(GroupSource), C.POINTER(Mlp)]),
    "papc_sa_mlp_f32": (_I, [C.POINTER(GroupSource), C.POINTER(Mlp), _vp, _I, _vp, _SZ, _vp]),
    "papc_mlp_stats_partial_rows": (_I64, [_I64]),
    "papc_mlp_layer_workspace_bytes": (_SZ, [C.c_int32, C.c_int32]),
    "papc_mlp_layer_forward_f32": (_I, [C.POINTER(GroupSource), _vp, _vp, _vp, _I64, C.c_int32,
                                         C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _SZ,
                                         _vp]),
    "papc_mlp_stats_reduce_f64": (_I, [_vp, _I64, C.c_int32, _vp, _vp]),
    "papc_bn_scale_shift_f32": (_I, [_vp, _D, _vp, _vp, _F, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "papc_bn_running_scale_shift_f32": (_I, [_vp, _vp, _vp, _vp, _F, C.c_int32, _vp, _vp, _vp]),
    "papc_sa_pool_finish_f32": (_I, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _I,
                                      _vp]),
    "papc_bn_relu_apply_f32": (_I, [_vp, _vp, _vp, _I64, C.c_int32, _vp, _vp]),
    "papc_fp_interpolate_f32": (_I, [_vp, _vp, _vp, _vp, _I, _I, _I, _I, _I, _I, _vp, _vp]),
    "papc_pointwise_mlp_workspace_bytes": (_SZ, [_I64, C.c_int32, C.POINTER(Mlp)]),
    "papc_pointwise_mlp_f32": (_I, [_vp, _I64, C.c_int32, C.POINTER(Mlp), _vp, _vp, _SZ, _vp]),
    "papc_voxelize_workspace_bytes": (_SZ, [_I, C.POINTER(_F), C.POINTER(_F), _I]),
    "papc_voxelize_f32": (_I, [_vp, _I, _I, C.POINTER(_F), C.POINTER(_F), _I, _I, _I, _vp, _vp, _vp,
                                _vp, _vp, _SZ, _vp]),
    "papc_pfn_workspace_bytes": (_SZ, [_I, _I]),
    "papc_pfn_f32": (_I, [_vp, _vp, _vp, _I, _I, _I, _F, _F, _F, _F, _vp, _vp, _vp, _vp, _vp, _vp, _I,
                           _F, _I, _vp, _vp, _vp, _vp, _vp, _SZ, _vp]),
    "papc_pfn_layer_workspace_bytes": (_SZ, [_I, _I, _I]),
    "papc_pfn_layer_f32": (_I, [_vp, _I, _I, _vp, _vp, _I, _I, _I, _F, _F, _F, _F, _vp, _vp, _vp, _vp, _vp, _vp, _I,
                                 _F, _I, _I, _vp, _vp, _vp, _vp, _vp, _SZ, _vp]),
    "papc_pillar_scatter_workspace_bytes": (_SZ, [_I, _I, _I]),
    "papc_pillar_scatter_f32": (_I, [_vp, _vp, _I, _I, _I, _I, _I, _vp, _vp, _vp, _SZ, _vp]),
    "papc_voxelize_batch_workspace_bytes": (_SZ, [_I, _I, C.POINTER(_F), C.POINTER(_F), _I]),
    "papc_voxelize_batch_f32": (_I, [_vp, C.POINTER(C.c_int32), _I, _I, C.POINTER(_F), C.POINTER(_F), _I, _I, _I,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _SZ, _vp]),
    "papc_anchors_mask_f32": (_I, [_vp, _I, _I, _vp, _I, _I, _vp, _I, C.POINTER(_F), C.POINTER(_F),
                                    C.POINTER(C.c_int32), _vp, _vp, _vp]),
    "papc_points_to_bev_workspace_bytes": (_SZ, [_I, C.POINTER(_F), C.POINTER(_F)]),
    "papc_points_to_bev_f32": (_I, [_vp, _I, _I, C.POINTER(_F), C.POINTER(_F), C.POINTER(_F), _I, _I, _vp, _vp,
                                     _SZ, _vp]),
    "papc_p2p_allgather_f32": (_I, [_vp, _I64, C.POINTER(C.c_void_p), _I, _I, _vp]),
    "papc_nms_workspace_bytes": (_SZ, [_I]),
    "papc_nms_f32": (_I, [_vp, _I, _I, _F, _vp, _vp, _vp, _SZ, _vp]),
    "papc_rotate_iou_f32": (_I, [_vp, _I, _vp, _I, _I, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load (once) and return the C-ABI library; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PapcError(
                f"{LIB_PATH} not found: the sm_100a CUDA library is the only back end of papc_b200 "
                "(no CPU fallback).  Build it: python -c 'import __graft_entry__ as g; g.build()'")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.papc_abi_version() != 1:
            raise PapcError("libpapc_b200.so ABI version mismatch")
        _LIB = handle
    return _LIB


def check(status: int, what: str):
    if status != PAPC_OK:
        l = lib()
        msg = l.papc_status_string(status).decode()
        if status == -3:
            msg += f" (cudaError {l.papc_last_cuda_error()})"
        raise PapcError(f"{what}: {msg}")


def prof_records():
    """All records of the launch profiler as dicts (synchronises their events)."""
    l = lib()
    out = []
    name = C.create_string_buffer(64)
    M, ci, co = C.c_int64(), C.c_int32(), C.c_int32()
    fl, by, ms = C.c_double(), C.c_double(), C.c_float()
    for i in range(l.papc_prof_count()):
        check(l.papc_prof_get(i, name, 64, C.byref(M), C.byref(ci), C.byref(co), C.byref(fl), C.byref(by),
                              C.byref(ms)), "prof_get")
        out.append({"name": name.value.decode(), "M": M.value, "cin": ci.value, "cout": co.value,
                    "flops": fl.value, "bytes": by.value, "ms": ms.value})
    return out


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PapcError("papc_b200 ops need CUDA tensors: there is no CPU fallback "
                            "(the CPU oracle lives in oracle/ and is test infrastructure only)")


def f32c(t):
    """Contiguous float32 view/copy of a CUDA tensor."""
    if t is None:
        return None
    if type(t) is not torch.Tensor:   # e.g. the paddle facade's Tensor subclass (papc_b200.compat): plain view
        t = t.as_subclass(torch.Tensor)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
