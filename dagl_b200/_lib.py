"""ctypes binding of include/dagl_b200.h.  No CPU fallback: if the library is
missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DAGL_B200_LIB") or os.path.join(_HERE, "libdagl_b200.so")   # env override: A/B builds in development

IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_TC4 = 0, 1, 2, 4
IMPL_BY_NAME = {"auto": IMPL_AUTO, "simt": IMPL_SIMT, "tc": IMPL_TC, "tc4": IMPL_TC4}

EXPORTS = [
    "dagl_abi_version", "dagl_last_error", "dagl_ce_workspace_bytes", "dagl_ce_forward_f32",
    "dagl_ce_forward_debug_f32", "dagl_ce_host_staging_bytes", "dagl_ce_forward_host_f32",
    "dagl_graph_attend_workspace_bytes", "dagl_graph_attend_f32", "dagl_ce_workspace_view",
    "dagl_last_impl", "dagl_last_launch_count", "dagl_profile_enable", "dagl_profile_read",
    "dagl_ce_num_query_tiles", "dagl_ce_forward_rows_f32", "dagl_ce_fold_rows_f32",
    "dagl_ces_heads_forward_f32", "dagl_ce_packed_weights_bytes", "dagl_ce_pack_weights_f32",
    "dagl_ces_workspace_bytes", "dagl_ce_rows_workspace_bytes",
    "dagl_graph_attend_backward_workspace_bytes", "dagl_graph_attend_backward_f32", "dagl_ce_workspace_bytes_ex",
    "dagl_resblock_packed_weights_bytes", "dagl_resblock_pack_weights_f32", "dagl_resblocks_workspace_bytes",
    "dagl_resblocks_forward_f32",
]


class DaglCEWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "g_w", "g_b", "theta_w", "theta_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
        "thr_w", "thr_b", "bias_w", "bias_b")] + [
        ("in_channels", C.c_int32), ("inter_channels", C.c_int32), ("ksize", C.c_int32),
        ("stride_q", C.c_int32), ("stride_k", C.c_int32), ("softmax_scale", C.c_float),
        ("packed_fc", C.c_void_p), ("legacy_topk", C.c_int32)]


class DaglResBlockWeights(C.Structure):
    _fields_ = [("conv1_w", C.c_void_p), ("conv1_b", C.c_void_p), ("prelu_w", C.c_void_p), ("prelu_n", C.c_int32),
                ("conv2_w", C.c_void_p), ("conv2_b", C.c_void_p), ("res_scale", C.c_float), ("packed", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    """Load libdagl_b200.so (built in-tree by ``python -m dagl_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m dagl_b200.build` "
            "(there is deliberately no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    L.dagl_abi_version.restype = i32
    L.dagl_last_error.restype = C.c_char_p
    L.dagl_last_impl.restype = C.c_char_p
    L.dagl_last_launch_count.restype = i32
    L.dagl_ce_workspace_bytes.restype = sz
    L.dagl_ce_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.dagl_ce_workspace_bytes_ex.restype = sz
    L.dagl_ce_workspace_bytes_ex.argtypes = [i32, i32, i32, i32, i32, i32]
    L.dagl_ce_host_staging_bytes.restype = sz
    L.dagl_ce_host_staging_bytes.argtypes = [i32, i32, i32, i32]
    L.dagl_ce_forward_f32.restype = i32
    L.dagl_ce_forward_f32.argtypes = [C.POINTER(DaglCEWeights), vp, vp, i32, i32, i32, vp, sz, i32, vp]
    L.dagl_ce_packed_weights_bytes.restype = sz
    L.dagl_ce_packed_weights_bytes.argtypes = []
    L.dagl_ce_pack_weights_f32.restype = i32
    L.dagl_ce_pack_weights_f32.argtypes = [C.POINTER(DaglCEWeights), vp, sz, vp]
    L.dagl_ces_heads_forward_f32.restype = i32
    L.dagl_ces_heads_forward_f32.argtypes = [C.POINTER(C.POINTER(DaglCEWeights)), i32, vp, vp, i32, i32, i32, vp, sz, i32, vp]
    L.dagl_ce_forward_debug_f32.restype = i32
    L.dagl_ce_forward_debug_f32.argtypes = [C.POINTER(DaglCEWeights), vp, vp, i32, i32, i32, vp, sz, i32, vp, vp, vp]
    L.dagl_ce_forward_host_f32.restype = i32
    L.dagl_ce_forward_host_f32.argtypes = [C.POINTER(DaglCEWeights), vp, vp, i32, i32, i32, vp, sz, i32, vp]
    L.dagl_graph_attend_workspace_bytes.restype = sz
    L.dagl_graph_attend_workspace_bytes.argtypes = [i32, i32, i32]
    L.dagl_graph_attend_f32.restype = i32
    L.dagl_graph_attend_f32.argtypes = [vp] * 7 + [i32, i32, i32, C.c_float, vp, sz, i32, vp, vp, vp]
    L.dagl_ce_workspace_view.restype = vp
    L.dagl_ce_workspace_view.argtypes = [vp, i32, i32, i32, i32, i32]
    L.dagl_ce_num_query_tiles.restype = i32
    L.dagl_ce_num_query_tiles.argtypes = [i32, i32]
    L.dagl_ce_forward_rows_f32.restype = i32
    L.dagl_ce_forward_rows_f32.argtypes = [C.POINTER(DaglCEWeights), vp, vp, i32, i32, i32, i32, i32, vp, sz, i32, vp]
    L.dagl_ce_rows_workspace_bytes.restype = sz
    L.dagl_ce_rows_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i32]
    L.dagl_ces_workspace_bytes.restype = sz
    L.dagl_ces_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.dagl_ce_fold_rows_f32.restype = i32
    L.dagl_ce_fold_rows_f32.argtypes = [vp, vp, i32, i32, i32, vp]
    L.dagl_graph_attend_backward_workspace_bytes.restype = sz
    L.dagl_graph_attend_backward_workspace_bytes.argtypes = [i32, i32, i32]
    L.dagl_graph_attend_backward_f32.restype = i32
    L.dagl_graph_attend_backward_f32.argtypes = [vp] * 11 + [i32, i32, i32, C.c_float, vp, sz, vp]
    L.dagl_resblock_packed_weights_bytes.restype = sz
    L.dagl_resblock_packed_weights_bytes.argtypes = []
    L.dagl_resblock_pack_weights_f32.restype = i32
    L.dagl_resblock_pack_weights_f32.argtypes = [C.POINTER(DaglResBlockWeights), vp, sz, vp]
    L.dagl_resblocks_workspace_bytes.restype = sz
    L.dagl_resblocks_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.dagl_resblocks_forward_f32.restype = i32
    L.dagl_resblocks_forward_f32.argtypes = [C.POINTER(DaglResBlockWeights), i32, vp, vp, i32, i32, i32, i32, vp, sz, i32, vp]
    L.dagl_profile_enable.restype = i32
    L.dagl_profile_enable.argtypes = [i32]
    L.dagl_profile_read.restype = i32
    L.dagl_profile_read.argtypes = [C.POINTER(C.c_float), i32]
    if L.dagl_abi_version() != 5:
        raise RuntimeError("libdagl_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().dagl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
