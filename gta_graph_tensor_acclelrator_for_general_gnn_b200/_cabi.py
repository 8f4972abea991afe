"""ctypes binding of libgta_b200.so (include/gta_b200.h).

This is the ONLY way the Python host reaches the device: plain pointers and sizes, no
torch types cross the boundary (torch only owns the memory and the stream).  There is no
CPU fallback: if the library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_TAG = os.environ.get("GTA_LIB_TAG", "")          # experiment builds, see build.py
LIB_PATH = os.path.join(_HERE, "libgta_b200" + ("_" + _TAG if _TAG else "") + ".so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, 1, 2, 3, 4
EPI_NONE, EPI_ELU, EPI_RELU = 0, 1, 2
W_NONE, W_EDGE, W_EDGE_DIV = 0, 1, 2
PHASE_MAIN, PHASE_RESET, PHASE_ALL, PHASE_STATIC = 1, 2, 3, 4
OPND_EDGE, OPND_DST, OPND_SRC = 0, 1, 2
BIN_ADD, BIN_MUL, BIN_DIV = 0, 1, 2
UN_EXP_LEAKY_RELU, UN_ELU, UN_RELU, UN_COPY = 0, 1, 2, 3

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f32 = C.c_float
_sz = C.c_size_t

MAX_RANKS = 16


class Exchange(C.Structure):
    """gta_exchange_t (include/gta_b200.h)."""
    _fields_ = [("world", _i32), ("rank", _i32), ("step", _i32), ("copy_ctas", _i32), ("slot_rows", _i64),
                ("row_bytes", _i64), ("table", _p), ("signals", _p), ("peer_table", _p * MAX_RANKS),
                ("slot_valid_rows", _i64 * MAX_RANKS)]


#: every symbol include/gta_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gta_last_error": (C.c_char_p, []),
    "gta_abi_version": (C.c_int, []),
    "gta_launch_count": (_i64, []),
    "gta_launch_count_reset": (None, []),
    "gta_csr_build_workspace": (_sz, [_i64, _i64]),
    "gta_csr_build": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "gta_tile_nnz": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _p, _p]),
    "gta_tile_nnz_max": (C.c_int, [_p, _p, _i64, _i64, _p, _sz, C.POINTER(_i32), _p]),
    "gta_partition": (C.c_int, [_p, _i64, _i32, _p, _p]),
    "gta_remap_sources": (C.c_int, [_p, _i64, _p, _i32, _i64, _i32, _p, _p]),
    "gta_ipc_alloc": (C.c_int, [_sz, C.POINTER(_p)]),
    "gta_ipc_free": (C.c_int, [_p]),
    "gta_ipc_export": (C.c_int, [_p, C.c_char_p]),
    "gta_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(_p)]),
    "gta_ipc_close": (C.c_int, [_p]),
    "gta_exchange_signal_bytes": (_sz, []),
    "gta_exchange_publish": (C.c_int, [_p, _i32, _i32, _i32, _i32, C.POINTER(_p), _p]),
    "gta_reorder_workspace": (_sz, [_i64]),
    "gta_reorder": (C.c_int, [_p, _i64, _p, _p, _sz, _p]),
    "gta_schedule_workspace": (_sz, [_i64, _i64, _i64]),
    "gta_schedule_max_items": (_i64, [_i64, _i64, _i32, _i64, _i64]),
    "gta_schedule_col_blocks": (_i32, [_i64, _i64]),
    "gta_schedule_build": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _i64, _p, _i64, _p, C.POINTER(_i64),
                                     C.POINTER(_i64), _p, _sz, _p]),
    "gta_schedule_build_cuts": (C.c_int, [_p, _p, _i64, _i64, _i32, C.POINTER(_i64), _i32, _p, _i64, _p, C.POINTER(_i64),
                                          C.POINTER(_i64), _p, _sz, _p]),
    "gta_gemm_workspace": (_sz, [_i32, _i32]),
    "gta_gemm_f32": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _p, _i32, _p, _p, _i64, _p, _sz,
                               _p]),
    "gta_gemm_f32_zbf16": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _p, _i32, _p, _p, _i64, _p, _sz,
                                     _p]),
    "gta_aggregate_bf16": (C.c_int, [_p, _i64, _p, _i64, _p, _i32, _p, _i32, _p, _p, _i64, _p, _i64, _i32, _i32,
                                     _p, _p, _p, _i32, _p]),
    "gta_gat_aggregate_bf16": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _i32, _f32, _p, _i64, _p, _i64,
                                         _i32, _i32, _p, _p, _p, _p, _p, _i64, _p, _i32, _p]),
    "gta_gemm_set_mode": (C.c_int, [C.c_int]),
    "gta_gemm_get_mode": (C.c_int, []),
    "gta_aggregate_f32": (C.c_int, [_p, _i64, _p, _i64, _p, _i32, _p, _i32, _p, _p, _i64, _p, _i64, _i32, _i32,
                                    _p, _p, _p, _i32, _p]),
    "gta_aggregate_edge_sum_f32": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _i64, _p, _i64, _p, _i64, _i32, _f32, _p, _i64,
                                             _i32, _i32, _p, _p, _i32, _p]),
    "gta_gat_partial_stride": (_i32, [_i32, _i32]),
    "gta_gather_peak_probe": (C.c_int, [_p, _i64, _i64, _i32, _i64, _p, _p]),
    "gta_er_stats": (C.c_int, [_p, _i64, _i64, _i64, _i32, _p, _p]),
    "gta_gat_aggregate_f32": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _i32, _f32, _p, _i64, _p, _i64,
                                        _i32, _i32, _p, _p, _p, _p, _p, _i64, _p, _i32, _p]),
    "gta_gat_logits_f32": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _i32, _f32, _i32, _p, _p, _p, _p]),
    "gta_edge_binary_f32": (C.c_int, [_p, _p, _i64, _i64, _i32, _p, _i32, _i32, _i64, _p, _i32, _i32, _i64, _p, _i32,
                                      _i64, _p]),
    "gta_edge_unary_f32": (C.c_int, [_p, _p, _i64, _i64, _i32, _f32, _p, _i32, _i32, _i64, _p, _i64, _p]),
    "gta_node_binary_f32": (C.c_int, [_i32, _p, _i32, _i64, _p, _i32, _i64, _p, _i32, _i64, _i64, _p]),
    "gta_node_unary_f32": (C.c_int, [_i32, _f32, _p, _i64, _p, _i64, _i32, _i64, _p]),
}


class GtaError(RuntimeError):
    def __init__(self, code: int, where: str, message: str):
        super().__init__(f"{where} failed with code {code}: {message}")
        self.code = code


class GtaUnsupported(GtaError):
    """Legal ISA, but no kernel for this shape yet (GTA_ERR_UNSUPPORTED)."""


_lib = None


def load() -> C.CDLL:
    """Load the shared library, binding every declared symbol.  Raises if it is absent --
    run ``python -m gta_graph_tensor_acclelrator_for_general_gnn_b200.build`` first."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. There is no CPU fallback; "
            "run `python -m gta_graph_tensor_acclelrator_for_general_gnn_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the .so diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, where: str) -> None:
    if code == OK:
        return
    msg = load().gta_last_error().decode("utf-8", "replace")
    if code == ERR_UNSUPPORTED:
        raise GtaUnsupported(code, where, msg)
    raise GtaError(code, where, msg)


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None stays None)."""
    if t is None:
        return None
    return t.data_ptr()
