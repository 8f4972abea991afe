"""ctypes loader of oracle/_ref/libgta_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY:
import from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libgta_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "gta_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        last = None
        for extra in ([], ["GTA_CC=gcc"], ["GTA_CC=gcc", "OMP="]):     # last resort: single-threaded
            last = subprocess.run(["make", "-C", HERE, "-B", "all"] + extra, capture_output=True, text=True)
            if last.returncode == 0:
                break
        else:
            raise RuntimeError("could not build the C oracle:\n" + last.stdout + last.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.gta_oracle_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def threads() -> int:
    return int(load().gta_oracle_threads())


def use_all_cores() -> int:
    """Lift the OMP_NUM_THREADS=1 that torchrun exports: OpenMP (this library) and the BLAS behind
    numpy both get every host core.  Returns the OpenMP thread count."""
    n = os.cpu_count() or 1
    load().gta_oracle_set_threads(C.c_int(n))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return threads()


def _suffix(dtype):
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]


def gemm(x, w, dtype=np.float32):
    x = np.ascontiguousarray(x, dtype=dtype)
    w = np.ascontiguousarray(w, dtype=dtype)
    n, k = x.shape
    f = w.shape[1]
    assert f <= 1024
    z = np.empty((n, f), dtype=dtype)
    getattr(load(), "gta_oracle_gemm_" + _suffix(dtype))(_p(x), C.c_int64(k), _p(w), _p(z), C.c_int64(f),
                                                         C.c_int64(n), C.c_int(k), C.c_int(f))
    return z


def proj(z, a, dtype=np.float32):
    z = np.ascontiguousarray(z, dtype=dtype)
    a = np.ascontiguousarray(a, dtype=dtype)
    n, f = z.shape
    h = a.shape[1]
    out = np.empty((n, h), dtype=dtype)
    getattr(load(), "gta_oracle_proj_" + _suffix(dtype))(_p(z), C.c_int64(f), _p(a), _p(out), C.c_int64(n),
                                                         C.c_int(f), C.c_int(h))
    return out


def spmm(indptr, indices, w, x, dtype=np.float32, row_begin=0, row_end=None):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=dtype)
    w = None if w is None else np.ascontiguousarray(w, dtype=dtype).reshape(-1)
    row_end = indptr.shape[0] - 1 if row_end is None else row_end
    f = x.shape[1]
    assert f <= 1024
    out = np.empty((row_end - row_begin, f), dtype=dtype)
    getattr(load(), "gta_oracle_spmm_" + _suffix(dtype))(_p(indptr), _p(indices), _p(w), _p(x), C.c_int64(f), _p(out),
                                                         C.c_int64(f), C.c_int64(row_begin), C.c_int64(row_end),
                                                         C.c_int(f))
    return out


def gat_edge_phase(indptr, indices, el, er, z, slope=0.2, dtype=np.float32, row_begin=0, row_end=None,
                   activation=True):
    """el is indexed by (row - row_begin); er and z by source id."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    el = np.ascontiguousarray(el, dtype=dtype)
    er = np.ascontiguousarray(er, dtype=dtype)
    z = np.ascontiguousarray(z, dtype=dtype)
    row_end = indptr.shape[0] - 1 if row_end is None else row_end
    f = z.shape[1]
    heads = er.shape[1]
    assert f <= 1024 and heads <= 64 and f % heads == 0
    out = np.empty((row_end - row_begin, f), dtype=dtype)
    ctype = C.c_float if np.dtype(dtype) == np.float32 else C.c_double
    getattr(load(), "gta_oracle_gat_" + _suffix(dtype))(_p(indptr), _p(indices), _p(el), _p(er), C.c_int(heads),
                                                        ctype(slope), _p(z), C.c_int64(f), _p(out), C.c_int64(f),
                                                        C.c_int64(row_begin), C.c_int64(row_end), C.c_int(f),
                                                        C.c_int(int(activation)))
    return out


def gat_edge_phase_scaled(indptr, indices, el, er, z, zabs, slope=0.2, row_begin=0, row_end=None, activation=True):
    """fp64 edge phase plus the per-element error scale ``sum_k alpha_k * zabs[src k]`` (``zabs = |X|.|W|``):
    returns ``(out, scale)``.  ``el`` is indexed by (row - row_begin); er, z, zabs by source id."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    el, er, z, zabs = (np.ascontiguousarray(a, dtype=np.float64) for a in (el, er, z, zabs))
    row_end = indptr.shape[0] - 1 if row_end is None else row_end
    f = z.shape[1]
    heads = er.shape[1]
    assert f <= 1024 and heads <= 64 and f % heads == 0 and zabs.shape == z.shape
    out = np.empty((row_end - row_begin, f), dtype=np.float64)
    scale = np.empty_like(out)
    load().gta_oracle_gat_scaled_f64(_p(indptr), _p(indices), _p(el), _p(er), C.c_int(heads), C.c_double(slope), _p(z),
                                     _p(zabs), C.c_int64(f), _p(out), _p(scale), C.c_int64(f), C.c_int64(row_begin),
                                     C.c_int64(row_end), C.c_int(f), C.c_int(int(activation)))
    return out, scale


def gat_layer(indptr, indices, x, w, al, ar, dtype=np.float32, row_begin=0, row_end=None):
    """Whole GAT layer (ops 0-13) on host cores: GEMM, projections, edge phase."""
    z = gemm(x, w, dtype)
    el = proj(z, al, dtype)
    er = proj(z, ar, dtype)
    row_end = indptr.shape[0] - 1 if row_end is None else row_end
    return gat_edge_phase(indptr, indices, el[row_begin:row_end], er, z, dtype=dtype, row_begin=row_begin,
                          row_end=row_end)


def tile_nnz(indptr, indices, n, tile_rows):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    tiles = -(-n // tile_rows)
    out = np.empty((tiles, n), dtype=np.int64)
    load().gta_oracle_tile_nnz(_p(indptr), _p(indices), C.c_int64(n), C.c_int64(tile_rows), _p(out))
    return out
