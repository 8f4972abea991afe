"""At-scale parity checker.  TEST INFRASTRUCTURE ONLY (see oracle/gta_oracle.py): imported by tests/,
``__graft_entry__.smoke()`` and bench.py's checker / CPU-baseline legs, never by the product package.

Compares a layer output computed on the GPU with the fp64 C oracle (oracle/gta_oracle.c, checked against
the numpy oracle in tests/test_cpu_oracle_c.py) under the stated fp32 tolerance

    |y - y64| <= 1e-5 |y64| + 1e-5 rowscale,      rowscale[i,c] = sum_k |coef_k| * (|X|.|W|)[src k, c]

where ``coef_k`` is the softmax coefficient (GAT) or the edge weight (GCN) and ``|X|.|W|`` bounds the
rounding error of the gathered ``Z = X.W`` entries -- the reduction behind one output element.  It works
on any subset of destination rows, so Reddit-size outputs are checked in seconds (all rows by default).
"""
from __future__ import annotations

import numpy as np

from . import c_oracle

RTOL = 1e-5


def host_tables(x, w, al=None, ar=None):
    """fp64 ``Z = X.W``, its error scale ``|X|.|W|`` and (GAT) ``el = Z.Al``, ``er = Z.Ar``."""
    x64 = np.asarray(x, dtype=np.float64)
    w64 = np.asarray(w, dtype=np.float64)
    z = x64 @ w64
    zabs = np.abs(x64) @ np.abs(w64)
    if al is None:
        return z, zabs, None, None
    return z, zabs, z @ np.asarray(al, dtype=np.float64), z @ np.asarray(ar, dtype=np.float64)


def select_rows(indptr: np.ndarray, max_edges: int, top: int = 512) -> np.ndarray:
    """Destination rows to check: all of them when they hold at most ``max_edges`` edges, else the ``top``
    highest-degree rows (the longest reductions) plus every k-th row, k chosen to fit the budget."""
    n = indptr.shape[0] - 1
    e = int(indptr[-1])
    if e <= max_edges or n == 0:
        return np.arange(n, dtype=np.int64)
    deg = np.diff(indptr)
    heavy = np.argsort(-deg, kind="stable")[:top]
    k = max(int(np.ceil(e / max(max_edges - int(deg[heavy].sum()), 1))), 1)
    return np.unique(np.concatenate([heavy, np.arange(0, n, k, dtype=np.int64)]))


def sub_csr(indptr: np.ndarray, indices: np.ndarray, rows: np.ndarray):
    """Compact CSR (indptr, indices) of the selected rows, edges in their original order."""
    if rows.shape[0] == indptr.shape[0] - 1:
        return np.ascontiguousarray(indptr, dtype=np.int64), np.ascontiguousarray(indices, dtype=np.int32)
    deg = (indptr[rows + 1] - indptr[rows]).astype(np.int64)
    out_ptr = np.zeros(rows.shape[0] + 1, dtype=np.int64)
    np.cumsum(deg, out=out_ptr[1:])
    pos = np.repeat(indptr[rows] - out_ptr[:-1], deg) + np.arange(int(out_ptr[-1]), dtype=np.int64)
    return out_ptr, np.ascontiguousarray(indices[pos], dtype=np.int32)


#: the bf16 storage mode's own tolerance: rtol 2e-2, atol 1e-2 * rowscale
BF16 = (2e-2, 1e-2)


def _report(y, y64, scale, rows, edges, rtol):
    y = np.asarray(y, dtype=np.float64)
    rtol, atol = rtol if isinstance(rtol, tuple) else (rtol, rtol)
    bound = rtol * np.abs(y64) + atol * scale + 1e-30
    err = np.abs(y - y64)
    finite = bool(np.all(np.isfinite(y)))
    worst = float(np.max(err / bound)) if err.size else 0.0
    return {"rows": int(rows), "edges": int(edges), "max_err_over_tol": worst if finite else float("inf"),
            "max_abs_err": float(err.max()) if err.size else 0.0, "finite": finite,
            "tolerance": f"|y-y64| <= {rtol:g}*|y64| + {atol:g}*rowscale, rowscale = sum_k |coef_k|*(|X|.|W|)[src k]",
            "oracle": "oracle/gta_oracle.c fp64 (OpenMP), ascending-source reduction"}


def check_gat(y_rows, indptr_s, indices_s, el_rows, er, z, zabs, slope=0.2, activation=True, rtol=RTOL):
    """``y_rows``: GPU output of the selected rows; ``indptr_s/indices_s``: their compact CSR (global source
    ids); ``el_rows``: fp64 el of the selected rows; ``er, z, zabs``: fp64 source tables."""
    y64, scale = c_oracle.gat_edge_phase_scaled(indptr_s, indices_s, el_rows, er, z, zabs, slope=slope,
                                                activation=activation)
    return _report(y_rows, y64, scale, indptr_s.shape[0] - 1, indptr_s[-1], rtol)


def check_gcn(y_rows, indptr_s, indices_s, edge_w_s, z, zabs, rtol=RTOL):
    """GCN-trans layer ``Y = A^ (X W)``: ``edge_w_s`` are the edge weights of the selected rows' edges."""
    w64 = None if edge_w_s is None else np.asarray(edge_w_s, dtype=np.float64).reshape(-1)
    y64 = c_oracle.spmm(indptr_s, indices_s, w64, z, dtype=np.float64)
    scale = c_oracle.spmm(indptr_s, indices_s, None if w64 is None else np.abs(w64), zabs, dtype=np.float64)
    return _report(y_rows, y64, scale, indptr_s.shape[0] - 1, indptr_s[-1], rtol)
