"""The at-scale parity checker (oracle/parity.py) itself: its fp64 C path equals the numpy oracle, a row subset
gives the same answer as the whole graph, an fp32-rounded correct result passes and a wrong one fails."""
import numpy as np

from oracle import c_oracle, gta_oracle as O, parity
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic


def _case(n=700, e=9000, fin=40, f=32, h=4, seed=3):
    g = synthetic.powerlaw_graph(n, e, seed=seed, i0=10.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    x, w, al, ar = synthetic.gat_tensors(n, fin, f, h, seed=seed)
    return indptr, indices, x, w, al, ar


def test_scaled_gat_oracle_matches_numpy_oracle_and_rowscale():
    indptr, indices, x, w, al, ar = _case()
    z, zabs, el, er = parity.host_tables(x, w, al, ar)
    ref = O.gat_layer(indptr, indices, x, w, al, ar)
    y64, scale = c_oracle.gat_edge_phase_scaled(indptr, indices, el, er, z, zabs)
    np.testing.assert_allclose(y64, ref["Y"], rtol=1e-12, atol=1e-14)
    want = O.segment_sum(O.head_broadcast(ref["alpha"], z.shape[1]) * zabs[indices], indptr)
    np.testing.assert_allclose(scale, want, rtol=1e-12, atol=1e-14)


def test_row_subset_equals_whole_graph_rows():
    indptr, indices, x, w, al, ar = _case()
    z, zabs, el, er = parity.host_tables(x, w, al, ar)
    full, _ = c_oracle.gat_edge_phase_scaled(indptr, indices, el, er, z, zabs)
    rows = parity.select_rows(indptr, max_edges=2000, top=16)
    assert 0 < rows.shape[0] < indptr.shape[0] - 1
    deg = np.diff(indptr)
    assert set(np.argsort(-deg, kind="stable")[:16]) <= set(rows)          # the longest reductions are in
    ip, ix = parity.sub_csr(indptr, indices, rows)
    assert ip[-1] == deg[rows].sum()
    sub, _ = c_oracle.gat_edge_phase_scaled(ip, ix, el[rows], er, z, zabs)
    assert np.array_equal(sub, full[rows])
    all_rows = parity.select_rows(indptr, max_edges=10**9)
    assert np.array_equal(all_rows, np.arange(indptr.shape[0] - 1))
    ip2, ix2 = parity.sub_csr(indptr, indices, all_rows)
    assert np.array_equal(ip2, indptr) and np.array_equal(ix2, indices)


def test_checker_accepts_fp32_result_and_rejects_a_wrong_one():
    indptr, indices, x, w, al, ar = _case()
    z, zabs, el, er = parity.host_tables(x, w, al, ar)
    y32 = c_oracle.gat_layer(indptr, indices, x, w, al, ar, dtype=np.float32)          # an honest fp32 computation
    ok = parity.check_gat(y32, indptr, indices, el, er, z, zabs)
    assert ok["finite"] and ok["max_err_over_tol"] <= 1.0 and ok["rows"] == indptr.shape[0] - 1
    bad = y32.copy()
    bad[5, 7] += 1e-2
    assert parity.check_gat(bad, indptr, indices, el, er, z, zabs)["max_err_over_tol"] > 1.0
    bad[5, 7] = np.nan
    assert parity.check_gat(bad, indptr, indices, el, er, z, zabs)["max_err_over_tol"] == float("inf")


def test_gcn_checker():
    indptr, indices, x, w, _, _ = _case()
    ew = synthetic.gcn_edge_norm(indptr, indices)
    z, zabs, _, _ = parity.host_tables(x, w)
    y32 = c_oracle.spmm(indptr, indices, ew, (x @ w).astype(np.float32), dtype=np.float32)
    rep = parity.check_gcn(y32, indptr, indices, ew, z, zabs)
    assert rep["max_err_over_tol"] <= 1.0
    rows = np.array([3, 10, 11, 500], dtype=np.int64)
    ip, ix = parity.sub_csr(indptr, indices, rows)
    pos = np.concatenate([np.arange(indptr[r], indptr[r + 1]) for r in rows])
    assert parity.check_gcn(y32[rows], ip, ix, ew[pos], z, zabs)["max_err_over_tol"] <= 1.0
    assert parity.check_gcn(y32[rows] * 1.001, ip, ix, ew[pos], z, zabs)["max_err_over_tol"] > 1.0
