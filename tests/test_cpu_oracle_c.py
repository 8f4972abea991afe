"""The C restatement (oracle/gta_oracle.c, the timed CPU baseline) against the numpy oracle."""
import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import gta_oracle as O
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic


@pytest.fixture(scope="module")
def small():
    n, e = 600, 9000
    g = synthetic.powerlaw_graph(n, e, seed=6, i0=8.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    return n, indptr, indices


def test_c_gemm_and_projections(small):
    n, _, _ = small
    x, w, al, ar = synthetic.gat_tensors(n, 602, 128, 4, seed=2)
    z64 = O.gemm(x, w)
    np.testing.assert_allclose(C.gemm(x, w, np.float64), z64, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(C.gemm(x, w, np.float32), z64, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(C.proj(z64, al, np.float64), z64 @ al.astype(np.float64), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("f", [16, 128, 500])
def test_c_spmm(small, f):
    n, indptr, indices = small
    rng = np.random.default_rng(f)
    x = rng.standard_normal((n, f))
    w = synthetic.gcn_edge_norm(indptr, indices).astype(np.float64)
    np.testing.assert_allclose(C.spmm(indptr, indices, w, x, np.float64), O.spmm(indptr, indices, w, x),
                               rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(C.spmm(indptr, indices, None, x, np.float64, 100, 250),
                               O.spmm(indptr, indices, None, x)[100:250], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("heads", [1, 4, 16])
def test_c_gat_layer(small, heads):
    n, indptr, indices = small
    x, w, al, ar = synthetic.gat_tensors(n, 96, 128, heads, seed=1)
    ref = O.gat_layer(indptr, indices, x, w, al, ar)
    got = C.gat_layer(indptr, indices, x, w, al, ar, np.float64)
    np.testing.assert_allclose(got, ref["Y"], rtol=1e-10, atol=1e-12)
    part = C.gat_layer(indptr, indices, x, w, al, ar, np.float64, 37, 411)
    np.testing.assert_allclose(part, ref["Y"][37:411], rtol=1e-10, atol=1e-12)
    f32 = C.gat_layer(indptr, indices, x, w, al, ar, np.float32)
    np.testing.assert_allclose(f32, ref["Y"], rtol=2e-3, atol=2e-4)


def test_c_tile_nnz(small):
    n, indptr, indices = small
    for sr in (1, 16, 100, 600):
        assert np.array_equal(C.tile_nnz(indptr, indices, n, sr), O.tile_nnz(indptr, indices, n, sr))
    assert C.threads() >= 1
