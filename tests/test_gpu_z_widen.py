"""SURVEY.md section 8(f)-2 on the GPU: the DGN and PNA op graphs (COMP_MM on edges, one-input binaries,
PNA-trans' self-referencing producers) built by this package's own generator + lowering, executed by
the CUDA kernels and compared with the op-by-op oracle.  (Named to sort last: newest coverage runs last.)"""
import numpy as np
import pytest

import test_cpu_executor as C
import test_gpu_executor as shared
from conftest import assert_close_rowscale
from oracle import gta_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("network,reorder,plan_kind", C.WIDE, ids=[f"{n}-{'trans' if r else 'original'}-{k}" for n, r, k in C.WIDE])
@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "stores-honoured"])
def test_dgn_pna_match_oracle(network, reorder, plan_kind, fuse):
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor

    class RT:
        pass
    rt = RT()
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph
    rt.graph = graph
    op_info, records = C.wide_program(network, reorder, plan_kind)
    g, indptr, indices, dg = shared._graph(rt, "cora")
    node_inputs, weights, edge_inputs = shared._inputs(op_info, g.num_nodes, g.num_edges)
    sem = O.NETWORK_SEMANTICS.get((network, reorder), {})
    ref, ref_scale = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs, semantics=sem, stabilize=True, return_scale=True)
    dev = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    out, log = executor.execute(records, op_info, dg, dev(node_inputs), dev(weights), dev(edge_inputs), network=network,
                                is_reorder=reorder, fuse_across_blocks=fuse, return_log=True)
    (p, y), = out.items()
    y64 = ref[p]
    assert_close_rowscale(y.cpu().numpy(), y64, ref_scale[p], what=str(log))
    if network == "DGN" and (fuse or plan_kind == "one-block"):      # linear edge phase: no E x F tensor at all
        assert not any(k == "gta_gemm_f32:edges" or k.startswith("gta_edge_") for k, _ in log), log
        assert any(k == "gta_aggregate_f32:scatter_sum" for k, _ in log), log
    else:
        assert any(k == "gta_gemm_f32:edges" for k, _ in log)
    if network == "PNA" and (fuse or plan_kind == "one-block"):      # ops 5-8 in one pass, nothing E x F written after op 2
        assert any(k == "gta_aggregate_edge_sum_f32" for k, _ in log), log
        assert not any(k.startswith("gta_edge_") for k, _ in log), log


def test_npz_ingest_matches_scipy(tmp_path):
    """the SciPy-CSR .npz the simulator reads -> DeviceGraph + edge values, bit-exact
    against SciPy's own canonical form, rows given with columns in arbitrary order."""
    import scipy.sparse as sp
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph, kernels
    a = sp.random(300, 300, density=0.05, random_state=5, format="csr", dtype=np.float32)
    a.setdiag(1.0)
    a = a.tocsr()
    # shuffle the columns inside every row: a legal CSR archive that is not in canonical order
    rng = np.random.default_rng(1)
    ind, dat = a.indices.copy(), a.data.copy()
    for r in range(300):
        lo, hi = a.indptr[r], a.indptr[r + 1]
        perm = rng.permutation(hi - lo)
        ind[lo:hi], dat[lo:hi] = ind[lo:hi][perm], dat[lo:hi][perm]
    np.savez(tmp_path / "adj.npz", data=dat, indices=ind, indptr=a.indptr, shape=np.array(a.shape), format=b"csr")
    a.sort_indices()
    for drop in (False, True):
        want = a.copy()
        if drop:
            want.setdiag(0)
            want.eliminate_zeros()
        g, w = graph.csr_from_npz(str(tmp_path / "adj.npz"), drop_diagonal=drop)
        assert g.num_nodes == 300 and g.num_edges == want.nnz
        assert np.array_equal(g.indptr.cpu().numpy(), want.indptr.astype(np.int64))
        assert np.array_equal(g.indices.cpu().numpy(), want.indices.astype(np.int32))
        assert np.array_equal(w.cpu().numpy(), want.data)
        x = torch.randn(300, 32, device="cuda")
        y = kernels.aggregate(g, kernels.to_table(x), w)
        np.testing.assert_allclose(y.cpu().numpy(), want @ x.cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "stores-honoured"])
def test_order_c_gather_on_device(fuse):
    """ORDER C gather (sum per source): CSC walk built by gta_csr_build over (source, edge id)."""
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph, lowering

    class RT:
        pass
    rt = RT()
    rt.graph = graph
    g, indptr, indices, dg = shared._graph(rt, "cora")
    n, e = g.num_nodes, g.num_edges
    op_info = C.column_gather_case()
    for op in op_info:      # re-stamp the generator's N/E to this graph
        op["INPUT"]["feature_number"] = [e if op["TYPE"] in ("applyedge", "gather") else n] * len(op["INPUT"]["feature_number"])
    records = lowering.lower(op_info, [[0], [1, 2, 3]], [[64, 1], [64, 1]], n)
    node_inputs, weights, edge_inputs = shared._inputs(op_info, n, e)
    ref, ref_scale = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs, return_scale=True)
    dev = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    out, log = executor.execute(records, op_info, dg, dev(node_inputs), dev(weights), dev(edge_inputs),
                                fuse_across_blocks=fuse, return_log=True)
    assert_close_rowscale(out[3].cpu().numpy(), ref[3], ref_scale[3])
    assert any(k == "gta_aggregate_f32:by_source" for k, _ in log)
    again = executor.execute(records, op_info, dg, dev(node_inputs), dev(weights), dev(edge_inputs), fuse_across_blocks=fuse)
    assert torch.equal(out[3], again[3])       # deterministic


@pytest.mark.parametrize("plan", [[[0, 1, 2, 3, 4, 5]], [[0, 1, 2, 3], [4, 5]]], ids=["one-block", "store-between"])
@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "stores-honoured"])
def test_edge_mm_feeding_a_gather_on_device(plan, fuse):
    """COMP_MM_COMP_ADD: the GEMM runs over N rows whatever the plan (linearity of the edge phase)."""
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph, lowering

    class RT:
        pass
    rt = RT()
    rt.graph = graph
    g, indptr, indices, dg = shared._graph(rt, "cora")
    n, e = g.num_nodes, g.num_edges
    op_info = C.mm_then_gather_case(n, e)
    records = lowering.lower(op_info, plan, [[64, 1]] * len(plan), n)
    node_inputs, weights, edge_inputs = shared._inputs(op_info, n, e)
    ref, ref_scale = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs, return_scale=True)
    dev = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    out, log = executor.execute(records, op_info, dg, dev(node_inputs), dev(weights), dev(edge_inputs),
                                fuse_across_blocks=fuse, return_log=True)
    assert_close_rowscale(out[5].cpu().numpy(), ref[5], ref_scale[5], what=str(log))
    names = [k for k, _ in log]
    stored_between = len(plan) == 2 and not fuse
    # MM over ADD(scatter, scatter) distributes into two N-row GEMMs + one segment sum: never an E-row GEMM; a stored
    # edge tensor is materialised once at the output width
    assert "gta_gemm_f32:edges" not in names, names
    assert ("gta_aggregate_f32:scatter_sum" in names or "gta_gemm_f32:after_gather" in names) == (not stored_between), names
    assert ("gta_edge_binary_f32" in names) == stored_between, names


def test_random_op_graphs_on_device():
    """The op-graph fuzz of tests/test_cpu_executor_fuzz.py with the CUDA kernels in place of the test double
    (tests/device_fuzz.py; its first 150 cases ran clean on a B200 in round 1)."""
    import device_fuzz
    bad, unsupported, messages = device_fuzz.run_cases(0, 60)
    assert bad == 0 and unsupported == 0, "\n".join(messages)
