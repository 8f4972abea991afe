"""SURVEY.md section 8(f)-2 on the GPU: the DGN and PNA op graphs (COMP_MM on edges, one-input binaries,
PNA-trans' self-referencing producers) built by this package's own generator + lowering, executed by
the CUDA kernels and compared with the op-by-op oracle.  (Named to sort last: newest coverage runs last.)"""
import numpy as np
import pytest

import test_cpu_executor as C
import test_gpu_executor as shared
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
    ref = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs, semantics=sem, stabilize=True)
    dev = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    out, log = executor.execute(records, op_info, dg, dev(node_inputs), dev(weights), dev(edge_inputs), network=network,
                                is_reorder=reorder, fuse_across_blocks=fuse, return_log=True)
    (p, y), = out.items()
    y64 = ref[p]
    np.testing.assert_allclose(y.cpu().numpy(), y64, rtol=1e-4, atol=2e-5 * np.abs(y64).max(), err_msg=str(log))
    assert any(k == "gta_gemm_f32:edges" for k, _ in log)
