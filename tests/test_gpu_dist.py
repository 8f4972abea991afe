"""Destination-partitioned execution, all ranks emulated on ONE GPU (no collective kernel waits
on another): every rank's partition is built, the gathered source table is assembled by hand, and
the per-rank results must reproduce the single-GPU result bit for bit."""
import numpy as np
import pytest

from oracle import gta_oracle as O
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partitioned_gat_layer_equals_single_gpu(world):
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import dist as gdist, graph, kernels
    n, e, fin, f, h = 3000, 90000, 96, 128, 4
    g = synthetic.powerlaw_graph(n, e, seed=2, i0=3.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    full = graph.csr_from_coo(g.dst, g.src, n)
    x, w, al, ar = synthetic.gat_tensors(n, fin, f, h, seed=1)
    dev = lambda a: torch.from_numpy(a).cuda()
    xd, wd, ald, ard = kernels.to_table(dev(x)), dev(w), dev(al), dev(ar)
    z, el, er = kernels.gemm(xd, wd, ald, ard)
    want = kernels.gat_aggregate(full, el, er, z)

    parts = [gdist.make_partition(full, r, world) for r in range(world)]
    bounds = parts[0].bounds
    assert bounds == [int(v) for v in O.partition_bounds(indptr, world)]
    stride = parts[0].stride
    z_all = torch.zeros((world * stride, f), device="cuda")
    er_all = torch.zeros((world * stride, h), device="cuda")
    el_loc = []
    for p in parts:       # every rank's local GEMM, written into its slot of the gathered tables
        zl, ell, erl = kernels.gemm(xd[p.row_begin:p.row_end], wd, ald, ard)
        z_all[p.rank * stride: p.rank * stride + p.rows] = zl
        er_all[p.rank * stride: p.rank * stride + p.rows] = erl
        el_loc.append(ell)
    got = torch.cat([kernels.gat_aggregate(p.local, el_loc[p.rank], er_all, z_all) for p in parts])
    assert torch.equal(got, want)
    # edge balance: no rank holds more than its share plus one row
    loads = [p.local.num_edges for p in parts]
    assert sum(loads) == e and max(loads) - e / world <= np.diff(indptr).max()
