"""Parity at the BASELINE shapes (round-1 verdict: nothing above 4 000 nodes had ever been compared with anything).

Full Flickr shape (89 250 nodes / 899 756 edges / 500 features) and full Reddit shape (232 965 nodes / 114.6 M edges /
602 features), executed from the reference-emitted ISA programs through ``execute()`` with the DEFAULT work list
(``schedule_for``: the Reddit-shape 119 MB source table is walked in 3 column blocks and every row is a 3-item chain),
compared with the fp64 C oracle under the stated tolerance 1e-5 |y| + 1e-5 rowscale (oracle/parity.py)."""
import os

import numpy as np
import pytest
import yaml

from oracle import gta_oracle as O, parity as P
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(rel):
    with open(os.path.join(GOLDEN, rel)) as f:
        return yaml.safe_load(f)


@pytest.fixture(scope="module")
def rt():
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph, kernels

    class NS:
        pass
    ns = NS()
    ns.torch, ns.ex, ns.graph, ns.k = torch, executor, graph, kernels
    ns.dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return ns


@pytest.fixture(scope="module")
def flickr(rt):
    n, e, fin = synthetic.SHAPES["flickr"]
    g = synthetic.shape_graph("flickr")
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    dg = rt.graph.csr_from_coo(g.dst, g.src, n)
    assert np.array_equal(dg.indptr.cpu().numpy(), indptr) and np.array_equal(dg.indices.cpu().numpy(), indices)
    return n, e, fin, indptr, indices, dg


def test_flickr_gcn_two_layers_chained(rt, flickr):
    """BASELINE config 3: the 2-layer Flickr GCN (500 -> 128 -> 64), layer 1's device output is layer 2's input.
    Each layer is checked against the oracle applied to the input that layer actually saw, and the final output
    against the oracle of the whole chain with the first layer's error scale carried through."""
    n, e, fin, indptr, indices, dg = flickr
    ew = synthetic.gcn_edge_norm(indptr, indices)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, fin), dtype=np.float32)
    w1, w2 = synthetic.glorot(rng, fin, 128), synthetic.glorot(rng, 128, 64)
    ew_d = rt.dev(ew)[:, None]
    outs = []
    x_d = rt.dev(x)
    for layer, w in ((1, w1), (2, w2)):
        op_info = _load(f"opgraph/GCN-flickr-layer{layer}-trans.yaml")
        prog = _load(f"isa/GCN-flickr-layer{layer}-trans__0_1-2-3.yaml")
        out, log = rt.ex.execute(prog, op_info, dg, {0: x_d}, {0: rt.dev(w)}, {2: ew_d}, network="GCN", is_reorder=True,
                                 return_log=True)
        assert [k for k, _ in log] == ["gta_gemm_f32", "gta_aggregate_f32:w"], log
        x_d = out[3]
        outs.append(out[3].cpu().numpy())
    # per layer, on the input the layer saw
    z, zabs, _, _ = P.host_tables(x, w1)
    rep1 = P.check_gcn(outs[0], indptr, indices, ew, z, zabs)
    z2, zabs2, _, _ = P.host_tables(outs[0], w2)
    rep2 = P.check_gcn(outs[1], indptr, indices, ew, z2, zabs2)
    assert rep1["max_err_over_tol"] <= 1.0 and rep2["max_err_over_tol"] <= 1.0, (rep1, rep2)
    # whole chain in fp64; the error scale of layer 1's output is layer 2's input scale
    from oracle import c_oracle
    y1 = c_oracle.spmm(indptr, indices, ew.astype(np.float64), z, dtype=np.float64)
    s1 = c_oracle.spmm(indptr, indices, ew.astype(np.float64), zabs, dtype=np.float64)
    rep = P.check_gcn(outs[1], indptr, indices, ew, y1 @ w2.astype(np.float64), 3.0 * s1 @ np.abs(w2).astype(np.float64))
    assert rep["max_err_over_tol"] <= 1.0, rep


def test_flickr_gat_layer(rt, flickr):
    """Flickr-shape GAT layer 1 as genGraphOP writes it (H = 16: lane-local-head kernel), full shape."""
    n, e, fin, indptr, indices, dg = flickr
    op_info = _load("opgraph/GAT-flickr-layer1-original.yaml")
    prog = _load("isa/GAT-flickr-layer1-original__0-1-2_4-5-6-7-8_3-9-10-11-12-13.yaml")
    heads = op_info[1]["OUTPUT"]["size_per_feature"] // 4
    x, w, al, ar = synthetic.gat_tensors(n, fin, 128, heads, seed=2)
    out, log = rt.ex.execute(prog, op_info, dg, {0: rt.dev(x)}, {0: rt.dev(w), 1: rt.dev(al), 2: rt.dev(ar)},
                             network="GAT", is_reorder=False, return_log=True)
    assert "gta_gat_aggregate_f32" in [k for k, _ in log]
    z, zabs, el, er = P.host_tables(x, w, al, ar)
    rep = P.check_gat(out[13].cpu().numpy(), indptr, indices, el, er, z, zabs)
    assert rep["max_err_over_tol"] <= 1.0 and rep["rows"] == n, rep


@pytest.fixture(scope="module")
def reddit(rt):
    n, e, fin = synthetic.SHAPES["reddit"]
    g = synthetic.shape_graph("reddit")
    dg = rt.graph.csr_from_coo(g.dst, g.src, n)
    del g
    return n, e, fin, dg.indptr.cpu().numpy(), dg.indices.cpu().numpy(), dg


def test_reddit_gat_layer_default_schedule(rt, reddit):
    """The benchmarked configuration: Reddit-shape GAT (H = 4) through the default schedule_for() path -- 3 column
    blocks, every destination row a chain of 3 items -- checked on the 512 highest-degree rows (reductions of up to
    ~12 000 terms) plus a strided sample, about 30 M edges."""
    n, e, fin, indptr, indices, dg = reddit
    op_info = _load("opgraph/GAT-reddit-restamped-h4.yaml")
    prog = _load("isa/GAT-reddit-layer1-original__0-1-2_4-5-6-7-8_3-9-10-11-12-13.yaml")
    x, w, al, ar = synthetic.gat_tensors(n, fin, 128, 4, seed=0)
    sched = dg.schedule_for(128 * 4)
    assert sched.num_blocks == 3 and sched.num_slots >= 3 * n - 3, (sched.num_blocks, sched.num_slots)
    run = lambda: rt.ex.execute(prog, op_info, dg, {0: rt.dev(x)}, {0: rt.dev(w), 1: rt.dev(al), 2: rt.dev(ar)},
                                network="GAT", is_reorder=False)[13]
    y = run()
    assert rt.torch.equal(y, run()), "not bitwise reproducible"
    rows = P.select_rows(indptr, 30_000_000)
    ip, ix = P.sub_csr(indptr, indices, rows)
    z, zabs, el, er = P.host_tables(x, w, al, ar)
    rep = P.check_gat(y.cpu().numpy()[rows], ip, ix, el[rows], er, z, zabs)
    assert rep["max_err_over_tol"] <= 1.0, rep


def test_reddit_gcn_layer_default_schedule(rt, reddit):
    n, e, fin, indptr, indices, dg = reddit
    op_info = _load("opgraph/GCN-reddit-layer1-trans.yaml")
    prog = _load("isa/GCN-reddit-layer1-trans__0_1-2-3.yaml")
    rng = np.random.default_rng(4)
    x = rng.standard_normal((n, fin), dtype=np.float32)
    w = synthetic.glorot(rng, fin, 128)
    deg = np.maximum(np.diff(indptr), 1).astype(np.float64)
    rows = P.select_rows(indptr, 30_000_000)
    ip, ix = P.sub_csr(indptr, indices, rows)
    pos = np.repeat(indptr[rows] - ip[:-1], np.diff(ip)) + np.arange(int(ip[-1]), dtype=np.int64)
    # edge weights on device (E x 1 fp32 = 458 MB), 1/sqrt(deg_i deg_j)
    deg_d = rt.dev(deg)
    row_of_edge = rt.torch.repeat_interleave(rt.torch.arange(n, device="cuda"), rt.dev(np.diff(indptr)))
    ew_d = (1.0 / rt.torch.sqrt(deg_d[row_of_edge] * deg_d[dg.indices.long()])).to(rt.torch.float32)[:, None].contiguous()
    del row_of_edge
    y = rt.ex.execute(prog, op_info, dg, {0: rt.dev(x)}, {0: rt.dev(w)}, {2: ew_d}, network="GCN", is_reorder=True)[3]
    z, zabs, _, _ = P.host_tables(x, w)
    rep = P.check_gcn(y.cpu().numpy()[rows], ip, ix, ew_d.cpu().numpy().reshape(-1)[pos], z, zabs)
    assert rep["max_err_over_tol"] <= 1.0, rep
