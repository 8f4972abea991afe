"""execute()'s HOST logic on CPU: every golden ISA program is run through the real executor with
the kernel wrappers replaced by the torch-CPU test double (tests/host_kernels.py) and compared with
the op-by-op oracle.  What this checks is the executor itself -- dataflow from the YAML, block order,
the pattern matching onto fused kernels, dead stores, refusals; the CUDA kernels are checked on the
GPU (test_gpu_executor.py runs the same programs there)."""
import json
import os

import numpy as np
import pytest
import torch
import yaml

import host_kernels
import test_gpu_executor as shared
from conftest import assert_close_rowscale
from oracle import gta_oracle as O
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import _cabi, executor, graph, isa, lowering, opgraph, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "manifest.json")) as _f:
    PROGRAMS = json.load(_f)["programs"]
N, E = 700, 9000


def _load(rel):
    with open(os.path.join(GOLDEN, rel)) as f:
        return yaml.safe_load(f)


@pytest.fixture()
def host(monkeypatch):
    """executor wired to the test double; work lists are a kernel-side structure, so none is built"""
    monkeypatch.setattr(executor, "kernels", host_kernels)
    monkeypatch.setattr(graph.DeviceGraph, "schedule", lambda self, *a, **k: None)
    g = synthetic.powerlaw_graph(N, E, seed=11, i0=20.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, N)
    dg = graph.DeviceGraph(N, g.num_edges, torch.from_numpy(indptr), torch.from_numpy(indices.astype(np.int32)),
                           num_sources=N)
    return g, indptr, indices, dg


def _t(d):
    up = lambda v: [torch.from_numpy(a) for a in v] if isinstance(v, list) else torch.from_numpy(v)
    return {k: up(v) for k, v in d.items()}


class _Ref(dict):
    """oracle outputs per op, with the matching error scales in ``.scale``"""


def _run(host, op_info, records, network, reorder, fuse=True, **kw):
    g, indptr, indices, dg = host
    node_inputs, weights, edge_inputs = shared._inputs(op_info, N, g.num_edges)
    sem = O.NETWORK_SEMANTICS.get((network, reorder), {})
    ref, ref_scale = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs, semantics=sem,
                                   stabilize=True, return_scale=True)
    ref = _Ref(ref)
    ref.scale = ref_scale      # first-order error scale per op: the rowscale of the stated 1e-5 tolerance
    out, log = executor.execute(records, op_info, dg, _t(node_inputs), _t(weights), _t(edge_inputs), network=network,
                                is_reorder=reorder, fuse_across_blocks=fuse, check_shapes=False, return_log=True, **kw)
    return out, ref, [k for k, _ in log]


@pytest.mark.parametrize("prog", PROGRAMS, ids=[p["file"].split("/")[-1][:-5] for p in PROGRAMS])
@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "stores-honoured"])
def test_golden_program_dataflow(host, prog, fuse):
    op_info = _load(prog["opgraph"])
    out, ref, names = _run(host, op_info, _load(prog["file"]), prog["network"], prog["reorder"], fuse)
    finals = [p for p in range(len(op_info)) if not op_info[p]["OUTPUT"]["output_list"]]
    assert sorted(out) == finals
    for p in finals:
        y64 = ref[p]
        assert_close_rowscale(out[p].numpy(), y64, ref.scale[p], what=str(names))
    if prog["network"] == "GAT" and fuse:
        f_out, heads = op_info[0]["OUTPUT"]["size_per_feature"] // 4, op_info[1]["OUTPUT"]["size_per_feature"] // 4
        # one pass whatever the layer: whole pieces per head, or the narrow heads of layer 3 (F = H = 16)
        assert (f_out // heads) % 4 == 0 or f_out // heads in (1, 2)
        assert "gta_gat_aggregate_f32" in names and not any(k.startswith("gta_edge_") for k in names), names
    if prog["network"] in ("GCN", "SGC", "GraphSAGE", "GIN"):
        assert any(k.startswith("gta_aggregate_f32") for k in names), names
        if fuse:
            assert not any(k.startswith("gta_edge_") for k in names), names


def test_generate_lower_execute_without_any_reference_file(host):
    """opgraph.build -> lowering.lower -> execute: the whole host chain of this package."""
    n_ref = synthetic.SHAPES["cora"][0]
    for network, reorder, plan, tiles in (("GCN", True, [[0], [1, 2, 3]], [[512, 1], [512, 1]]),
                                          ("GraphSAGE", False, [[0, 1, 2, 3], [4, 5, 6]], [[64, 1], [64, 1]])):
        op_info = opgraph.build(*synthetic.SHAPES["cora"], network, 2, reorder, repair=True)
        records = lowering.lower(op_info, plan, tiles, n_ref)
        out, ref, names = _run(host, op_info, records, network, reorder)
        (p, y), = out.items()
        assert_close_rowscale(y.numpy(), ref[p], ref.scale[p])


def test_refusals_need_no_gpu(host):
    g, indptr, indices, dg = host
    prog = next(p for p in PROGRAMS if p["file"].endswith("GCN-cora-layer1-original__0_1-2-3.yaml"))
    op_info, records = _load(prog["opgraph"]), _load(prog["file"])
    node_inputs, weights, edge_inputs = shared._inputs(op_info, N, g.num_edges)
    with pytest.raises(executor.ExecutionError, match="max_edge_bytes"):
        executor.execute(records, op_info, dg, _t(node_inputs), _t(weights), _t(edge_inputs), network="GCN",
                         fuse_across_blocks=False, max_edge_bytes=1 << 20, check_shapes=False)
    with pytest.raises(executor.ExecutionError, match="generated for"):
        executor.execute(records, op_info, dg, _t(node_inputs), _t(weights), _t(edge_inputs), network="GCN")
    with pytest.raises(executor.ExecutionError, match="weights"):
        executor.execute(records, op_info, dg, _t(node_inputs), {}, _t(edge_inputs), network="GCN", check_shapes=False)
    with pytest.raises(executor.ExecutionError, match="edge_inputs"):
        executor.execute(records, op_info, dg, _t(node_inputs), _t(weights), {}, network="GCN", check_shapes=False)
    col = [dict(op) for op in op_info]
    col[2] = dict(col[2], ORDER="X")
    with pytest.raises(isa.IsaError, match="ORDER"):
        executor.execute(records, col, dg, _t(node_inputs), _t(weights), _t(edge_inputs), network="GCN", check_shapes=False)
    with pytest.raises(isa.IsaError):
        executor.execute([[dict(records[0][0], TYPE="COMP_FOO")]], op_info, dg, {}, {})


# ---- DGN / PNA: COMP_MM on edges, one-input binaries, PNA-trans' self references ---------------
WIDE = [("DGN", False, "per-op"), ("DGN", False, "one-block"), ("PNA", False, "per-op"), ("PNA", False, "one-block"),
        ("PNA", True, "per-op")]


def wide_program(network, reorder, plan_kind, layer=2):
    """(op graph, ISA records) built by this package's own generator + lowering (Cora shape)."""
    shape = synthetic.SHAPES["cora"]
    op_info = opgraph.build(*shape, network, layer, reorder)
    n_ops = len(op_info)
    plan = [[i] for i in range(n_ops)] if plan_kind == "per-op" else [list(range(n_ops))]
    return op_info, lowering.lower(op_info, plan, [[64, 1]] * len(plan), shape[0])


@pytest.mark.parametrize("network,reorder,plan_kind", WIDE, ids=[f"{n}-{'trans' if r else 'original'}-{k}" for n, r, k in WIDE])
@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "stores-honoured"])
def test_dgn_pna_dataflow(host, network, reorder, plan_kind, fuse):
    op_info, records = wide_program(network, reorder, plan_kind)
    out, ref, names = _run(host, op_info, records, network, reorder, fuse)
    (p, y), = out.items()
    assert_close_rowscale(y.numpy(), ref[p], ref.scale[p], what=str(names))
    if network == "DGN" and (fuse or plan_kind == "one-block"):
        # the whole edge phase is linear: MM distributed over the ADD of scatters (two N-row GEMMs), the gather of the
        # four scatters = one plain segment sum + degree x own rows -- no E-row GEMM, no generic edge kernel
        assert "gta_gemm_f32:edges" not in names and "gta_aggregate_f32:scatter_sum" in names, names
        assert not any(k.startswith("gta_edge_") for k in names), names
    else:
        assert "gta_gemm_f32:edges" in names          # PNA op 2 (edge features), DGN with every STORE honoured
    if network == "PNA" and not reorder and fuse:
        # ops 3/4 = MM(scatter(x)): commuted to scatter(MM(x)) -> two N-row GEMMs, one E-row GEMM (op 2)
        assert names.count("gta_gemm_f32:edges") == 1 and names.count("gta_gemm_f32") == 3, names
    if network == "PNA" and (fuse or plan_kind == "one-block"):
        # ops 5-8 = gather(SF(edge + scatterC + scatterR)): one pass over the edges, no E x F intermediate written
        assert "gta_aggregate_edge_sum_f32" in names and not any(k.startswith("gta_edge_") for k in names), names


def test_edge_mm_respects_the_edge_budget(host):
    g, indptr, indices, dg = host
    # PNA op 2 is a GEMM over edge features: an E x F tensor that cannot be rewritten away
    op_info, records = wide_program("PNA", False, "one-block")
    node_inputs, weights, edge_inputs = shared._inputs(op_info, N, g.num_edges)
    with pytest.raises(executor.ExecutionError, match="max_edge_bytes"):
        executor.execute(records, op_info, dg, _t(node_inputs), _t(weights), _t(edge_inputs), network="PNA",
                         max_edge_bytes=1 << 16, check_shapes=False)
    # DGN's edge phase is linear and never materialises E x F (even with a tiny budget); a missing input still raises
    op_info, records = wide_program("DGN", False, "one-block")
    node_inputs, weights, edge_inputs = shared._inputs(op_info, N, g.num_edges)
    executor.execute(records, op_info, dg, _t(node_inputs), _t(weights), _t(edge_inputs), network="DGN",
                     max_edge_bytes=1 << 16, check_shapes=False)
    node_inputs.pop(9)
    with pytest.raises(executor.ExecutionError, match=r"node_inputs\[9\]"):
        executor.execute(records, op_info, dg, _t(node_inputs), _t(weights), _t(edge_inputs), network="DGN", check_shapes=False)


# ---- ORDER C gathers (sum per source) -----------------------------------------------------------
def _host_csr_from_coo(dst, src, num_nodes, want_perm=False):
    """torch-CPU stand-in for graph.csr_from_coo (sort by (dst, src)), test double like host_kernels"""
    key = dst.long() * (int(src.max()) + 1 if src.numel() else 1) + src.long()
    order = torch.argsort(key, stable=True)
    counts = torch.bincount(dst.long(), minlength=num_nodes)
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(counts, 0)
    return graph.DeviceGraph(num_nodes, int(dst.numel()), indptr, src[order].contiguous(), order if want_perm else None,
                             num_sources=num_nodes)


def column_gather_case():
    """A small program whose reduction is column-wise: x -> scatter R -> x edge weight -> gather C -> MM,
    i.e. Y = (A^T diag-weighted X) W: every node sums what it SENT.  The reference has the lowering rules
    for ORDER C gathers (interpreter.py:55-129) but no shipped network uses one."""
    op_info = opgraph.build(N, E, 32, "GCN", 2, False)
    op_info[0]["ORDER"] = "R"
    op_info[2]["ORDER"] = "C"
    return op_info


@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "stores-honoured"])
def test_order_c_gather_sums_per_source(host, monkeypatch, fuse):
    monkeypatch.setattr(executor, "csr_from_coo", _host_csr_from_coo)
    op_info = column_gather_case()
    records = lowering.lower(op_info, [[0], [1, 2, 3]], [[64, 1], [64, 1]], N)
    out, ref, names = _run(host, op_info, records, None, False, fuse)
    assert_close_rowscale(out[3].numpy(), ref[3], ref.scale[3])
    assert "gta_aggregate_f32:by_source" in names
    # and the oracle's column-wise sum is what it says: every node sums what it sent
    g, indptr, indices, dg = host
    node_inputs, weights, edge_inputs = shared._inputs(op_info, N, g.num_edges)
    x = node_inputs[0].astype(np.float64)
    rows = np.repeat(np.arange(N), np.diff(indptr))
    direct = np.zeros((N, x.shape[1]))
    np.add.at(direct, indices, edge_inputs[1].astype(np.float64) * x[rows])
    np.testing.assert_allclose(ref[2], direct, rtol=1e-12)


# ---- COMP_MM_COMP_ADD: an edge GEMM feeding a gather ----------------------------------------------
def mm_then_gather_case(n, e, fin=32, fout=16):
    """scatter C + scatter R -> ADD -> MM (edges) -> gather -> SF: the (applyedge, gather) / (MM, ADD) pattern of the
    reference's fusion table (hardware_info.yaml:27-30), which no shipped network exercises."""
    op = opgraph.gen_one_op
    return [op(0, "NONE", "scatter", "C", [n], [], 1, 0, [], [], [fin * 4], [2], e, fin * 4),
            op(1, "NONE", "scatter", "R", [n], [], 1, 0, [], [], [fin * 4], [2], e, fin * 4),
            op(2, "ADD", "applyedge", "R", [e, e], [0, 1], 2, 0, [], [], [fin * 4, fin * 4], [3], e, fin * 4),
            op(3, "MM", "applyedge", "R", [e], [2], 1, 1, [], [fin * fout * 4], [fin * 4], [4], e, fout * 4),
            op(4, "ADD", "gather", "R", [e], [3], 1, 0, [], [], [fout * 4], [5], n, fout * 4),
            op(5, "SF", "applynode", "R", [n], [4], 1, 0, [], [], [fout * 4], [], n, fout * 4)]


@pytest.mark.parametrize("plan", [[[0, 1, 2, 3, 4, 5]], [[0, 1, 2], [3, 4], [5]], [[0, 1, 2, 3], [4, 5]]],
                         ids=["one-block", "mm+gather-block", "store-between"])
@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "stores-honoured"])
def test_edge_mm_feeding_a_gather_reduces_first(host, plan, fuse):
    g, indptr, indices, dg = host
    op_info = mm_then_gather_case(N, g.num_edges)
    records = lowering.lower(op_info, plan, [[64, 1]] * len(plan), N)
    if plan != [[0, 1, 2, 3], [4, 5]]:
        assert any(i["TYPE"] == "COMP_MM_COMP_ADD" for b in records for i in b)
    out, ref, names = _run(host, op_info, records, None, False, fuse)
    assert_close_rowscale(out[5].numpy(), ref[5], ref.scale[5], what=str(names))
    stored_between = plan == [[0, 1, 2, 3], [4, 5]] and not fuse
    # never an E-row GEMM unless the edge tensor is stored: MM over ADD(scatter, scatter) distributes into two N-row
    # GEMMs and one segment sum (or, for a non-linear edge value, the gather runs first and the GEMM after it)
    assert ("gta_aggregate_f32:scatter_sum" in names or "gta_gemm_f32:after_gather" in names) == (not stored_between), names
    # a stored edge tensor is materialised once, at the output width (the GEMM still runs on N rows)
    assert "gta_gemm_f32:edges" not in names, names
    assert ("gta_edge_binary_f32" in names) == (not fuse and len(plan) > 1), names      # some STORE_E is honoured


# ---- corrupted inputs are rejected, never crashed on (SURVEY 8b "Errors") ---------------------------
_JUNK = [None, 5, "x", [], {}, "COMP_", "LOAD_Q", "9_applyedge", "a_b_c", -1, 3.5, ["a"], {"TYPE": 1}, "COMP_MM",
         "STORE_E", "99_gather_0", "STORE_N", "LOAD_E", "1_scatter_0", "2_gather_0", "COMP_MUL_COMP_ADD", [99], [0, 0, 0],
         "scatter", "gather", "MM", "C", 4, 6, [-1, -1]]
_REJECTIONS = (isa.IsaError, executor.ExecutionError, _cabi.GtaUnsupported)


def _corrupt_program(rng, rec):
    for _ in range(rng.randint(1, 3)):
        if not isinstance(rec, list) or not rec:
            return rec
        b = rng.randrange(len(rec))
        if not isinstance(rec[b], list) or not rec[b]:
            continue
        i = rng.randrange(len(rec[b]))
        m = rng.random()
        if m < 0.2 and isinstance(rec[b][i], dict) and rec[b][i]:
            del rec[b][i][rng.choice(list(rec[b][i]))]
        elif m < 0.6 and isinstance(rec[b][i], dict) and rec[b][i]:
            rec[b][i][rng.choice(["TYPE", "ID"] + list(rec[b][i]))] = rng.choice(_JUNK)
        elif m < 0.7:
            rec[b][i] = rng.choice(_JUNK)
        elif m < 0.9:
            del rec[b][i]
        else:
            rec[b] = rng.choice(_JUNK)
    return rec


def _corrupt_op_graph(rng, op_info):
    for _ in range(rng.randint(1, 2)):
        i = rng.randrange(len(op_info))
        op = op_info[i]
        if not isinstance(op, dict) or not op:
            continue
        m = rng.random()
        if m < 0.25:
            del op[rng.choice(list(op))]
        elif m < 0.5:
            op[rng.choice(list(op))] = rng.choice(_JUNK)
        elif m < 0.9:
            sub = op.get(rng.choice(["INPUT", "OUTPUT"]))
            if isinstance(sub, dict) and sub:
                k = rng.choice(list(sub))
                if rng.random() < 0.3:
                    del sub[k]
                else:
                    sub[k] = rng.choice(_JUNK)
        else:
            op_info[i] = rng.choice(_JUNK)
    return op_info


@pytest.mark.parametrize("what", ["program", "op-graph"])
def test_corrupted_inputs_are_rejected_not_crashed_on(host, what):
    """Random damage to an ISA program or to an op graph: execute() either runs or raises one of its three
    documented error types -- never a KeyError / TypeError / IndexError from deep inside."""
    import copy
    import random
    g, indptr, indices, dg = host
    progs = [p for p in PROGRAMS if p["dataset"] == "cora" and (p["layer"] > 1 or p["network"] == "GCN")]
    progs = [p for p in progs if p["layer"] > 1][:4] + [p for p in progs if p["layer"] == 1][:1]
    loaded = {p["file"]: (_load(p["file"]), _load(p["opgraph"])) for p in progs}
    tensors = {f: tuple(_t(d) for d in shared._inputs(op, N, g.num_edges)) for f, (_, op) in loaded.items()}
    rng = random.Random(5 if what == "program" else 6)
    ran = refused = 0
    for trial in range(400):
        p = rng.choice(progs)
        records, op_info = copy.deepcopy(loaded[p["file"]][0]), copy.deepcopy(loaded[p["file"]][1])
        node_inputs, weights, edge_inputs = tensors[p["file"]]
        if what == "program":
            records = _corrupt_program(rng, records)
        else:
            op_info = _corrupt_op_graph(rng, op_info)
        try:
            executor.execute(records, op_info, dg, node_inputs, weights, edge_inputs, network=p["network"],
                             is_reorder=p["reorder"], check_shapes=False, fuse_across_blocks=bool(trial % 2))
            ran += 1
        except _REJECTIONS:
            refused += 1
    assert refused > 100 and ran + refused == 400


def test_whole_chain_on_files_without_the_reference(host, tmp_path, monkeypatch):
    """INTEGRATION.md's stand-alone flow, on disk and CWD-relative like the reference's: gen_yaml -> unfused_plan ->
    interpret -> execute_files (the arguments of simulate())."""
    monkeypatch.chdir(tmp_path)
    g, indptr, indices, dg = host
    n, e, f = synthetic.SHAPES["cora"]
    op_info = opgraph.gen_yaml(opgraph.network_path("GAT", "cora", 2, False), n, e, f, "GAT", 2, False)
    op_array, tile_size_list = lowering.unfused_plan(op_info)
    lowering.interpret("cora", "GAT", False, "layer2", op_array, tile_size_list)
    assert os.path.exists("Results/Insts/GAT-cora-layer2-original.yaml")
    node_inputs, weights, edge_inputs = shared._inputs(op_info, N, g.num_edges)
    out = executor.execute_files(tile_size_list, "cora", "GAT", "layer2", False, dg, _t(node_inputs), _t(weights),
                                 _t(edge_inputs), check_shapes=False)
    ref, ref_scale = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs,
                        semantics=O.NETWORK_SEMANTICS[("GAT", False)], stabilize=True, return_scale=True)
    assert_close_rowscale(out[13].numpy(), ref[13], ref_scale[13])
