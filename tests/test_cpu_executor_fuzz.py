"""Random op graphs x random fusion plans through the real ``execute()`` (CPU test double for the kernels,
tests/host_kernels.py) against the op-by-op oracle.

The golden programs cover the seven shipped networks; the executor's pattern matching (which lazy
expression becomes which fused kernel, what may stay virtual, what a STORE_* forces) must also be right
for op graphs nobody wrote by hand.  Each case draws a typed DAG of scatter / applyedge / gather /
applynode ops (incl. COMP_MM on nodes and edges, ORDER C, '-1' edge weights, MUL-that-means-divide),
lowers it under a random plan with this package's ``lowering.lower`` and compares every final output.
"""
import os
import random

import numpy as np
import pytest
import torch

import host_kernels
import test_cpu_executor as C
from oracle import gta_oracle as O
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph, isa, lowering, opgraph, synthetic

N, E = 120, 900
CASES = int(os.environ.get("GTA_FUZZ_CASES", "300"))


class _Val:
    def __init__(self, pos, on_edges, width, bound, positive):
        self.pos, self.on_edges, self.width, self.bound, self.positive = pos, on_edges, width, bound, positive
        self.consumers = []


def _random_graph(rng, n, e, max_deg):
    """Returns (op_info, executor semantics, oracle semantics)."""
    ops, vals, sem_x, sem_o = [], [], {}, {}

    def emit(comp, kind, order, producers, in_widths, out_width, bound, positive):
        pos = len(ops)
        rows_in = e if kind in ("applyedge", "gather") else n
        rows_out = e if kind in ("scatter", "applyedge") else n
        ops.append(opgraph.gen_one_op(pos, comp, kind, order, [rows_in] * len(in_widths), list(producers), len(in_widths),
                                      int(comp == "MM"), [], [in_widths[0] * out_width * 4] if comp == "MM" else [],
                                      [w * 4 for w in in_widths], [], rows_out, out_width * 4))
        for q in producers:
            if q != -1:
                ops[q]["OUTPUT"]["output_list"].append(pos)
                vals[q].consumers.append(pos)
        v = _Val(pos, kind in ("scatter", "applyedge"), out_width, bound, positive)
        vals.append(v)
        return v

    width0 = rng.choice([4, 8, 16])
    emit("NONE", "scatter", rng.choice("RC"), [], [width0], width0, 0.5, False)       # external node tensor -> edges
    target = rng.randint(4, 14)
    guard = 0
    while len(ops) < target and guard < 200:
        guard += 1
        nodes = [v for v in vals if not v.on_edges]
        edges = [v for v in vals if v.on_edges]
        move = rng.choice(["source", "scatter", "edge_bin", "edge_w", "edge_sf", "edge_mm", "gather", "node_mm",
                           "node_bin", "node_sf", "gather", "edge_bin", "gat"])
        if move == "source":
            w = rng.choice([4, 8, 16])
            if rng.random() < 0.5:
                emit("NONE", "scatter", rng.choice("RC"), [], [w], w, 0.5, False)
            else:
                wo = rng.choice([4, 8])
                emit("MM", "applynode", "R", [], [w], wo, 0.5 * w * 0.3, False)
        elif move == "scatter" and nodes:
            a = rng.choice(nodes)
            emit("NONE", "scatter", rng.choice("RC"), [a.pos], [a.width], a.width, a.bound, a.positive)
        elif move == "edge_bin" and len(edges) >= 1:
            a = rng.choice(edges)
            mates = [b for b in edges if b is not a and (b.width == a.width or (a.width % b.width == 0 and b.width in (1, 4)))]
            if not mates:
                continue
            b = rng.choice(mates)
            comp = rng.choice(["ADD", "MUL", "MUL"])
            if comp == "MUL" and b.positive and b.bound < 1e4 and rng.random() < 0.5:
                pos = len(ops)
                sem_x[pos], sem_o[pos] = "div", "div_first_by_second"
                emit("MUL", "applyedge", "R", [a.pos, b.pos], [a.width, b.width], a.width, a.bound * 50, a.positive)
            elif comp == "MUL":
                emit("MUL", "applyedge", "R", [a.pos, b.pos], [a.width, b.width], a.width, a.bound * b.bound,
                     a.positive and b.positive)
            else:
                emit("ADD", "applyedge", "R", [a.pos, b.pos], [a.width, b.width], a.width, a.bound + b.bound,
                     a.positive and b.positive)
        elif move == "edge_w" and edges:
            a = rng.choice(edges)
            emit("MUL", "applyedge", "R", [a.pos, -1], [a.width, a.width], a.width, a.bound, a.positive)
        elif move == "edge_sf" and edges:
            a = rng.choice([v for v in edges if v.bound <= 4.0] or [None])
            if a is not None:
                emit("SF", "applyedge", "R", [a.pos], [a.width], a.width, float(np.exp(a.bound)), True)
        elif move == "edge_mm" and edges:
            a = rng.choice(edges)
            wo = rng.choice([4, 8])
            emit("MM", "applyedge", "R", [a.pos], [a.width], wo, a.bound * a.width * 0.3, False)
        elif move == "gather" and edges:
            a = rng.choice(edges)
            order = "C" if rng.random() < 0.2 else "R"
            emit("ADD", "gather", order, [a.pos], [a.width], a.width, a.bound * max_deg, a.positive)
        elif move == "node_mm" and nodes:
            a = rng.choice(nodes)
            wo = rng.choice([4, 8, 16])
            emit("MM", "applynode", "R", [a.pos], [a.width], wo, a.bound * a.width * 0.3, False)
        elif move == "node_bin" and len(nodes) >= 2:
            a, b = rng.sample(nodes, 2)
            if a.width % b.width and b.width % a.width:
                continue
            if a.width < b.width:
                a, b = b, a
            comp = rng.choice(["ADD", "MUL"])
            if comp == "MUL" and b.positive and rng.random() < 0.5:
                pos = len(ops)
                sem_x[pos], sem_o[pos] = "div", "div_first_by_second"
                emit("MUL", "applynode", "R", [a.pos, b.pos], [a.width, b.width], a.width, a.bound * 50, a.positive)
            else:
                emit(comp, "applynode", "R", [a.pos, b.pos], [a.width, b.width], a.width,
                     a.bound * b.bound if comp == "MUL" else a.bound + b.bound, a.positive and b.positive)
        elif move == "gat" and nodes:
            # the attention motif with random omissions: logits from two projections, exp, row sum, optional
            # normalisation (edge-side divide, node-side divide, or none), weighted aggregation of a third table
            base = rng.choice(nodes)
            heads = rng.choice([1, 2, 4])
            width = heads * rng.choice([1, 4, 8])
            z = emit("MM", "applynode", "R", [base.pos], [base.width], width, 2.0, False)
            el = emit("MM", "applynode", "R", [z.pos], [width], heads, 0.5, False)
            er = emit("MM", "applynode", "R", [z.pos], [width], heads, 0.5, False)
            a = emit("NONE", "scatter", "R", [el.pos], [heads], heads, 0.5, False)
            b = emit("NONE", "scatter", "C", [er.pos], [heads], heads, 0.5, False)
            s_ = emit("ADD", "applyedge", "R", [a.pos, b.pos] if rng.random() < 0.7 else [b.pos, a.pos], [heads, heads],
                      heads, 1.0, False)
            p_ = emit("SF", "applyedge", "R", [s_.pos], [heads], heads, 3.0, True)
            zs = emit("NONE", "scatter", "C", [z.pos], [width], width, 2.0, False)
            style = rng.choice(["edge_div", "node_div", "none"])
            if style == "edge_div":
                S = emit("ADD", "gather", "R", [p_.pos], [heads], heads, 3.0 * max_deg, True)
                Sk = emit("NONE", "scatter", "R", [S.pos], [heads], heads, 3.0 * max_deg, True)
                pos = len(ops)
                sem_x[pos], sem_o[pos] = "div", "div_first_by_second"
                al = emit("MUL", "applyedge", "R", [p_.pos, Sk.pos], [heads, heads], heads, 1.0, True)
                m = emit("MUL", "applyedge", "R", [zs.pos, al.pos] if rng.random() < 0.7 else [al.pos, zs.pos],
                         [width, heads], width, 2.0, False)
                emit("ADD", "gather", "R", [m.pos], [width], width, 2.0, False)
            else:
                m = emit("MUL", "applyedge", "R", [zs.pos, p_.pos], [width, heads], width, 6.0, False)
                num = emit("ADD", "gather", "R", [m.pos], [width], width, 6.0 * max_deg, False)
                if style == "node_div":
                    S = emit("ADD", "gather", "R", [p_.pos], [heads], heads, 3.0 * max_deg, True)
                    pos = len(ops)
                    if rng.random() < 0.5:
                        sem_x[pos], sem_o[pos] = "rdiv", "div_second_by_first"
                        emit("MUL", "applynode", "R", [S.pos, num.pos], [heads, width], width, 2.0, False)
                    else:
                        sem_x[pos], sem_o[pos] = "div", "div_first_by_second"
                        emit("MUL", "applynode", "R", [num.pos, S.pos], [width, heads], width, 2.0, False)
        elif move == "node_sf" and nodes:
            a = rng.choice(nodes)
            emit("SF", "applynode", "R", [a.pos], [a.width], a.width, a.bound, False)
    # every dangling EDGE value gets reduced so that the finals are node tensors (or small edge tensors)
    for v in list(vals):
        if v.on_edges and not v.consumers:
            emit("ADD", "gather", "R", [v.pos], [v.width], v.width, v.bound * max_deg, v.positive)
    return ops, sem_x, sem_o


def _random_plan(rng, n_ops):
    order = list(range(n_ops))
    blocks, i = [], 0
    while i < n_ops:
        k = rng.choice([1, 2, 3, 5, n_ops])
        blocks.append(order[i:i + k])
        i += k
    return blocks, [[16 * rng.randint(1, 8), 1] for _ in blocks]


@pytest.fixture(scope="module")
def small_graph():
    g = synthetic.powerlaw_graph(N, E, seed=3, i0=6.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, N)
    dg = graph.DeviceGraph(N, g.num_edges, torch.from_numpy(indptr), torch.from_numpy(indices.astype(np.int32)),
                           num_sources=N)
    return g, indptr, indices, dg


@pytest.mark.parametrize("seed", range(CASES))
def test_random_op_graph_matches_oracle(small_graph, monkeypatch, seed):
    monkeypatch.setattr(executor, "kernels", host_kernels)
    monkeypatch.setattr(executor, "csr_from_coo", C._host_csr_from_coo)
    monkeypatch.setattr(graph.DeviceGraph, "schedule", lambda self, *a, **k: None)
    g, indptr, indices, dg = small_graph
    dg.schedules.clear()
    rng = random.Random(1000 + seed)
    max_deg = int(max(np.diff(indptr).max(), np.bincount(indices, minlength=N).max()))
    op_info, sem_x, sem_o = _random_graph(rng, N, g.num_edges, max_deg)
    n_ops = len(op_info)
    plan, tiles = _random_plan(rng, n_ops)
    try:
        records = lowering.lower(op_info, plan, tiles, N)
        # with several fusable pairs in one block the reference's instruction fusion (mirrored byte for byte)
        # can drop COMP records; such a program names fewer ops than the op graph and execute() refuses it
        isa.Program.from_records(records).block_ops(op_info)
    except (lowering.LoweringError, isa.IsaError):
        plan, tiles = [[i] for i in range(n_ops)], [[32, 1]] * n_ops        # the unfused plan always lowers
        records = lowering.lower(op_info, plan, tiles, N)
    data = np.random.default_rng(seed)
    node_inputs, weights, edge_inputs = {}, {}, {}
    for pos, op in enumerate(op_info):
        win = op["INPUT"]["size_per_feature"][0] // 4
        if op["COMP_TYPE"] == "MM":
            weights[pos] = data.uniform(-0.3, 0.3, size=(win, op["OUTPUT"]["size_per_feature"] // 4)).astype(np.float32)
        if not op["INPUT"]["input_g_list"]:
            node_inputs[pos] = data.uniform(-0.5, 0.5, size=(N, win)).astype(np.float32)
        if -1 in op["INPUT"]["input_g_list"]:
            edge_inputs[pos] = data.uniform(0.1, 1.0, size=(g.num_edges, 1)).astype(np.float32)
    ref, ref_scale = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs, semantics=sem_o,
                                   stabilize=False, fix_gat_op10=False, return_scale=True)
    fuse = bool(seed % 2)
    out, log = executor.execute(records, op_info, dg, C._t(node_inputs), C._t(weights), C._t(edge_inputs),
                                semantics=sem_x, stabilize=False, fuse_across_blocks=fuse, return_log=True)
    finals = [p for p in range(n_ops) if not op_info[p]["OUTPUT"]["output_list"]]
    assert sorted(out) == finals
    for p in finals:
        want = ref[p] if ref[p].ndim == 2 else ref[p][:, None]
        got = out[p].numpy()
        assert got.shape == want.shape, (p, got.shape, want.shape)
        sc = ref_scale[p] if ref_scale[p].ndim == 2 else ref_scale[p][:, None]
        C.assert_close_rowscale(got, want, sc, what=f"op {p}, plan {plan}, fuse {fuse}, kernels {log}")
