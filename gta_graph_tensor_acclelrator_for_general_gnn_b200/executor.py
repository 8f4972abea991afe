"""``execute()``: functional execution of a GTA ISA program on B200.

This is the entry point that sits next to the reference's ``simulate()`` in the
``compile -> interpret -> simulate`` flow (vTCAD/code/test.py:10-15): it consumes exactly what
``simulate`` consumes -- the op graph (``Network/...yaml``) and the instruction program
(``Results/Insts/...yaml``, interpreter.py:809-853) -- and, where the simulator only counts
cycles (simulator.py:281-355), it computes the tensors.

How a program becomes kernels
-----------------------------
The op graph gives the dataflow, the ISA program gives the fused blocks and the points
where a tensor is materialised (``STORE_*`` / ``LOAD_*``).  Blocks are run in dependency
order (the reference sorts them by size, compiler.py:60, which is not an execution order).
Inside a block every op first becomes a *lazy value*; scatters stay virtual exactly as
``fuse_fetch`` removed their FETCH (interpreter.py:768-806).  Values are forced at the
block's ``STORE_*`` instructions, and forcing pattern-matches onto the fused kernels:

=============================================  ==========================================
lazy expression                                  kernel (include/gta_b200.h)
=============================================  ==========================================
MM(x, W) [+ MM(., Al), MM(., Ar) same block]     gta_gemm_f32 (el/er fused)
gather(MUL(scatterC(x), w))  COMP_MUL_COMP_ADD   gta_aggregate_f32 (W_EDGE)
gather(MUL(scatterC(x), p / scatterR(S)))        gta_aggregate_f32 (W_EDGE_DIV)
SF(ADD(scatterR(el), scatterC(er))) + its sum    gta_gat_logits_f32
whole GAT edge phase (ops 3..13 / trans 3..12)   gta_gat_aggregate_f32 (single pass)
applynode SF after a gather                      epilogue of the aggregate kernel
MM(scatter(x), W) on edges (PNA ops 3/4)         scatter(gta_gemm_f32(x, W)): N rows instead of E, same bits
MM(e, W) on any other edge tensor (DGN op 3)     gta_gemm_f32 over the E rows
anything else                                    gta_edge_* / gta_node_* generic kernels
gather with ORDER C (sum per SOURCE)             gta_csr_build over (source, edge id) once, then
                                                 gta_aggregate_f32 gathering rows of the edge tensor
=============================================  ==========================================

``fuse_across_blocks=True`` (default) treats a ``STORE_E``/``LOAD_E`` pair whose tensor has
no consumer outside the program as a dead store, which lets the two GAT edge blocks
collapse into the single-pass kernel; ``False`` honours every ``STORE_*`` of the program.
Plans that would materialise an ``E x F`` tensor larger than ``max_edge_bytes`` are refused
(Reddit would need 58.7 GB per tensor).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch
import yaml

from . import _cabi, kernels
from .graph import DeviceGraph, csr_from_coo
from .isa import IsaError, Program, stamp_comp_types, validate_op_graph

#: COMP_TYPE -> arithmetic where the YAML alone is ambiguous (the reference only names ops)
DEFAULT_SEMANTICS = {
    ("applyedge", "SF"): "exp_leaky_relu",   # GAT op 7: template/GAT_op.png 'f' then 'exp'
    ("applynode", "SF"): "elu",              # GAT op 13; activation unspecified upstream, pinned
    ("applyedge", "MUL"): "mul", ("applynode", "MUL"): "mul",
    ("applyedge", "ADD"): "add", ("applynode", "ADD"): "add",
}
#: MULs that are DIVIDES in GAT (template/GAT_op.png '/'), keyed (network, isReorder) -> {op: kind}
NETWORK_SEMANTICS = {
    ("GAT", False): {9: "div"},      # alpha = p / S[dst]        inputs [7, 10]
    ("GAT", True): {11: "rdiv"},     # O = O' / S                inputs [9 (S), 10 (O')]
    # PNA's edge SF closes the message MLP (genGraphOP.py:130,142); unnamed upstream, pinned to ReLU
    ("PNA", False): {7: "relu"},
    ("PNA", True): {7: "relu"},
}


class ExecutionError(RuntimeError):
    pass


def read_yaml(path: str):
    with open(path) as f:
        return yaml.safe_load(f)


def repair_op_graph(op_info, network: str | None, is_reorder: bool):
    """Return producer lists per op position with the one wiring quirk repaired:
    genGraphOP's GAT-original op 10 lists op 7 as input but consumes op 8's row sums
    (genGraphOP.py:59 vs V2/GAT_Cora.yaml:250-252)."""
    prods = [list(op["INPUT"]["input_g_list"]) for op in op_info]
    if (network == "GAT" and not is_reorder and len(op_info) == 14 and prods[10] == [7]
            and op_info[8]["TYPE"] == "gather" and op_info[10]["TYPE"] == "scatter"):
        prods[10] = [8]
    for pos, ins in enumerate(prods):
        if ins == [pos]:
            # PNA-trans ops 0/1 name themselves as producer (genGraphOP.py:136-137): an external input
            prods[pos] = []
    return prods


# ---- lazy values -----------------------------------------------------------------------------

@dataclass(eq=False)
class Value:
    kind: str                    # node | edge | scatter | edge_expr | gather | node_expr | mm
    op: str = ""                 # arithmetic: add mul div rdiv exp_leaky_relu elu relu
    args: tuple = ()
    side: str = ""               # scatter: 'R' (dst) or 'C' (src)
    width: int = 0
    tensor: torch.Tensor | None = None
    weight: torch.Tensor | None = None
    pos: int = -1
    extra: dict = field(default_factory=dict)

    @property
    def forced(self) -> bool:
        return self.tensor is not None

    @property
    def on_edges(self) -> bool:
        return self.kind in ("edge", "scatter", "edge_expr", "edge_mm")


class _Run:
    def __init__(self, graph: DeviceGraph, opts):
        self.g = graph
        self.o = opts
        self.kernel_log = []

    # -- helpers
    def _guard(self, width: int, what: str):
        need = self.g.num_edges * width * 4
        if need > self.o["max_edge_bytes"]:
            raise ExecutionError(
                f"plan materialises {what} as an E x {width} edge tensor ({need / 2**30:.1f} GiB > "
                f"max_edge_bytes); fuse it (COMP_MUL_COMP_ADD) or raise the budget")

    def _operand(self, v: Value):
        """(tensor, operand kind) for the generic edge kernels, scatters stay virtual."""
        if v.kind == "scatter":
            t = self.force(v.args[0])
            if v.side == "C":
                t = self.o["source_table"](t)
            return t, (_cabi.OPND_DST if v.side == "R" else _cabi.OPND_SRC)
        return self.force(v), _cabi.OPND_EDGE

    def _is_softmax_numerator(self, v: Value):
        """p = SF(ADD(scatterR(el), scatterC(er))) -> (el_value, er_value) or None."""
        if v.kind != "edge_expr" or v.op != "exp_leaky_relu":
            return None
        s = v.args[0]
        if s.kind != "edge_expr" or s.op != "add" or s.forced:
            return None
        a, b = s.args
        if a.kind == "scatter" and b.kind == "scatter" and {a.side, b.side} == {"R", "C"}:
            return (a.args[0], b.args[0]) if a.side == "R" else (b.args[0], a.args[0])
        return None

    @staticmethod
    def _split_mul(v: Value):
        """MUL(scatterC(x), w) in either order -> (x_value, w_value) or None."""
        if v.kind != "edge_expr" or v.op != "mul" or v.forced:
            return None
        a, b = v.args
        for wide, other in ((a, b), (b, a)):
            if wide.kind == "scatter" and wide.side == "C" and not wide.forced and other.width <= wide.width:
                return wide.args[0], other
        return None

    # -- forcing
    def force(self, v: Value, epilogue: int = _cabi.EPI_NONE) -> torch.Tensor:
        if v.forced:
            if epilogue != _cabi.EPI_NONE:
                raise ExecutionError("internal: epilogue requested on a forced value")
            return v.tensor
        k = kernels
        if v.kind == "mm":
            x = k.to_table(self.force(v.args[0]))
            zt = self.o["feature_dtype"] if v.pos in self.o["gathered"] and v.pos not in self.o["wanted"] else torch.float32
            v.tensor = k.gemm(x, v.weight, z_dtype=zt)
            self.kernel_log.append(("gta_gemm_f32" if zt == torch.float32 else "gta_gemm_f32_zbf16", v.pos))
        elif v.kind == "edge_mm":
            self._guard(v.args[0].width, f"the input of edge COMP_MM op {v.pos}")
            self._guard(v.width, f"edge COMP_MM op {v.pos}")
            v.tensor = k.gemm(k.to_table(self.force(v.args[0])), v.weight)
            self.kernel_log.append(("gta_gemm_f32:edges", v.pos))
        elif v.kind == "scatter":
            self._guard(v.width, f"scatter op {v.pos}")
            t, kind = self._operand(v)
            v.tensor = k.edge_unary(self.g, _cabi.UN_COPY, t, kind)
            self.kernel_log.append(("gta_edge_unary_f32:copy", v.pos))
        elif v.kind == "edge_expr":
            v.tensor = self._force_edge_expr(v)
        elif v.kind == "gather":
            v.tensor = self._force_gather(v, epilogue)
            return v.tensor
        elif v.kind == "node_expr":
            v.tensor = self._force_node_expr(v)
        else:
            raise ExecutionError(f"cannot force value kind {v.kind}")
        if epilogue != _cabi.EPI_NONE:
            raise ExecutionError("internal: epilogue requested on a non-gather value")
        return v.tensor

    def _force_edge_expr(self, v: Value) -> torch.Tensor:
        k = kernels
        sm = self._is_softmax_numerator(v)
        if sm is not None:
            el, er = self.force(sm[0]), self.o["source_table"](self.force(sm[1]))
            p, rowmax, rowsum = k.gat_logits(self.g, el, er, self.o["slope"], self.o["stabilize"])
            v.extra["rowsum"], v.extra["rowmax"] = rowsum, rowmax
            self.kernel_log.append(("gta_gat_logits_f32", v.pos))
            return p
        self._guard(v.width, f"applyedge op {v.pos}")
        if v.op in ("exp_leaky_relu", "elu", "relu"):
            t, kind = self._operand(v.args[0])
            code = {"exp_leaky_relu": _cabi.UN_EXP_LEAKY_RELU, "elu": _cabi.UN_ELU, "relu": _cabi.UN_RELU}[v.op]
            self.kernel_log.append(("gta_edge_unary_f32", v.pos))
            return k.edge_unary(self.g, code, t, kind, self.o["slope"])
        a, b = v.args
        if v.op == "rdiv":
            a, b = b, a
        ta, ka = self._operand(a)
        tb, kb = self._operand(b)
        code = {"add": _cabi.BIN_ADD, "mul": _cabi.BIN_MUL, "div": _cabi.BIN_DIV, "rdiv": _cabi.BIN_DIV}[v.op]
        self.kernel_log.append(("gta_edge_binary_f32", v.pos))
        return k.edge_binary(self.g, code, ta, ka, tb, kb)

    def _single_pass_head_width(self, f: int, heads: int) -> bool:
        """Shapes the single-pass GAT kernels take: per-head width of whole 4-feature pieces, or -- fp32 tables of
        whole pieces only -- heads of 1 or 2 features (layer 3 of the reference's GAT, F = H = 16:
        ``genGraphOP.py:31-32,50``).  Anything else runs on the generic kernels."""
        width = f // heads
        return width % 4 == 0 or (width in (1, 2) and f % 4 == 0 and self.o["feature_dtype"] == torch.float32)

    def _gat_single_pass(self, numer: Value, z_value: Value, epilogue: int, pos: int):
        sm = self._is_softmax_numerator(numer)
        el = self.force(sm[0])
        ex = self.o["source_table"]
        self.kernel_log.append(("gta_gat_aggregate_f32", pos))
        if hasattr(ex, "gather_pair"):
            # partitioned run: [z | er] travel in one gathered table.  NCCL exchange: complete on return (gate is
            # None).  Fused exchange: only this rank's slot is there yet; the aggregation launch pulls the rest
            z, er, gate = ex.gather_pair(kernels.to_table(self.force(z_value)), self.force(sm[1]))
            return kernels.gat_aggregate(self.g, el, er, z, self.o["slope"], epilogue, exchange=gate)
        er = ex(self.force(sm[1]))
        z = ex(kernels.to_table(self.force(z_value)))
        return kernels.gat_aggregate(self.g, el, er, z, self.o["slope"], epilogue)

    def _gather_sources(self, x: torch.Tensor):
        """(table, gate) of a source-side table for the fused aggregate kernels (gate: see dist.Gate)."""
        ex = self.o["source_table"]
        if hasattr(ex, "gather_one") and x.shape[0] == ex.part.rows:
            return ex.gather_one(x)
        return ex(x), None

    def _scatter_sum_leaves(self, v: Value):
        """The leaves of an ADD tree over edges whose leaves are all virtual scatters of one width (DGN's whole edge
        phase after the MM has been distributed), or None.  Inner nodes must be lazy, private to the tree and not
        asked for as outputs -- otherwise they have to exist as edge tensors anyway."""
        if v.kind != "edge_expr" or v.op != "add" or v.forced:
            return None
        leaves, stack = [], list(v.args)
        while stack:
            a = stack.pop()
            if a.kind == "scatter" and not a.forced:
                leaves.append(a)
            elif (a.kind == "edge_expr" and a.op == "add" and not a.forced and a.extra.get("consumers", 1) == 1
                  and a.pos not in self.o["wanted"]):
                stack.extend(a.args)
            else:
                return None
        if len(leaves) < 2 or len({l.width for l in leaves}) != 1:
            return None
        return leaves

    def _gather_scatter_sum(self, leaves, epilogue: int, pos: int) -> torch.Tensor:
        """gather(sum of scatters): the sources' part is ONE plain segment sum over the added node tables, the
        destinations' part is the row's own value times its degree -- no edge tensor at all (the generic path makes
        five passes over E x F for DGN)."""
        k = kernels
        total = lambda vals: self._sum_node_tables(vals, pos)
        by_src = [l for l in leaves if l.side == "C"]
        by_dst = [l for l in leaves if l.side == "R"]
        out = None
        if by_src:
            x, gate = self._gather_sources(total(by_src))
            self.kernel_log.append(("gta_aggregate_f32:scatter_sum", pos))
            out = k.aggregate(self.g, x, None, None, _cabi.EPI_NONE if by_dst else epilogue, exchange=gate)
            if not by_dst:
                return out
        if "degree" not in self.g.schedules:      # graph metadata, once: the row's edge count as an [N, 1] table
            deg = (self.g.indptr[1:] - self.g.indptr[:-1]).to(torch.float32)[:, None]
            self.g.schedules["degree"] = k.to_table(deg.contiguous())
        self.kernel_log.append(("gta_node_binary_f32", pos))
        own = k.node_binary(_cabi.BIN_MUL, total(by_dst), self.g.schedules["degree"])
        if out is not None:
            self.kernel_log.append(("gta_node_binary_f32", pos))
            own = k.node_binary(_cabi.BIN_ADD, out, own)
        if epilogue != _cabi.EPI_NONE:
            self.kernel_log.append(("gta_node_unary_f32", pos))
            own = k.node_unary(_cabi.UN_ELU if epilogue == _cabi.EPI_ELU else _cabi.UN_RELU, own, self.o["slope"])
        return own

    def _sum_node_tables(self, scatters, pos: int):
        """The node tables behind a list of scatters, added up (None for an empty list)."""
        t = None
        for val in scatters:
            cur = kernels.to_table(self.force(val.args[0]))
            if t is None:
                t = cur
            else:
                self.kernel_log.append(("gta_node_binary_f32", pos))
                t = kernels.node_binary(_cabi.BIN_ADD, t, cur)
        return t

    _EDGE_SUM_UNARY = {"relu": _cabi.UN_RELU, "elu": _cabi.UN_ELU}

    def _edge_sum_terms(self, v: Value):
        """The operand of a gather_R as ``unary(edge + sum of scatters)`` (PNA ops 5-7, ``genGraphOP.py:110-147``):
        ``(unary code, C-side scatters, R-side scatters, edge value or None)``, or None when it is not of that form.
        Every inner node must be lazy, private and not asked for as an output; a pure sum of scatters without a unary
        is ``_scatter_sum_leaves`` (cheaper: the destinations' part needs no pass over the edges at all)."""
        def private(a):
            return not a.forced and a.extra.get("consumers", 1) == 1 and a.pos not in self.o["wanted"]
        unary = _cabi.UN_COPY
        if v.kind == "edge_expr" and v.op in self._EDGE_SUM_UNARY and private(v):
            unary, v = self._EDGE_SUM_UNARY[v.op], v.args[0]
        if v.kind != "edge_expr" or v.op != "add" or not private(v):
            return None
        by_src, by_dst, edges, stack = [], [], [], list(v.args)
        while stack:
            a = stack.pop()
            if a.kind == "scatter" and not a.forced:
                (by_src if a.side == "C" else by_dst).append(a)
            elif a.kind == "edge_expr" and a.op == "add" and private(a):
                stack.extend(a.args)
            else:
                edges.append(a)
        if len(edges) > 1 or not (by_src or by_dst) or (unary == _cabi.UN_COPY and not edges):
            return None
        if len({t.width for t in by_src + by_dst + edges}) != 1:
            return None
        return unary, by_src, by_dst, (edges[0] if edges else None)

    def _gather_edge_sum(self, terms, epilogue: int, pos: int) -> torch.Tensor:
        """gather_R(unary(edge + scatterC(a) + scatterR(b))) in ONE pass over the edges: the E x F operand (if any) is
        read once, the three E x F intermediates of the generic path (two adds, the unary) never exist."""
        unary, by_src, by_dst, edge = terms
        k = kernels
        et = k.to_table(self.force(edge)) if edge is not None else None
        x = self._sum_node_tables(by_src, pos)
        if x is not None:
            x = self.o["source_table"](x)
        r = self._sum_node_tables(by_dst, pos)
        self.kernel_log.append(("gta_aggregate_edge_sum_f32", pos))
        return k.aggregate_edge_sum(self.g, et, x, r, unary, self.o["slope"], epilogue)

    def _by_source(self) -> DeviceGraph:
        """The CSC walk as a graph over EDGE ids: row j lists the CSR positions of the edges whose
        source is j, ascending (= ascending destination), so an ORDER C gather is the same
        deterministic segment sum, gathering rows of the edge tensor instead of rows of a node table."""
        if "by_source" not in self.g.schedules:
            e = self.g.num_edges
            ids = torch.arange(max(e, 1), dtype=torch.int32, device=self.g.indices.device)[:e]
            t = csr_from_coo(self.g.indices, ids, self.g.num_sources or self.g.num_nodes)
            self.g.schedules["by_source"] = DeviceGraph(t.num_nodes, e, t.indptr, t.indices, num_sources=max(e, 1))
        return self.g.schedules["by_source"]

    def _force_gather(self, v: Value, epilogue: int) -> torch.Tensor:
        k = kernels
        src = v.args[0]
        if v.side == "C":
            et = k.to_table(self.force(src))
            self.kernel_log.append(("gta_aggregate_f32:by_source", v.pos))
            return k.aggregate(self._by_source(), et, None, None, epilogue)
        # S = gather(p) where p came out of the logits kernel: the row sums are already there
        if src.forced and "rowsum" in src.extra and epilogue == _cabi.EPI_NONE:
            return src.extra["rowsum"]
        if src.kind == "scatter" and src.side == "C" and not src.forced:
            x, gate = self._gather_sources(k.to_table(self.force(src.args[0])))
            self.kernel_log.append(("gta_aggregate_f32:sum", v.pos))
            return k.aggregate(self.g, x, None, None, epilogue, exchange=gate)
        leaves = self._scatter_sum_leaves(src)
        if leaves is not None:
            return self._gather_scatter_sum(leaves, epilogue, v.pos)
        terms = self._edge_sum_terms(src)
        if terms is not None:
            return self._gather_edge_sum(terms, epilogue, v.pos)
        sp = self._split_mul(src)
        if sp is not None:
            xv, wv = sp
            # w = p / scatterR(gather(p)) with p still lazy  ->  the whole GAT edge phase, one pass
            if (wv.kind == "edge_expr" and wv.op == "div" and not wv.forced):
                p, den = wv.args
                if (den.kind == "scatter" and den.side == "R" and den.args[0].kind == "gather"
                        and den.args[0].side == "R" and den.args[0].args[0] is p and not p.forced and not den.args[0].forced
                        and self._is_softmax_numerator(p) is not None and xv.width % p.width == 0
                        and self._single_pass_head_width(xv.width, p.width)):
                    return self._gat_single_pass(p, xv, epilogue, v.pos)
                if den.kind == "scatter" and den.side == "R" and p.width == den.width:
                    x = self.o["source_table"](k.to_table(self.force(xv)))
                    pt = self.force(p)
                    dt = self.force(den.args[0])
                    if pt.shape[1] == 1 or (x.shape[1] // pt.shape[1]) % 4 == 0:
                        self.kernel_log.append(("gta_aggregate_f32:w/rowden", v.pos))
                        return k.aggregate(self.g, x, pt, dt, epilogue)
            wt = self.force(wv)
            xl = k.to_table(self.force(xv))
            if wt.dim() == 1 or wt.shape[1] == 1 or (xl.shape[1] // wt.shape[1]) % 4 == 0:
                x, gate = self._gather_sources(xl)
                self.kernel_log.append(("gta_aggregate_f32:w", v.pos))
                return k.aggregate(self.g, x, wt, None, epilogue, exchange=gate)
        if (src.kind == "edge_mm" and not src.forced and src.extra.get("consumers", 1) == 1
                and src.pos not in self.o["wanted"]):
            # COMP_MM_COMP_ADD (hardware_info.yaml:27-30): sum_k (e_k W) = (sum_k e_k) W -- reduce first, then the
            # GEMM runs over N rows instead of E and the E x Fout tensor never exists
            inner = Value("gather", args=(src.args[0],), side="R", width=src.args[0].width, pos=v.pos,
                          extra={"consumers": 1})
            reduced = k.to_table(self._force_gather(inner, _cabi.EPI_NONE))
            self.kernel_log.append(("gta_gemm_f32:after_gather", src.pos))
            out = k.gemm(reduced, src.weight)
            if epilogue != _cabi.EPI_NONE:
                self.kernel_log.append(("gta_node_unary_f32", v.pos))
                out = k.node_unary(_cabi.UN_ELU if epilogue == _cabi.EPI_ELU else _cabi.UN_RELU, out, self.o["slope"])
            return out
        # generic: materialise the edge tensor and segment-sum it (identity gather)
        et = k.to_table(self.force(src))
        if "arange" not in self.g.schedules:
            self.g.schedules["arange"] = torch.arange(max(self.g.num_edges, 1), dtype=torch.int32, device=et.device)
        ident = DeviceGraph(self.g.num_nodes, self.g.num_edges, self.g.indptr, self.g.schedules["arange"],
                            schedules=self.g.schedules)
        self.kernel_log.append(("gta_aggregate_f32:segment_sum", v.pos))
        return k.aggregate(ident, et, None, None, epilogue, sched=self.g.schedule())

    def _force_node_expr(self, v: Value) -> torch.Tensor:
        k = kernels
        if v.op in ("elu", "relu", "exp_leaky_relu"):
            a = v.args[0]
            if (a.kind == "gather" and not a.forced and v.op in ("elu", "relu") and a.extra.get("consumers", 1) == 1
                    and a.pos not in self.o["wanted"]):
                a.extra["absorbed"] = True
                return self.force(a, _cabi.EPI_ELU if v.op == "elu" else _cabi.EPI_RELU)
            if (a.kind == "node_expr" and not a.forced and v.op in ("elu", "relu") and a.extra.get("consumers", 1) == 1
                    and a.pos not in self.o["wanted"]):
                fused = self._try_gat_trans(a, _cabi.EPI_ELU if v.op == "elu" else _cabi.EPI_RELU)
                if fused is not None:
                    return fused
            code = {"elu": _cabi.UN_ELU, "relu": _cabi.UN_RELU, "exp_leaky_relu": _cabi.UN_EXP_LEAKY_RELU}[v.op]
            self.kernel_log.append(("gta_node_unary_f32", v.pos))
            return k.node_unary(code, k.to_table(self.force(a)), self.o["slope"])
        fused = self._try_gat_trans(v, _cabi.EPI_NONE)
        if fused is not None:
            return fused
        a, b = v.args
        if v.op == "rdiv":
            a, b = b, a
        code = {"add": _cabi.BIN_ADD, "mul": _cabi.BIN_MUL, "div": _cabi.BIN_DIV, "rdiv": _cabi.BIN_DIV}[v.op]
        ta, tb = k.to_table(self.force(a)), k.to_table(self.force(b))
        self.kernel_log.append(("gta_node_binary_f32", v.pos))
        return k.node_binary(code, ta, tb)

    def _try_gat_trans(self, v: Value, epilogue: int):
        """GAT-trans op 11: O = gather(p (x) Z[src]) / gather(p) with p still lazy."""
        if v.kind != "node_expr" or v.op not in ("div", "rdiv") or v.forced:
            return None
        num, den = v.args if v.op == "div" else (v.args[1], v.args[0])
        if num.kind != "gather" or den.kind != "gather" or num.forced or den.forced or "C" in (num.side, den.side):
            return None
        sp = self._split_mul(num.args[0])
        if sp is None:
            return None
        xv, wv = sp
        if wv is not den.args[0] or wv.forced or self._is_softmax_numerator(wv) is None:
            return None
        if xv.width % wv.width or not self._single_pass_head_width(xv.width, wv.width):
            return None
        return self._gat_single_pass(wv, xv, epilogue, v.pos)


def _width(nbytes: int) -> int:
    if nbytes % 4:
        raise IsaError(f"size_per_feature {nbytes} is not a whole number of fp32 elements")
    return nbytes // 4


def execute(program, op_info, graph: DeviceGraph, node_inputs: dict, weights: dict, edge_inputs: dict | None = None,
            network: str | None = None, is_reorder: bool = False, semantics: dict | None = None,
            fuse_across_blocks: bool = True, stabilize: bool = True, slope: float = kernels.LEAKY_SLOPE,
            max_edge_bytes: int = 8 << 30, outputs=None, source_table=None, return_log: bool = False,
            check_shapes: bool = True, legacy_comp_types=None, feature_dtype=torch.float32):
    """Run an ISA program functionally.

    program      : isa.Program, a path to ``Results/Insts/*.yaml`` or the raw list interpret() built
    op_info      : the op-graph list (``Network/.../*.yaml``) or a path to it
    graph        : DeviceGraph (CSR by destination)
    node_inputs  : {op position: [N, F] tensor} for ops with an external graph input
    weights      : {op position: [Fin, Fout] tensor} for COMP_MM ops
    edge_inputs  : {op position: [E] or [E, w] tensor} for '-1' entries of input_g_list
    outputs      : op positions to return (default: ops with an empty output_list)
    legacy_comp_types : COMP_TYPE per op position for V1/V2-era op graphs that lack the field
                   (V2/simpletest.yaml: ``isa.LEGACY_SIMPLETEST_COMP_TYPES``)
    feature_dtype : ``torch.bfloat16`` = bf16 STORAGE MODE: the output of a COMP_MM that is
                   gathered over the edges (Z) is stored in bf16 and accumulated in fp32 by the fused aggregate
                   kernels; tolerance rtol 2e-2, atol 1e-2 rowscale.  Fused plans only; across GPUs with the fused
                   exchange, whose tables then hold (and move) bf16 rows.
    Returns {op position: tensor} (and the kernel log with ``return_log``).
    """
    if isinstance(op_info, (str, os.PathLike)):
        op_info = read_yaml(op_info)
    if isinstance(program, (str, os.PathLike)):
        program = Program.load(program)
    elif isinstance(program, list):
        program = Program.from_records(program)
    if legacy_comp_types is not None:
        op_info = stamp_comp_types(op_info, legacy_comp_types)
    edge_inputs = edge_inputs or {}
    sem = dict(NETWORK_SEMANTICS.get((network, bool(is_reorder)), {}))
    sem.update(semantics or {})
    validate_op_graph(op_info)
    for pos, op in enumerate(op_info):
        if check_shapes:
            want = graph.num_edges if op["TYPE"] in ("applyedge", "gather") else graph.num_nodes
            for cnt in op["INPUT"]["feature_number"]:
                if cnt != want:
                    raise ExecutionError(f"op {pos} was generated for {cnt} {'edges' if want == graph.num_edges else 'nodes'}"
                                         f" but the graph has {want} (pass check_shapes=False to run it anyway)")
    prods = repair_op_graph(op_info, network, is_reorder)
    block_ops = program.block_ops(op_info)
    stored = program.stored_ops()
    n_ops = len(op_info)
    for ops in stored:
        for p in ops:
            if not 0 <= p < n_ops:
                raise IsaError(f"a STORE_* names op {p}, the op graph has {n_ops}")
    finals = [p for p in range(n_ops) if not op_info[p]["OUTPUT"]["output_list"]]
    wanted = list(outputs) if outputs is not None else finals

    # block order: dependency order, ties by program order
    owner = {p: b for b, ops in enumerate(block_ops) for p in ops}
    deps = [set() for _ in block_ops]
    for p in range(n_ops):
        for q in prods[p]:
            if q != -1 and owner[q] != owner[p]:
                deps[owner[p]].add(owner[q])
    order, done = [], set()
    while len(order) < len(block_ops):
        ready = [b for b in range(len(block_ops)) if b not in done and deps[b] <= done]
        if not ready:
            raise ExecutionError("the blocks of the program depend on each other cyclically")
        order.append(ready[0])
        done.add(ready[0])

    consumers = {p: 0 for p in range(n_ops)}
    for p in range(n_ops):
        for q in prods[p]:
            if q != -1:
                consumers[q] += 1

    if feature_dtype not in (torch.float32, torch.bfloat16):
        raise ExecutionError("feature_dtype is torch.float32 or torch.bfloat16")
    if feature_dtype == torch.bfloat16 and source_table is not None and not getattr(source_table, "fused", False):
        raise _cabi.GtaUnsupported(_cabi.ERR_UNSUPPORTED, "execute",
                                   "the bf16 storage mode needs the fused exchange across GPUs (dist.FusedExchange)")
    # ops whose output is gathered over the edges (consumed by a scatter): the tables the storage mode applies to
    gathered = {q for p in range(n_ops) if op_info[p]["TYPE"] == "scatter" for q in prods[p] if q != -1}
    opts = {"slope": float(slope), "stabilize": bool(stabilize), "max_edge_bytes": int(max_edge_bytes),
            "source_table": source_table or (lambda t: t), "wanted": set(wanted), "feature_dtype": feature_dtype,
            "gathered": gathered}
    run = _Run(graph, opts)
    env: dict[int, Value] = {}

    def make_value(pos: int) -> Value:
        op = op_info[pos]
        typ, comp, order_ = op["TYPE"], op["COMP_TYPE"], op["ORDER"]
        wout = _width(op["OUTPUT"]["size_per_feature"])
        args = []
        n_ext = 0
        for slot, q in enumerate(prods[pos]):
            if q == -1:
                # external ('-1') input: one tensor, or a list when the op has several (GIN op 3: [x, eps])
                table, label = (edge_inputs, "edge_inputs") if typ in ("applyedge", "gather") else (node_inputs, "node_inputs")
                if pos not in table:
                    raise ExecutionError(f"op {pos} needs an external input ({label}[{pos}])")
                t = table[pos]
                if isinstance(t, (list, tuple)):
                    if n_ext >= len(t):
                        raise ExecutionError(f"op {pos} has more external inputs than {label}[{pos}] provides")
                    t = t[n_ext]
                n_ext += 1
                t = t if t.dim() == 2 else t[:, None]
                kind = "edge" if typ in ("applyedge", "gather") else "node"
                args.append(Value(kind, tensor=t.contiguous() if kind == "edge" else t, width=int(t.shape[1]), pos=pos))
            else:
                args.append(env[q])
        on_edges = typ in ("applyedge", "gather")

        def external(what):
            table, label = (edge_inputs, "edge_inputs") if on_edges else (node_inputs, "node_inputs")
            if pos not in table:
                raise ExecutionError(f"op {pos} {what}: pass {label}[{pos}]")
            t = table[pos]
            t = t if t.dim() == 2 else t[:, None]
            return Value("edge" if on_edges else "node", tensor=t.contiguous() if on_edges else t,
                         width=int(t.shape[1]), pos=pos)

        if not prods[pos]:
            args.append(external("has no producer"))
        want_edges = typ in ("applyedge", "gather")
        for a in args:
            if a.on_edges != want_edges:
                raise IsaError(f"op {pos} ({typ}) reads {'an edge' if a.on_edges else 'a node'} tensor "
                               f"(from op {a.pos}); it needs {'edge' if want_edges else 'node'} tensors")
            if a.kind in ("node", "edge") and a.tensor is not None:
                rows = graph.num_edges if a.kind == "edge" else graph.num_rows
                if a.tensor.dim() != 2 or int(a.tensor.shape[0]) != rows:
                    raise ExecutionError(f"op {pos}: external {a.kind} input is {tuple(a.tensor.shape)}, "
                                         f"expected [{rows}, width]")
        if typ == "scatter":
            if order_ not in ("R", "C"):
                raise IsaError(f"op {pos}: ORDER {order_!r}")
            return Value("scatter", args=(args[0],), side=order_, width=args[0].width, pos=pos)
        if typ == "gather":
            if order_ not in ("R", "C"):
                raise IsaError(f"op {pos}: ORDER {order_!r}")
            if order_ == "C" and hasattr(opts["source_table"], "part"):
                raise _cabi.GtaUnsupported(_cabi.ERR_UNSUPPORTED, "execute",
                                           f"op {pos}: an ORDER C gather reduces over destinations, which a "
                                           f"destination-partitioned run does not own")
            return Value("gather", args=(args[0],), side=order_, width=args[0].width, pos=pos,
                         extra={"consumers": consumers[pos]})
        if comp == "MM":
            if pos not in weights:
                raise ExecutionError(f"op {pos} is COMP_MM: pass weights[{pos}]")
            w = weights[pos]
            if w.dim() != 2:
                raise ExecutionError(f"op {pos}: weight must be [Fin, Fout], got {tuple(w.shape)}")
            if int(w.shape[0]) != args[0].width:
                raise ExecutionError(f"op {pos}: weight is {tuple(w.shape)} but the input is {args[0].width} wide")
            if typ == "applyedge":
                a = args[0]
                if a.kind == "scatter" and not a.forced:
                    # row k of scatter(x).W is the row x[src k].W of x.W: transform the N node rows once
                    # and keep the scatter virtual (what the reference's PNA-trans graph does by hand)
                    inner = Value("mm", args=(a.args[0],), weight=w, width=int(w.shape[1]), pos=pos)
                    return Value("scatter", args=(inner,), side=a.side, width=inner.width, pos=pos)
                if (a.kind == "edge_expr" and a.op == "add" and not a.forced and consumers.get(a.pos, 1) == 1
                        and a.pos not in opts["wanted"] and len(a.args) == 2
                        and all(s_.kind == "scatter" and not s_.forced for s_ in a.args)):
                    # DGN op 3: MM(ADD(scatter(x), scatter(y)), W) = ADD(scatter(MM(x, W)), scatter(MM(y, W))) -- the
                    # product is linear, so it runs over the N node rows of each operand and no E x Fin tensor exists
                    parts = tuple(Value("scatter", side=s_.side, width=int(w.shape[1]), pos=pos,
                                        args=(Value("mm", args=(s_.args[0],), weight=w, width=int(w.shape[1]), pos=pos),))
                                  for s_ in a.args)
                    return Value("edge_expr", op="add", args=parts, width=int(w.shape[1]), pos=pos,
                                 extra={"consumers": consumers[pos]})
                return Value("edge_mm", args=(a,), weight=w, width=int(w.shape[1]), pos=pos,
                             extra={"consumers": consumers[pos]})
            return Value("mm", args=(args[0],), weight=w, width=int(w.shape[1]), pos=pos)
        kind = sem.get(pos, DEFAULT_SEMANTICS.get((typ, comp)))
        if kind is None:
            raise IsaError(f"op {pos}: no semantics for {typ} COMP_{comp}")
        if kind in ("add", "mul", "div", "rdiv") and len(args) == 1:
            # one declared input on a binary op (DGN/PNA op 9, the degree scaler, genGraphOP.py:120,132):
            # the second operand is the external tensor supplied under the op's own position
            args.append(external(f"is COMP_{comp} with one graph input; its second operand is external"))
        if kind in ("add", "mul", "div", "rdiv") and len(args) != 2:
            raise IsaError(f"op {pos}: COMP_{comp} needs two inputs, has {len(args)}")
        width = max(a.width for a in args)
        return Value("edge_expr" if typ == "applyedge" else "node_expr", op=kind, args=tuple(args), width=width,
                     pos=pos, extra={"consumers": consumers[pos]})

    def fuse_mm_chain(ops):
        """MM(x,W) feeding MM(.,Al) and MM(.,Ar) inside one block -> one gta_gemm_f32 call."""
        for p in ops:
            v = env[p]
            if v.kind != "mm" or v.forced:
                continue
            kids = [q for q in ops if env[q].kind == "mm" and env[q].args[0] is v and not env[q].forced]
            if len(kids) == 2 and env[kids[0]].width == env[kids[1]].width and env[kids[0]].width <= 16:
                x = kernels.to_table(run.force(v.args[0]))
                ex = opts["source_table"]
                views = None
                zt = opts["feature_dtype"] if p in opts["gathered"] and p not in opts["wanted"] else torch.float32
                if hasattr(ex, "local_views") and x.shape[0] == ex.part.rows:
                    # partitioned run: Z and er go straight into this rank's slot of the gathered table
                    views = ex.local_views(int(v.weight.shape[1]), env[kids[1]].width, x.device, zt)
                z, el, er = kernels.gemm(x, v.weight, env[kids[0]].weight, env[kids[1]].weight,
                                         out=views[0] if views else None, er_out=views[1] if views else None, z_dtype=zt)
                v.tensor, env[kids[0]].tensor, env[kids[1]].tensor = z, el, er
                run.kernel_log.append(("gta_gemm_f32+el/er", p))

    for b in order:
        ops = block_ops[b]
        pending = list(ops)
        guard = 0
        while pending:
            guard += 1
            if guard > len(ops) * len(ops) + 8:
                raise ExecutionError(f"block {b} has a dependency cycle")
            p = pending.pop(0)
            if any(q != -1 and q not in env for q in prods[p]):
                pending.append(p)
                continue
            env[p] = make_value(p)
        fuse_mm_chain(ops)
        for p in stored[b]:
            v = env[p]
            needed_outside = p in wanted
            if fuse_across_blocks and not needed_outside and v.kind != "mm":
                continue           # dead store unless somebody forces it later
            if v.extra.get("absorbed"):
                continue
            run.force(v)
    result = {}
    for p in wanted:
        result[p] = run.force(env[p])
    if return_log:
        return result, run.kernel_log
    return result


def execute_files(tile_size_list, dataset, network, layer, isReorder, graph: DeviceGraph, node_inputs, weights,
                  edge_inputs=None, root: str = ".", **kw):
    """Same positional arguments and CWD-relative files as the reference's
    ``simulate(tile_size_list, dataset, network, layer, isReorder, ...)`` (simulator.py:423-482):
    reads ``Network/<net>/<net>-<ds>/<net>-<map>/<net>-<layer>-<map>.yaml`` and
    ``Results/Insts/<net>-<ds>-<layer>-<map>.yaml`` below ``root``."""
    del tile_size_list   # tiling is an ASIC buffer decision; B200 kernels tile by work item
    m = "trans" if isReorder else "original"
    op_path = os.path.join(root, "Network", network, f"{network}-{dataset}", f"{network}-{m}", f"{network}-{layer}-{m}.yaml")
    isa_path = os.path.join(root, "Results", "Insts", f"{network}-{dataset}-{layer}-{m}.yaml")
    return execute(isa_path, op_path, graph, node_inputs, weights, edge_inputs, network=network,
                   is_reorder=isReorder, **kw)


class GraphedExecution:
    """A program execution captured into a CUDA graph (launch-bound regime: a Cora- or an
    8-GPU-Reddit-size layer is ~1 ms of kernels, comparable to the Python and launch overhead of
    issuing them one by one).  ``fn()`` must read only buffers that stay allocated (update them in
    place between replays) and return its output tensors; NCCL all-gathers inside are captured too.

        g = GraphedExecution(lambda: execute(program, op_info, graph, {0: x_static}, weights, ...))
        out = g.replay()          # same tensors every time, refreshed contents
    """

    def __init__(self, fn, warmup: int = 2):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):     # builds work lists / workspaces (these synchronise) outside capture
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

    def replay(self):
        self.graph.replay()
        return self.outputs
