"""Deterministic synthetic graphs and tensors of the BASELINE shapes.

The reference ships no datasets; every config in BASELINE.json is a
*shape*.  These generators are pure functions of ``(shape, seed)`` so the CPU box
and the GPU box build identical inputs without shipping anything large; each
graph carries a checksum that runs log so both sides can prove it.

Shapes follow the dataset table of the reference's op-graph generator
(vTCAD/GraphOP/genGraphOP.py:184-199) plus Flickr (vTCAD/code/interpreter.py:811-819)
and the RMAT config of BASELINE.json.

Edge convention (template/ISA_defination.yaml:35): adjacency
``A[row = dst i, col = src j]``; an edge ``k = (i <- j)`` is stored as ``dst[k] = i``,
``src[k] = j``.  Graphs are symmetric, self-loop free and duplicate free, so a dense
``N x N`` adjacency (what code/preprocessing.py:12-40 consumes) represents them exactly.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import numpy as np

# name -> (nodes, directed edges, input feature width)
SHAPES = {
    "cora": (2708, 10556, 1433),
    "citeseer": (3327, 9104, 3703),
    "pubmed": (19717, 88648, 500),
    "flickr": (89250, 899756, 500),
    "reddit": (232965, 114615892, 602),
}

# hidden widths per layer: [Fin, 128, 64, 16]  (genGraphOP.py:31-32)
LAYER_WIDTH = {1: 128, 2: 64, 3: 16}


@dataclass
class CooGraph:
    """Directed edge list, unsorted, int32."""

    num_nodes: int
    dst: np.ndarray
    src: np.ndarray
    name: str = "synthetic"

    @property
    def num_edges(self) -> int:
        return int(self.dst.shape[0])

    def checksum(self) -> str:
        h = hashlib.sha256()
        h.update(np.int64(self.num_nodes).tobytes())
        h.update(np.ascontiguousarray(self.dst, dtype=np.int32).tobytes())
        h.update(np.ascontiguousarray(self.src, dtype=np.int32).tobytes())
        return h.hexdigest()[:16]


def _sorted_unique(keys: np.ndarray) -> np.ndarray:
    """Sorted distinct values (np.sort + mask; np.unique's hash path is ~20x slower at 6e7 keys)."""
    if keys.shape[0] == 0:
        return keys
    keys = np.sort(keys, kind="stable")
    keep = np.empty(keys.shape[0], dtype=bool)
    keep[0] = True
    np.not_equal(keys[1:], keys[:-1], out=keep[1:])
    return keys[keep]


def _powerlaw_endpoints(rng: np.random.Generator, n: int, num_nodes: int, gamma: float, i0: float) -> np.ndarray:
    """Sample node ranks with probability proportional to (rank + i0)^-gamma (inverse CDF)."""
    u = rng.random(n)
    e = 1.0 - gamma
    a = (num_nodes + i0) ** e
    b = i0 ** e
    x = (u * (a - b) + b) ** (1.0 / e) - i0
    r = np.floor(x).astype(np.int64)
    np.clip(r, 0, num_nodes - 1, out=r)
    return r


def powerlaw_graph(num_nodes: int, num_edges: int, seed: int = 0, gamma: float = 0.5,
                   i0: float = 100.0, name: str = "powerlaw") -> CooGraph:
    """Chung-Lu style graph: ``num_edges/2`` distinct undirected pairs, symmetrised.

    ``gamma``/``i0`` shape the degree skew; the defaults give a Reddit-shape graph a
    maximum degree of about 25x the mean (real Reddit: 21,657 vs 492).
    Node ids are shuffled so degree is not correlated with id.
    """
    if num_edges % 2:
        raise ValueError("num_edges must be even (symmetrised pairs)")
    pairs = num_edges // 2
    if pairs > num_nodes * (num_nodes - 1) // 2:
        raise ValueError("more edges than a simple graph can hold")
    rng = np.random.default_rng(seed)
    relabel = rng.permutation(num_nodes).astype(np.int64)
    keys = np.empty(0, dtype=np.int64)
    need = pairs
    while keys.shape[0] < pairs:
        m = int(need * 1.08) + 1024
        a = _powerlaw_endpoints(rng, m, num_nodes, gamma, i0)
        b = _powerlaw_endpoints(rng, m, num_nodes, gamma, i0)
        keep = a != b
        a, b = a[keep], b[keep]
        lo = np.minimum(a, b)
        hi = np.maximum(a, b)
        keys = _sorted_unique(np.concatenate([keys, lo * num_nodes + hi]))
        need = max(pairs - keys.shape[0], 0) + 1024
    if keys.shape[0] > pairs:
        drop = rng.choice(keys.shape[0], size=keys.shape[0] - pairs, replace=False)
        mask = np.ones(keys.shape[0], dtype=bool)
        mask[drop] = False
        keys = keys[mask]
    lo = relabel[keys // num_nodes]
    hi = relabel[keys % num_nodes]
    dst = np.concatenate([lo, hi]).astype(np.int32)
    src = np.concatenate([hi, lo]).astype(np.int32)
    # interleave deterministically so the edge list is NOT pre-sorted (the CSR build must sort it)
    order = rng.permutation(dst.shape[0])
    return CooGraph(num_nodes, dst[order], src[order], name)


def rmat_graph(scale: int, edge_factor: int = 16, seed: int = 0,
               abcd=(0.57, 0.19, 0.19, 0.05), name: str | None = None) -> CooGraph:
    """RMAT (Graph500 parameters) directed graph, duplicates and self-loops removed,
    topped up to exactly ``edge_factor * 2**scale`` edges (BASELINE config 5 at scale 24)."""
    n = 1 << scale
    target = edge_factor * n
    rng = np.random.default_rng(seed)
    a, b, c, _ = abcd
    keys = np.empty(0, dtype=np.int64)
    need = target
    while keys.shape[0] < target:
        m = int(need * 1.15) + 1024
        row = np.zeros(m, dtype=np.int64)
        col = np.zeros(m, dtype=np.int64)
        for _bit in range(scale):
            u = rng.random(m)
            rbit = u >= a + b            # quadrants c, d
            cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)   # quadrants b, d
            row = (row << 1) | rbit
            col = (col << 1) | cbit
        keep = row != col
        keys = _sorted_unique(np.concatenate([keys, row[keep] * n + col[keep]]))
        need = max(target - keys.shape[0], 0) + 1024
    if keys.shape[0] > target:
        drop = rng.choice(keys.shape[0], size=keys.shape[0] - target, replace=False)
        mask = np.ones(keys.shape[0], dtype=bool)
        mask[drop] = False
        keys = keys[mask]
    relabel = rng.permutation(n).astype(np.int64)
    dst = relabel[keys // n].astype(np.int32)
    src = relabel[keys % n].astype(np.int32)
    order = rng.permutation(dst.shape[0])
    return CooGraph(n, dst[order], src[order], name or f"rmat{scale}")


#: degree skew of the ``<shape>-heavy`` variants: P(rank) ~ (rank + 600)^-0.9.  At the Reddit shape the expected
#: degrees run 109 .. 23 400 around the mean of 492 (max 47.6x the mean, median 0.41x) -- the tail of the real Reddit
#: graph (max 21 657 = 44x, long low-degree tail), which the default (gamma 0.5, i0 100: 251 .. 12 121, max 24.6x,
#: median 0.72x) flatters.  A textbook alpha ~ 2.1 sequence is not realisable at this density: its head would need
#: degrees beyond N - 1 and E/2 = 57 M DISTINCT pairs cannot be drawn from it.
HEAVY_TAIL = {"gamma": 0.9, "i0": 600.0}


def shape_graph(name: str, seed: int = 0, scale: float = 1.0) -> CooGraph:
    """Graph of a named dataset shape (``<shape>-heavy``: the heavier-tailed degree sequence, HEAVY_TAIL);
    ``scale`` < 1 shrinks nodes and edges together (used for bounded CPU samples) while keeping the mean degree."""
    base, heavy = (name[:-6], True) if name.endswith("-heavy") else (name, False)
    n, e, _ = SHAPES[base]
    if scale != 1.0:
        n = max(int(n * scale), 16)
        e = max(int(e * scale) // 2 * 2, 2)
    kw = HEAVY_TAIL if heavy else {}
    return powerlaw_graph(n, e, seed=seed, name=name if scale == 1.0 else f"{name}x{scale:g}", **kw)


def degree_stats(dst: np.ndarray, num_nodes: int) -> dict:
    """min / median / mean / max in-degree of an edge list (what the bench line states about its workload)."""
    deg = np.bincount(dst, minlength=num_nodes)
    return {"min": int(deg.min()), "median": float(np.median(deg)), "mean": float(deg.mean()), "max": int(deg.max()),
            "max_over_mean": float(deg.max() / max(deg.mean(), 1e-30))}


def rmat_graph_device(scale: int, edge_factor: int = 16, seed: int = 0, abcd=(0.57, 0.19, 0.19, 0.05), device="cuda"):
    """RMAT edge list generated ON THE DEVICE (BASELINE config 5: scale 24 = 16.8 M nodes, 268 M edges -- the numpy
    generator above needs minutes and tens of GB per rank for that).  Same quadrant probabilities and the same random
    relabelling, ``torch`` CUDA generator seeded with ``seed`` so every rank of a box draws the identical list.
    Self loops are redirected to the next node id; duplicate edges are KEPT (a multigraph, which the CSR builder and
    the kernels handle; about 1 % of the edges at scale 24).  Returns ``(dst int32 [E], src int32 [E], N)``."""
    import torch
    n = 1 << scale
    m = edge_factor * n
    gen = torch.Generator(device=device).manual_seed(seed)
    a, b, c, _ = abcd
    row = torch.zeros(m, dtype=torch.int32, device=device)
    col = torch.zeros(m, dtype=torch.int32, device=device)
    for _bit in range(scale):
        u = torch.rand(m, device=device, generator=gen)
        rbit = (u >= a + b).to(torch.int32)
        cbit = (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int32)
        row = (row << 1) | rbit
        col = (col << 1) | cbit
        del u, rbit, cbit
    col = torch.where(row == col, (col + 1) % n, col)
    relabel = torch.randperm(n, device=device, generator=gen).to(torch.int32)
    dst = relabel[row.long()]
    src = relabel[col.long()]
    del row, col
    order = torch.randperm(m, device=device, generator=gen)
    return dst[order].contiguous(), src[order].contiguous(), n


def glorot(rng: np.random.Generator, fan_in: int, fan_out: int) -> np.ndarray:
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def gat_tensors(num_nodes: int, fin: int, fout: int, heads: int, seed: int = 0, dense_attention: bool = False):
    """X ~ N(0,1), W Glorot, attention weights U(-0.1, 0.1).

    ``A_l``/``A_r`` are ``[fout, heads]`` weights of GAT ops 1 and 2 (applynode MM,
    genGraphOP.py:50-51).  The default is the block-diagonal multi-head form (head h
    only sees its own ``fout/heads`` slice); ``dense_attention`` fills every entry.
    """
    rng = np.random.default_rng(seed + 1)
    x = rng.standard_normal((num_nodes, fin), dtype=np.float32)
    w = glorot(rng, fin, fout)
    al = rng.uniform(-0.1, 0.1, size=(fout, heads)).astype(np.float32)
    ar = rng.uniform(-0.1, 0.1, size=(fout, heads)).astype(np.float32)
    if not dense_attention:
        d = fout // heads
        mask = np.zeros((fout, heads), dtype=np.float32)
        for h in range(heads):
            mask[h * d:(h + 1) * d, h] = 1.0
        al *= mask
        ar *= mask
    return x, w, al, ar


def gcn_edge_norm(indptr: np.ndarray, indices: np.ndarray, rows: np.ndarray | None = None) -> np.ndarray:
    """GCN edge weight 1/sqrt(deg_i * deg_j) in CSR edge order (the '-1' external
    edge input of GCN op 1, genGraphOP.py:36)."""
    deg = np.diff(indptr).astype(np.float64)
    deg = np.maximum(deg, 1.0)
    if rows is None:
        rows = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
    return (1.0 / np.sqrt(deg[rows] * deg[indices])).astype(np.float32)


def opgraph_inputs(op_info, n: int, e: int, seed: int = 0):
    """Random fp32 tensors for every external input and weight of an op graph: ``(node_inputs, weights,
    edge_inputs)`` keyed by op position, as ``executor.execute`` takes them.  External = an op without a producer
    (or PNA-trans' self reference), a ``-1`` entry of ``input_g_list`` (edge weights in (0.05, 1]; GIN's ``[x, eps]``
    pair), and the second operand of a binary op that declares one input (DGN / PNA degree scaler)."""
    rng = np.random.default_rng(seed)
    node_inputs, weights, edge_inputs = {}, {}, {}
    for pos, op in enumerate(op_info):
        widths = [s // 4 for s in op["INPUT"]["size_per_feature"]]
        ins = op["INPUT"]["input_g_list"]
        if op["COMP_TYPE"] == "MM":
            fout = op["OUTPUT"]["size_per_feature"] // 4
            weights[pos] = glorot(rng, widths[0], fout) if fout > 16 else \
                rng.uniform(-0.1, 0.1, size=(widths[0], fout)).astype(np.float32)
        on_edges = op["TYPE"] in ("applyedge", "gather")
        if not ins or ins == [pos]:
            if on_edges:
                edge_inputs[pos] = rng.standard_normal((e, widths[0]), dtype=np.float32)
            else:
                node_inputs[pos] = rng.standard_normal((n, widths[0]), dtype=np.float32)
        elif op["COMP_TYPE"] in ("MUL", "ADD") and len(ins) == 1:
            (edge_inputs if on_edges else node_inputs)[pos] = \
                rng.uniform(0.5, 1.5, size=(e if on_edges else n, 1)).astype(np.float32)
        ext = []
        for slot, q in enumerate(ins):
            if q == -1:
                if on_edges:
                    ext.append(rng.uniform(0.05, 1.0, size=(e, 1)).astype(np.float32))
                else:
                    ext.append(rng.standard_normal((n, widths[slot]), dtype=np.float32))
        if ext:
            (edge_inputs if on_edges else node_inputs)[pos] = ext[0] if len(ext) == 1 else ext
    return node_inputs, weights, edge_inputs
