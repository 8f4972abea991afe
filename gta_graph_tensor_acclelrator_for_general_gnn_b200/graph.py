"""Device-resident graph: CSR by destination, tile tables, partition, reorder, work list.

Host-side mirror of the reference's graph preprocessing entry points
(code/preprocessing.py:12-72 ``calculate_sparsity`` / ``cal_min_sparsity`` / ``gen_size``),
re-implemented on device through the C ABI, plus the CSR / partition / reorder steps the
north star adds.  torch owns memory and streams only; every computation is a
``libgta_b200.so`` call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _cabi

DEFAULT_CHUNK = 1024   # edges per work item (fixed => deterministic reduction shape)


def auto_chunk(num_edges: int) -> int:
    """Edges per work item.  Measured, not derived: 1024 is the best item size at every size tried.  One GPU, Reddit
    shape (114.6 M edges), kernel alone: 1024 3.91 ms, 512 4.00, 256 4.07, 128 5.83 -- an item costs about 5 us of
    exposed latency (item record, first ids, first er rows, chain state) whatever its length.  8 GPUs (14.3 M edges per
    rank, exchange inside the launch): 1024 0.946 ms, 512 0.933, 256 1.002, 128 2.54 -- the tail of a 1024-edge item at
    the end of the list is NOT what limits the small problem.  Kept as a function of the graph so
    the reduction shape stays fixed run to run; ``GTA_CHUNK`` pins another size for experiments (tools/scale.sh)."""
    import os
    if os.environ.get("GTA_CHUNK"):
        return int(os.environ["GTA_CHUNK"])
    return 1024


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gta_b200 runs on CUDA tensors only (there is no CPU fallback)")


@dataclass
class Schedule:
    """Work list of the aggregation kernels (gta_schedule_build)."""
    items: torch.Tensor      # int32 [num_items, 4]
    row_slots: torch.Tensor  # int32 [rows + 1]
    num_items: int
    num_slots: int
    chunk: int
    col_block: int
    row_begin: int
    row_end: int
    block_begin: list = None  # first item of every column block, then the item count (host ints)
    col_cuts: tuple = ()      # explicit column-block cut points (source ids) instead of a uniform col_block

    @property
    def num_blocks(self) -> int:
        return len(self.block_begin) - 1


#: a gathered table larger than this is walked in column blocks of about COL_BLOCK_BYTES so the
#: slice being gathered stays L2 resident (B200: 126 MB L2 over two dies; measured on the
#: Reddit shape: 119 MB table -> 60 % L2 hits unblocked)
COL_BLOCK_THRESHOLD = 72 << 20
COL_BLOCK_BYTES = 40 << 20


#: a (row, column block) segment should still hold a batch or two of edges, or the work list
#: degenerates into tiny items (measured: RMAT-20, mean degree 16, 13 blocks -> 3x slower)
COL_BLOCK_MIN_EDGES = 48


def column_block_rows(num_sources: int, row_bytes: int, mean_degree: float = float("inf")) -> int:
    """Source ids per column block for a table of ``num_sources`` rows of ``row_bytes`` (0 = none)."""
    total = num_sources * row_bytes
    if total <= COL_BLOCK_THRESHOLD:
        return 0
    blocks = min(-(-total // COL_BLOCK_BYTES), 32, int(mean_degree // COL_BLOCK_MIN_EDGES))
    if blocks <= 1:
        return 0
    return -(-num_sources // blocks)


@dataclass
class DeviceGraph:
    num_nodes: int
    num_edges: int
    indptr: torch.Tensor                  # int64 [rows+1]
    indices: torch.Tensor                 # int32 [E]  (source ids, ascending inside a row)
    perm: torch.Tensor | None = None      # int64 [E]  input position of CSR edge k
    num_sources: int | None = None        # rows of the source-side tables (== num_nodes unless partitioned)
    schedules: dict = field(default_factory=dict)

    @property
    def num_rows(self) -> int:
        return int(self.indptr.shape[0]) - 1

    def schedule(self, chunk: int | None = None, col_block: int = 0, col_cuts=None) -> Schedule:  # noqa: D401
        """Cached work list; ``col_block`` = source ids per column block (0 = no blocking), or ``col_cuts`` =
        explicit ascending cut points (column block b = sources in [cuts[b-1], cuts[b])); ``chunk`` = edges per
        item (default: ``auto_chunk`` of this graph)."""
        chunk = auto_chunk(self.num_edges) if chunk is None else chunk
        cuts = tuple(int(c) for c in col_cuts) if col_cuts else ()
        key = (chunk, col_block, cuts)
        if key not in self.schedules:
            self.schedules[key] = build_schedule(self.indptr, self.indices, 0, self.num_rows, self.num_edges,
                                                 self.num_sources or self.num_nodes, chunk, col_block, cuts)
        return self.schedules[key]

    def schedule_for(self, row_bytes: int, chunk: int | None = None) -> Schedule:
        """Work list whose column blocks keep a gathered table of ``row_bytes`` per source L2 resident."""
        mean_degree = self.num_edges / max(self.num_rows, 1)
        return self.schedule(chunk, column_block_rows(self.num_sources or self.num_nodes, row_bytes, mean_degree))


def csr_from_coo(dst, src, num_nodes: int, want_perm: bool = False) -> DeviceGraph:
    """COO (int32 device tensors, or numpy arrays which are uploaded) -> DeviceGraph."""
    lib = _cabi.load()
    if isinstance(dst, np.ndarray):
        dst = torch.from_numpy(np.ascontiguousarray(dst, dtype=np.int32)).cuda()
    if isinstance(src, np.ndarray):
        src = torch.from_numpy(np.ascontiguousarray(src, dtype=np.int32)).cuda()
    _require_cuda(dst, src)
    if dst.dtype != torch.int32 or src.dtype != torch.int32:
        raise TypeError("dst/src must be int32")
    if dst.shape != src.shape or dst.dim() != 1:
        raise ValueError("dst and src must be 1-D and of equal length")
    dst = dst.contiguous()
    src = src.contiguous()
    e = int(dst.shape[0])
    dev = dst.device
    indptr = torch.empty(num_nodes + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(max(e, 1), dtype=torch.int32, device=dev)[:e]
    perm = torch.empty(max(e, 1), dtype=torch.int64, device=dev)[:e] if want_perm else None
    ws_bytes = lib.gta_csr_build_workspace(e, num_nodes)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _cabi.check(lib.gta_csr_build(_cabi.ptr(dst), _cabi.ptr(src), e, num_nodes, _cabi.ptr(indptr),
                                  _cabi.ptr(indices), _cabi.ptr(perm), _cabi.ptr(ws), ws_bytes, _stream()),
                "gta_csr_build")
    return DeviceGraph(num_nodes, e, indptr, indices, perm, num_sources=num_nodes)


def read_csr_npz(file_path: str):
    """Host half of the ``.npz`` ingest: the SciPy-CSR archive the reference simulator reads
    (``read_csr_npz``, vTCAD/code/simulator.py:74-84: arrays ``data``, ``indices``, ``indptr``,
    ``shape``; ``KeyError`` when one is missing) expanded to COO.  Row = destination, column = source
    (SURVEY.md Appendix A).  Returns ``(dst int32 [E], src int32 [E], data [E], N)`` in file order."""
    with np.load(file_path) as z:
        missing = [k for k in ("data", "indices", "indptr", "shape") if k not in z.files]
        if missing:
            raise KeyError(f"The required keys are not found in the .npz file: {missing}")
        data, indices, indptr, shape = z["data"], z["indices"], z["indptr"], tuple(int(v) for v in z["shape"])
    if len(shape) != 2 or shape[0] != shape[1]:
        raise ValueError(f"{file_path}: adjacency must be square, shape is {shape}")
    n = shape[0]
    if indptr.shape[0] != n + 1 or int(indptr[0]) != 0 or int(indptr[-1]) != indices.shape[0] \
            or data.shape[0] != indices.shape[0] or np.any(np.diff(indptr) < 0):
        raise ValueError(f"{file_path}: inconsistent CSR arrays")
    if indices.size and (int(indices.min()) < 0 or int(indices.max()) >= n):
        raise ValueError(f"{file_path}: column index outside [0, {n})")
    if n >= 2**31 or indices.shape[0] >= 2**31:
        raise ValueError(f"{file_path}: int32 ids cannot hold this graph")
    dst = np.repeat(np.arange(n, dtype=np.int32), np.diff(indptr).astype(np.int64))
    return dst, indices.astype(np.int32), data, n


def csr_from_npz(file_path: str, drop_diagonal: bool = False):
    """SciPy-CSR ``.npz`` -> ``(DeviceGraph, edge values)``: the COO expansion is re-sorted on device by
    ``gta_csr_build`` (the file's rows need not have ascending columns), and the stored values come back
    as an fp32 ``[E]`` device tensor in CSR edge order (GCN's edge weight input).  ``drop_diagonal``
    removes self loops first, which is what the tile tables count (``A - diag``, preprocessing.py:20)."""
    dst, src, data, n = read_csr_npz(file_path)
    if drop_diagonal:
        keep = dst != src
        dst, src, data = dst[keep], src[keep], data[keep]
    g = csr_from_coo(dst, src, n, want_perm=True)
    values = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(g.indptr.device)
    return g, (values[g.perm] if g.num_edges else values)


def build_schedule(indptr: torch.Tensor, indices: torch.Tensor, row_begin: int, row_end: int, num_edges: int,
                   num_sources: int, chunk: int = DEFAULT_CHUNK, col_block: int = 0, col_cuts=()) -> Schedule:
    lib = _cabi.load()
    _require_cuda(indptr, indices)
    rows = row_end - row_begin
    if col_cuts:      # sized as num_cuts + 1 uniform blocks
        ws_sources, ws_block = len(col_cuts) + 1, 1
    else:
        ws_sources, ws_block = num_sources, col_block
    cap = int(lib.gta_schedule_max_items(rows, num_edges, chunk, ws_sources, ws_block))
    items = torch.empty((max(cap, 1), 4), dtype=torch.int32, device=indptr.device)
    row_slots = torch.zeros(rows + 1, dtype=torch.int32, device=indptr.device)
    ws_bytes = lib.gta_schedule_workspace(rows, ws_sources, ws_block)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=indptr.device)
    counts = (C.c_int64 * 2)()
    n_cb = int(lib.gta_schedule_col_blocks(ws_sources, ws_block))
    block_begin = (C.c_int64 * (n_cb + 1))()
    if col_cuts:
        cuts = (C.c_int64 * len(col_cuts))(*col_cuts)
        _cabi.check(lib.gta_schedule_build_cuts(_cabi.ptr(indptr), _cabi.ptr(indices), row_begin, row_end, chunk, cuts,
                                                len(col_cuts), _cabi.ptr(items), cap, _cabi.ptr(row_slots), counts,
                                                block_begin, _cabi.ptr(ws), ws_bytes, _stream()), "gta_schedule_build_cuts")
    else:
        _cabi.check(lib.gta_schedule_build(_cabi.ptr(indptr), _cabi.ptr(indices), row_begin, row_end, num_sources, chunk,
                                           col_block, _cabi.ptr(items), cap, _cabi.ptr(row_slots), counts, block_begin,
                                           _cabi.ptr(ws), ws_bytes, _stream()), "gta_schedule_build")
    n_items, n_slots = int(counts[0]), int(counts[1])
    return Schedule(items[:max(n_items, 1)], row_slots, n_items, n_slots, chunk, col_block, row_begin, row_end,
                    [int(v) for v in block_begin], tuple(col_cuts))


# ---- tile tables: calculate_sparsity / cal_min_sparsity / gen_size -------------------------

def gen_size(start: int, end: int) -> list[int]:
    """Tile-size list ``start*k`` up to the first value >= end (code/preprocessing.py:65-72)."""
    sizes = [start]
    k = 1
    while sizes[-1] < end:
        k += 1
        sizes.append(start * k)
    return sizes


def calculate_sparsity(g: DeviceGraph, row: int, col: int = 1, tile_begin: int = 0,
                       tile_end: int | None = None) -> torch.Tensor:
    """Device ``calculate_sparsity(row, col=1, ...)`` (code/preprocessing.py:12-40): int32
    ``[tiles, N]`` table of per-(row tile x 1 column) non-zero counts of ``A - diag``."""
    if col != 1:
        raise NotImplementedError("the reference only ever calls col = 1 (code/preprocessing.py:86)")
    if row <= 0:
        raise ValueError("row tile size must be positive")
    _require_simple(g)
    lib = _cabi.load()
    tiles = -(-g.num_nodes // row)
    tile_end = tiles if tile_end is None else tile_end
    out = torch.empty((max(tile_end - tile_begin, 0), g.num_nodes), dtype=torch.int32, device=g.indptr.device)
    _cabi.check(lib.gta_tile_nnz(_cabi.ptr(g.indptr), _cabi.ptr(g.indices), g.num_nodes, row, tile_begin, tile_end,
                                 _cabi.ptr(out), _stream()), "gta_tile_nnz")
    return out


def _require_simple(g: DeviceGraph) -> None:
    """The tile tables count non-zeros of a dense adjacency upstream (np.count_nonzero, preprocessing.py:37): a
    repeated (dst, src) pair is ONE entry there but one count per edge here.  Checked once per graph."""
    if "simple" not in g.schedules:
        e = g.num_edges
        dup = False
        if e > 1:
            same_src = g.indices[1:] == g.indices[:-1]
            row_start = torch.zeros(e, dtype=torch.bool, device=g.indices.device)
            starts = g.indptr[1:-1]
            row_start[starts[starts < e]] = True
            dup = bool((same_src & ~row_start[1:]).any().item())
        g.schedules["simple"] = not dup
    if not g.schedules["simple"]:
        raise ValueError("the graph has duplicate (dst, src) edges: the reference's tile tables are defined on a dense "
                         "adjacency, where a repeated pair is one non-zero; deduplicate the edge list first")


def cal_min_sparsity(g: DeviceGraph, tile_size: int, workspace_bytes: int = 256 << 20) -> int:
    """Maximum tile nnz for one tile size (code/preprocessing.py:53-63; the reference's name
    says min, its code takes the max), streamed in bounded batches of row tiles."""
    _require_simple(g)
    lib = _cabi.load()
    need = 256 + 4 * g.num_nodes
    ws_bytes = max(need, min(workspace_bytes, 256 + 4 * g.num_nodes * (-(-g.num_nodes // tile_size))))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.indptr.device)
    out = C.c_int32(0)
    _cabi.check(lib.gta_tile_nnz_max(_cabi.ptr(g.indptr), _cabi.ptr(g.indices), g.num_nodes, tile_size, _cabi.ptr(ws),
                                     ws_bytes, C.byref(out), _stream()), "gta_tile_nnz_max")
    return int(out.value)


def tile_tables(g: DeviceGraph, start: int, end: int):
    """``sizelist`` and ``maxlist`` as the preprocessing CLI writes them (preprocessing.py:83-96)."""
    sizes = gen_size(start, end)
    return sizes, [cal_min_sparsity(g, s) for s in sizes]


def write_tile_tables(g: DeviceGraph, dataset: str, sizes, root: str = ".") -> tuple:
    """Write the files the reference's preprocessing CLI writes (code/preprocessing.py:83-96), in its
    format (``yaml.dump`` of plain int lists): ``dataset/<ds>/adj_<ds>_<SR>_1.yaml`` for every tile
    size, ``sizelist_<ds>.yaml`` and ``maxlist_<ds>.yaml`` -- computed on device from the CSR, no dense
    ``N x N`` adjacency.  Returns ``(sizelist, maxlist)``; these are what ``compile()`` reads
    (vTCAD/code/compiler.py:504-505) and what ``simulate()`` reads (simulator.py:481-482)."""
    import os

    import yaml
    d = os.path.join(root, "dataset", dataset)
    os.makedirs(d, exist_ok=True)
    sizes = [int(s) for s in sizes]
    maxlist = []
    for sr in sizes:
        table = calculate_sparsity(g, sr).cpu().tolist()
        with open(os.path.join(d, f"adj_{dataset}_{sr}_1.yaml"), "w") as f:
            yaml.dump(table, f)
        maxlist.append(max((max(row) for row in table if row), default=0))
    with open(os.path.join(d, f"sizelist_{dataset}.yaml"), "w") as f:
        yaml.dump(sizes, f)
    with open(os.path.join(d, f"maxlist_{dataset}.yaml"), "w") as f:
        yaml.dump(maxlist, f)
    return sizes, maxlist


# ---- partition / reorder -----------------------------------------------------------------------

def partition_bounds(g: DeviceGraph, parts: int) -> torch.Tensor:
    """int64 [parts+1] destination-range bounds balanced by edge count."""
    lib = _cabi.load()
    bounds = torch.empty(parts + 1, dtype=torch.int64, device=g.indptr.device)
    _cabi.check(lib.gta_partition(_cabi.ptr(g.indptr), g.num_rows, parts, _cabi.ptr(bounds), _stream()),
                "gta_partition")
    return bounds


def degree_reorder(g: DeviceGraph) -> torch.Tensor:
    """int64 [N] permutation, ``perm[new] = old``, descending in-degree, stable."""
    lib = _cabi.load()
    perm = torch.empty(g.num_rows, dtype=torch.int64, device=g.indptr.device)
    ws_bytes = lib.gta_reorder_workspace(g.num_rows)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.indptr.device)
    _cabi.check(lib.gta_reorder(_cabi.ptr(g.indptr), g.num_rows, _cabi.ptr(perm), _cabi.ptr(ws), ws_bytes, _stream()),
                "gta_reorder")
    return perm


def slice_rows(g: DeviceGraph, row_begin: int, row_end: int) -> DeviceGraph:
    """Destination rows [row_begin,row_end) as a zero-based local CSR (sources keep global ids)."""
    e0 = int(g.indptr[row_begin].item())
    e1 = int(g.indptr[row_end].item())
    indptr = (g.indptr[row_begin:row_end + 1] - e0).contiguous()
    indices = g.indices[e0:e1].contiguous()
    return DeviceGraph(g.num_nodes, e1 - e0, indptr, indices, None, num_sources=g.num_sources)
