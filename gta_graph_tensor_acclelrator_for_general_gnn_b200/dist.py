"""Destination-range partitioning over the GPUs of one box.

One process per GPU (torchrun).  Rank p owns CSR rows ``[bounds[p], bounds[p+1])`` -- bounds balanced
by edge count (gta_partition) -- and the matching rows of every node tensor.  The only exchange per
layer is a replication of the source-side tables (``[Z | er]`` for GAT, ``Z`` for GCN); destination rows
are disjoint, so there is no reduction.

Two exchanges implement it:

* :class:`FusedExchange` (default): the transfer happens INSIDE the aggregation launch.  Every rank keeps
  a gathered table ``[world, stride, ld]`` whose slot k holds the rows of rank ``(rank + k) % world`` (its own
  rows first), published to the peers through CUDA IPC.  The GEMM writes slot 0, ``gta_exchange_publish``
  tells the peers, and the aggregation kernel's first CTAs pull the peers' slots over NVLink in ring order
  while the other CTAs already reduce the column block of the rank's own sources and wait, slot by slot, for
  the rest (csrc/exchange.cuh).  No NCCL call, no communication kernel, no launch boundary in the step.
* :class:`SourceExchange`: one NCCL all-gather per layer between the GEMM and the aggregation (the
  round-1 path; kept as the baseline the fused exchange is measured against, ``bench.py --exchange nccl``).

Source ids are remapped once at setup (gta_remap_sources).  With the rotated layout of the fused exchange
every row is re-sorted by the new ids: a fixed reduction order, deterministic, but not the single-GPU one.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _cabi, kernels
from .graph import DeviceGraph, _stream, csr_from_coo, partition_bounds, slice_rows

#: a slot must be a whole number of 128-byte lines (a cached line never straddles two slots, one of which
#: may not have landed yet): row pitches are multiples of 16 bytes, so 8 rows always are
SLOT_ROW_MULTIPLE = 8


@dataclass
class Partition:
    rank: int
    world: int
    bounds: list          # python ints, len world+1
    stride: int           # padded rows per rank in gathered tables (multiple of SLOT_ROW_MULTIPLE)
    local: DeviceGraph    # zero-based local CSR, sources remapped into the gathered table; .perm maps its
                          # edges to positions in the rank's slice of the global CSR edge order
    num_nodes: int
    rotate: bool = True   # slot k = rank (rank + k) % world (fused exchange); False: slot k = rank k (all-gather)

    @property
    def row_begin(self) -> int:
        return self.bounds[self.rank]

    @property
    def row_end(self) -> int:
        return self.bounds[self.rank + 1]

    @property
    def rows(self) -> int:
        return self.row_end - self.row_begin

    def slot_of(self, owner: int) -> int:
        return (owner - self.rank) % self.world if self.rotate else owner

    def owner_of(self, slot: int) -> int:
        return (self.rank + slot) % self.world if self.rotate else slot

    def permute_edges(self, t: torch.Tensor) -> torch.Tensor:
        """Edge tensor given in the global CSR order of this rank's rows -> the local graph's edge order."""
        return t if self.local.perm is None else t[self.local.perm]


def _stride_for(bounds, world: int) -> int:
    stride = max(bounds[p + 1] - bounds[p] for p in range(world))
    return max((stride + SLOT_ROW_MULTIPLE - 1) // SLOT_ROW_MULTIPLE * SLOT_ROW_MULTIPLE, SLOT_ROW_MULTIPLE)


def _localise(rows_of_edge: torch.Tensor, src: torch.Tensor, b: torch.Tensor, bounds, rank: int, world: int,
              num_nodes: int, rotate: bool) -> Partition:
    """Local CSR of one rank from its edges (local destination row, GLOBAL source id)."""
    lib = _cabi.load()
    stride = _stride_for(bounds, world)
    rows = bounds[rank + 1] - bounds[rank]
    remapped = torch.empty_like(src)
    _cabi.check(lib.gta_remap_sources(_cabi.ptr(src), int(src.shape[0]), _cabi.ptr(b), world, stride,
                                      rank if rotate else 0, _cabi.ptr(remapped), _stream()), "gta_remap_sources")
    local = csr_from_coo(rows_of_edge, remapped, rows, want_perm=True)
    local.num_nodes = num_nodes
    local.num_sources = world * stride
    return Partition(rank, world, bounds, stride, local, num_nodes, rotate)


def make_partition(full: DeviceGraph, rank: int, world: int, rotate: bool = True) -> Partition:
    """Cut a (replicated) full graph into this rank's destination range."""
    b = partition_bounds(full, world)
    bounds = [int(v) for v in b.cpu().tolist()]
    local = slice_rows(full, bounds[rank], bounds[rank + 1])
    deg = local.indptr[1:] - local.indptr[:-1]
    rows_of_edge = torch.repeat_interleave(torch.arange(local.num_rows, dtype=torch.int32, device=deg.device), deg)
    return _localise(rows_of_edge, local.indices, b, bounds, rank, world, full.num_nodes, rotate)


def partition_from_coo(dst: torch.Tensor, src: torch.Tensor, num_nodes: int, rank: int, world: int,
                       rotate: bool = True) -> Partition:
    """Partition ON BUILD: from the (replicated, device-resident) edge list straight to this rank's local CSR.
    Only a degree histogram of the whole graph is formed, never its CSR (RMAT-24: 1 GB of indices per rank
    saved, and one sort of E/world edges instead of E).  Same bounds as :func:`make_partition`."""
    lib = _cabi.load()
    deg = torch.bincount(dst, minlength=num_nodes)
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=dst.device)
    torch.cumsum(deg, 0, out=indptr[1:])
    b = torch.empty(world + 1, dtype=torch.int64, device=dst.device)
    _cabi.check(lib.gta_partition(_cabi.ptr(indptr), num_nodes, world, _cabi.ptr(b), _stream()), "gta_partition")
    bounds = [int(v) for v in b.cpu().tolist()]
    mine = (dst >= bounds[rank]) & (dst < bounds[rank + 1])
    rows_of_edge = (dst[mine] - bounds[rank]).to(torch.int32)
    return _localise(rows_of_edge, src[mine].contiguous(), b, bounds, rank, world, num_nodes, rotate)


class SourceExchange:
    """NCCL all-gather of a local ``[rows, F]`` table into the gathered ``[world, stride, F]`` table (slot k =
    rank k: needs a partition built with ``rotate=False``).  Buffers are cached per width, so the steady state
    allocates nothing."""

    fused = False

    def __init__(self, part: Partition, group=None):
        if part.rotate and part.world > 1:
            raise ValueError("the NCCL all-gather lays slots out in rank order: build the partition with rotate=False")
        self.part = part
        self.group = group
        self._buf = {}

    # -- layout ---------------------------------------------------------------------------------
    def buffer(self, width: int, device) -> torch.Tensor:
        key = (width, torch.device(device))
        if key not in self._buf:
            p = self.part
            self._buf[key] = torch.zeros((p.world, p.stride, kernels.pad4(width)), dtype=torch.float32, device=device)
        return self._buf[key]

    def local_views(self, f: int, h: int, device, dtype=torch.float32):
        """``(z [rows, f], er [rows, h])`` views of this rank's slot of the gathered ``[F | H]`` table, for
        producers (the GEMM) that write there directly."""
        if dtype != torch.float32:
            raise _cabi.GtaUnsupported(_cabi.ERR_UNSUPPORTED, "SourceExchange", "the NCCL all-gather exchange is fp32 only")
        p = self.part
        slot = self.buffer(f + kernels.pad4(h), device)[p.slot_of(p.rank), :p.rows]
        return slot[:, :f], slot[:, f:f + h]

    def _store(self, t: torch.Tensor, width: int, col: int) -> None:
        p = self.part
        self.buffer(width, t.device)[p.slot_of(p.rank), :p.rows, col:col + int(t.shape[1])].copy_(t)

    # -- collective -------------------------------------------------------------------------------
    def _gather(self, width: int, device) -> torch.Tensor:
        p = self.part
        buf = self.buffer(width, device)
        if p.world > 1:
            dist.all_gather_into_tensor(buf.view(-1, buf.shape[-1]), buf[p.rank], group=self.group)
        return buf.view(-1, buf.shape[-1])[:, :width]

    @kernels._timed("nccl_all_gather")
    def gather_pair(self, z: torch.Tensor, er: torch.Tensor):
        """One gathered table for the two source-side tensors of a GAT layer: every source row is
        ``[z (F) | er (H)]``.  Returns ``(z_view, er_view, gate)`` (strided views; gate is None: complete)."""
        f, h = int(z.shape[1]), int(er.shape[1])
        width = f + kernels.pad4(h)
        views = self.local_views(f, h, z.device)
        if not (views[0].data_ptr() == z.data_ptr() and views[1].data_ptr() == er.data_ptr()):
            self._store(z, width, 0)          # the producer did not write into the slot: copy
            self._store(er, width, f)
        full = self._gather(width, z.device)
        return full[:, :f], full[:, f:f + h], None

    @kernels._timed("nccl_all_gather")
    def gather_one(self, t: torch.Tensor):
        """``(table, gate)`` for a single source-side tensor (GCN: Z)."""
        width = int(t.shape[1])
        self._store(t, width, 0)
        return self._gather(width, t.device), None

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        """Executor ``source_table`` hook for consumers that need the whole table at once."""
        if t.shape[0] != self.part.rows:
            return t            # already a gathered table
        return self.gather_one(t)[0]

    def last_table(self, width: int) -> torch.Tensor:
        """The gathered ``[world*stride, width]`` table of the last gather_* call of that width."""
        buf = self._buf[next(k for k in self._buf if k[0] == width)]
        return buf.view(-1, buf.shape[-1])[:, :width]


class _IpcBuffer:
    """Buffer cudaMalloc'ed by the library (so its CUDA IPC handle names a base pointer), exposed to torch
    through ``__cuda_array_interface__`` as fp32."""

    def __init__(self, shape):
        self.lib = _cabi.load()
        self.shape = tuple(int(v) for v in shape)
        n = 4
        for v in self.shape:
            n *= v
        ptr = C.c_void_p()
        _cabi.check(self.lib.gta_ipc_alloc(n, C.byref(ptr)), "gta_ipc_alloc")
        self.ptr = int(ptr.value)
        self.nbytes = n
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self) -> torch.Tensor:
        return torch.as_tensor(self, device=torch.device("cuda", torch.cuda.current_device()))

    def handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _cabi.check(self.lib.gta_ipc_export(self.ptr, buf), "gta_ipc_export")
        return buf.raw

    def close(self):
        if self.ptr:
            self.lib.gta_ipc_free(self.ptr)
            self.ptr = 0


def slot_groups(world: int) -> list:
    """Column blocks of a fused-exchange work list, as lists of slots: the rank's own slot alone (work that needs
    no transfer), then the peers' slots in ring order in two groups, the first about half the size of the second
    (it only has to outlast the arrival of the second).  Per-slot blocks were measured on 8 B200: a Reddit-shape
    row became 8 items of 61 edges and the per-item overhead cost 47 % (profiles/r02_bench_n8_fused_per_slot_blocks)."""
    if world <= 1:
        return [[0]]
    first = max((world - 1) * 3 // 7, 1) if world > 2 else 1
    groups = [[0], list(range(1, 1 + first))]
    if 1 + first < world:
        groups.append(list(range(1 + first, world)))
    return groups


class Gate:
    """What an aggregation launch needs to pull the peers' slots itself: the ``gta_exchange_t`` of one step."""

    def __init__(self, struct: _cabi.Exchange, keep=()):
        self.struct = struct
        self._keep = keep

    @property
    def slot_rows(self) -> int:
        return int(self.struct.slot_rows)

    @property
    def block_cuts(self) -> tuple:
        """Cut points (source ids) of the work list's column blocks: the ends of all slot groups but the last."""
        return tuple((g[-1] + 1) * self.slot_rows for g in slot_groups(int(self.struct.world))[:-1])

    def byref(self):
        return C.byref(self.struct)


class FusedExchange:
    """In-kernel exchange (module docstring).  ``peers`` (tests, single-GPU emulation of every rank): the other
    ranks' FusedExchange objects of this process; otherwise the peers are the ranks of ``group`` and the tables
    and signal blocks are mapped through CUDA IPC.  Raises ``RuntimeError`` on every rank if any rank cannot map
    its peers (the caller falls back to the NCCL all-gather)."""

    fused = True

    def __init__(self, part: Partition, group=None, copy_ctas: int = 0):
        if not part.rotate and part.world > 1:
            raise ValueError("the fused exchange keeps the rank's own rows in slot 0: build the partition with rotate=True")
        if part.world > _cabi.MAX_RANKS:
            raise ValueError(f"at most {_cabi.MAX_RANKS} ranks")
        self.part = part
        self.group = group
        self.copy_ctas = copy_ctas
        self.step = 0
        self._state = {}
        self._emulated = None

    def emulate_with(self, peers) -> None:
        """Single-process emulation: ``peers[q]`` is rank q's FusedExchange (this object at [rank])."""
        self._emulated = list(peers)

    # -- setup: tables (two step parities) + signal block, published to / mapped from the peers ------------
    @staticmethod
    def _row_bytes(f: int, h: int, dtype) -> int:
        """Bytes of one table row: ``f`` features of ``dtype`` padded to 16 bytes, then ``h`` fp32 (er) padded to 16."""
        es = 2 if dtype == torch.bfloat16 else 4
        return (f * es + 15) // 16 * 16 + kernels.pad4(h) * 4

    def _alloc(self, row_bytes: int, device):
        p = self.part
        lib = _cabi.load()
        st = {"row_bytes": row_bytes, "tables": [_IpcBuffer((p.world, p.stride, row_bytes // 4)) for _ in range(2)],
              "signals": _IpcBuffer(((int(lib.gta_exchange_signal_bytes()) + 3) // 4,))}
        st["tables_t"] = [t.tensor() for t in st["tables"]]          # fp32-typed raw rows [world, stride, row_bytes / 4]
        return st

    def _views(self, st, parity: int, f: int, h: int, dtype):
        """``(z [world*stride, f] of dtype, er [world*stride, h] fp32 | None)`` strided views of one table."""
        raw = st["tables_t"][parity].view(-1, st["row_bytes"] // 4)
        es = 2 if dtype == torch.bfloat16 else 4
        z = (raw.view(torch.bfloat16) if dtype == torch.bfloat16 else raw)[:, :f]
        off = ((f * es + 15) // 16 * 16) // 4
        return z, (raw[:, off:off + h] if h else None)

    def _setup(self, row_bytes: int, device):
        key = (row_bytes, torch.device(device))
        if key in self._state:
            return self._state[key]
        p = self.part
        lib = _cabi.load()
        if self._emulated is not None or p.world == 1:
            st = self._alloc(row_bytes, device)
            st["own"] = True
            self._state[key] = st
            return st
        ok, err, st = 1, "", None
        try:
            st = self._alloc(row_bytes, device)
            handles = [t.handle() for t in st["tables"]] + [st["signals"].handle()]
        except Exception as exc:          # keep going to the collective below so every rank agrees
            ok, err, handles = 0, str(exc), [b"", b"", b""]
        gathered = [None] * p.world
        dist.all_gather_object(gathered, handles, group=self.group)
        peer = [[0, 0, 0] for _ in range(p.world)]
        if ok:
            try:
                for q in range(p.world):
                    for i in range(3):
                        if q == p.rank:
                            peer[q][i] = (st["tables"] + [st["signals"]])[i].ptr
                        else:
                            mapped = C.c_void_p()
                            _cabi.check(lib.gta_ipc_open(gathered[q][i], C.byref(mapped)), "gta_ipc_open")
                            peer[q][i] = int(mapped.value)
            except Exception as exc:
                ok, err = 0, str(exc)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            raise RuntimeError("fused exchange unavailable on at least one rank" + (": " + err if err else ""))
        st["peer"] = peer
        self._state[key] = st
        return st

    def _peer_ptrs(self, row_bytes: int, device, parity: int):
        """(table pointer of every rank for this parity, signal pointer of every rank)."""
        p = self.part
        st = self._setup(row_bytes, device)
        if self._emulated is not None:
            others = [pe._setup(row_bytes, device) for pe in self._emulated]
            return [o["tables"][parity].ptr for o in others], [o["signals"].ptr for o in others]
        if p.world == 1:
            return [st["tables"][parity].ptr], [st["signals"].ptr]
        return [st["peer"][q][parity] for q in range(p.world)], [st["peer"][q][2] for q in range(p.world)]

    # -- per step -----------------------------------------------------------------------------------
    def local_views(self, f: int, h: int, device, dtype=torch.float32):
        """Views of slot 0 (this rank's rows) of the table the NEXT gather_* call will publish: ``(z [rows, f] of
        ``dtype``, er [rows, h] fp32)``.  ``dtype=torch.bfloat16``: the bf16 storage mode travels over NVLink too
        (half the bytes of the pull; er stays fp32 behind the bf16 features of its row)."""
        st = self._setup(self._row_bytes(f, h, dtype), device)
        z, er = self._views(st, (self.step + 1) & 1, f, h, dtype)
        return z[:self.part.rows], (er[:self.part.rows] if er is not None else None)

    def _publish(self, st, device, er: torch.Tensor | None) -> Gate:
        p = self.part
        lib = _cabi.load()
        self.step += 1
        parity = self.step & 1
        tables, signals = self._peer_ptrs(st["row_bytes"], device, parity)
        sig_arr = (C.c_void_p * p.world)(*signals)
        stats = kernels.er_stats(er, 0) if er is not None and p.rows > 0 else None      # None: no power-of-two head count
        heads = int(er.shape[1]) if stats is not None else 0
        _cabi.check(lib.gta_exchange_publish(_cabi.ptr(stats), heads, p.rank, p.world, self.step, sig_arr, _stream()),
                    "gta_exchange_publish")
        ex = _cabi.Exchange()
        ex.world, ex.rank, ex.step, ex.copy_ctas = p.world, p.rank, self.step, self.copy_ctas
        ex.slot_rows, ex.row_bytes = p.stride, st["row_bytes"]
        ex.table, ex.signals = st["tables"][parity].ptr, st["signals"].ptr
        for k in range(p.world):
            q = p.owner_of(k)
            ex.peer_table[k] = tables[q]
            ex.slot_valid_rows[k] = p.bounds[q + 1] - p.bounds[q]
        return Gate(ex, keep=(st,))

    @kernels._timed("exchange_publish")
    def gather_pair(self, z: torch.Tensor, er: torch.Tensor):
        """``(z_table, er_table, gate)``: views of this step's gathered table -- only slot 0 is valid until the
        aggregation launch that is handed ``gate`` has pulled the rest."""
        f, h = int(z.shape[1]), int(er.shape[1])
        st = self._setup(self._row_bytes(f, h, z.dtype), z.device)
        views = self.local_views(f, h, z.device, z.dtype)
        if not (views[0].data_ptr() == z.data_ptr() and views[1].data_ptr() == er.data_ptr()):
            views[0].copy_(z)          # the producer did not write into the slot: copy
            views[1].copy_(er)
        gate = self._publish(st, z.device, views[1])
        self._last = (st, f, h, z.dtype)
        zt, ert = self._views(st, self.step & 1, f, h, z.dtype)
        return zt, ert, gate

    @kernels._timed("exchange_publish")
    def gather_one(self, t: torch.Tensor):
        f = int(t.shape[1])
        st = self._setup(self._row_bytes(f, 0, t.dtype), t.device)
        self.local_views(f, 0, t.device, t.dtype)[0].copy_(t)
        gate = self._publish(st, t.device, None)
        self._last = (st, f, 0, t.dtype)
        return self._views(st, self.step & 1, f, 0, t.dtype)[0], gate

    def last_table(self, width: int = 0) -> torch.Tensor:
        """The gathered ``[world*stride, f]`` feature table of the last step (complete once the aggregation launch
        that pulled it has finished)."""
        st, f, h, dtype = self._last
        return self._views(st, self.step & 1, f, h, dtype)[0]

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        """Whole table at once for the generic kernels (any legal plan runs): an NCCL all-gather in rank order,
        rolled into this rank's slot order.  Not the fast path."""
        p = self.part
        if t.shape[0] != p.rows:
            return t
        width = int(t.shape[1])
        buf = torch.zeros((p.world, p.stride, kernels.pad4(width)), dtype=torch.float32, device=t.device)
        buf[p.rank, :p.rows, :width].copy_(t)
        if p.world > 1:
            if self._emulated is not None:
                raise RuntimeError("the emulated fused exchange has no whole-table gather")
            dist.all_gather_into_tensor(buf.view(-1, buf.shape[-1]), buf[p.rank].clone(), group=self.group)
        return torch.roll(buf, -p.rank, 0).view(-1, buf.shape[-1])[:, :width]
