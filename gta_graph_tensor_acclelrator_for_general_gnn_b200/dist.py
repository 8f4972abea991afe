"""Destination-range partitioning over the GPUs of one box.

One process per GPU (torchrun).  Rank p owns CSR rows ``[bounds[p], bounds[p+1])`` -- bounds
balanced by edge count (gta_partition) -- and the matching rows of every node tensor.  The only
exchange per layer is an all-gather of the source-side tables (``[Z | er]`` for GAT, ``Z`` for GCN)
over NVLink (NCCL); destination rows are disjoint, so there is no reduction.

Overlap.  Every rank's rows are padded to ``stride`` and cut into ``chunks`` equal pieces; the
gathered table is laid out ``[chunks, world, stride/chunks, F]`` so that chunk q of EVERY rank is
one contiguous NCCL all-gather and one COLUMN BLOCK of the aggregation work list.  The all-gathers
run on a communication stream, chunk after chunk; the aggregation kernel of column block q waits
only for chunk q's event, so the transfer of chunk q+1 hides under the gathers of chunk q
(measured on 8 B200: one 119 MB all-gather is 0.2 ms of a 1.0 ms layer).  Source ids are remapped
once at setup (gta_remap_sources) and every row is re-sorted by the new ids: a fixed reduction
order, deterministic, but not the single-GPU order when chunks > 1.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _cabi, kernels
from .graph import DeviceGraph, _stream, csr_from_coo, partition_bounds, slice_rows


@dataclass
class Partition:
    rank: int
    world: int
    bounds: list          # python ints, len world+1
    stride: int           # padded rows per rank in gathered tables (multiple of 4*chunks)
    local: DeviceGraph    # zero-based local CSR, sources remapped into the gathered table
    num_nodes: int
    chunks: int = 1

    @property
    def row_begin(self) -> int:
        return self.bounds[self.rank]

    @property
    def row_end(self) -> int:
        return self.bounds[self.rank + 1]

    @property
    def rows(self) -> int:
        return self.row_end - self.row_begin

    @property
    def chunk_rows(self) -> int:
        return self.stride // self.chunks

    @property
    def col_block(self) -> int:
        """Source ids per column block of the work list = one gathered chunk (0: no chunking)."""
        return self.world * self.chunk_rows if self.chunks > 1 else 0


def make_partition(full: DeviceGraph, rank: int, world: int, chunks: int = 1) -> Partition:
    """Cut a (replicated) full graph into this rank's destination range."""
    lib = _cabi.load()
    b = partition_bounds(full, world)
    bounds = [int(v) for v in b.cpu().tolist()]
    stride = max(bounds[p + 1] - bounds[p] for p in range(world))
    unit = 4 * chunks
    stride = (stride + unit - 1) // unit * unit
    local = slice_rows(full, bounds[rank], bounds[rank + 1])
    remapped = torch.empty_like(local.indices)
    _cabi.check(lib.gta_remap_sources(_cabi.ptr(local.indices), local.num_edges, _cabi.ptr(b), world, stride, chunks,
                                      _cabi.ptr(remapped), _stream()), "gta_remap_sources")
    if chunks > 1:
        # the chunked layout is not monotonic in the source id: re-sort every row by the new ids
        rows = local.num_rows
        deg = (local.indptr[1:] - local.indptr[:-1])
        row_of_edge = torch.repeat_interleave(torch.arange(rows, dtype=torch.int32, device=deg.device), deg)
        local = csr_from_coo(row_of_edge, remapped, rows)
        local.num_nodes = full.num_nodes
    else:
        local.indices = remapped
    local.num_sources = world * stride
    return Partition(rank, world, bounds, stride, local, full.num_nodes, chunks)


class SourceExchange:
    """All-gather of a local ``[rows, F]`` table into the gathered ``[chunks, world, stride/chunks, F]``
    table.  Buffers are cached per width, so the steady state allocates nothing."""

    def __init__(self, part: Partition, group=None):
        self.part = part
        self.group = group
        self._buf = {}
        self._comm = None
        self._events = None

    # -- layout ---------------------------------------------------------------------------------
    def buffer(self, width: int, device) -> torch.Tensor:
        key = (width, torch.device(device))
        if key not in self._buf:
            p = self.part
            ld = (width + 3) // 4 * 4
            self._buf[key] = torch.zeros((p.chunks, p.world, p.chunk_rows, ld), dtype=torch.float32, device=device)
        return self._buf[key]

    def table(self, width: int, device) -> torch.Tensor:
        """The gathered table as a flat ``[chunks*world*chunk_rows, width]`` view (what kernels index)."""
        buf = self.buffer(width, device)
        return buf.view(-1, buf.shape[-1])[:, :width]

    def store_local(self, t: torch.Tensor, width: int, col: int = 0) -> None:
        """Copy this rank's ``[rows, w]`` table into columns ``[col, col+w)`` of its chunk slots."""
        p = self.part
        buf = self.buffer(width, t.device)
        cs = p.chunk_rows
        w = int(t.shape[1])
        for q in range(p.chunks):
            lo, hi = q * cs, min((q + 1) * cs, p.rows)
            if hi > lo:
                buf[q, p.rank, :hi - lo, col:col + w].copy_(t[lo:hi])

    def local_views(self, f: int, h: int, device):
        """``(z [rows, f], er [rows, h])`` views of this rank's slot of the gathered ``[F | H]`` table,
        for producers (the GEMM) that write there directly.  Only when the slot is one contiguous
        piece (chunks == 1); otherwise None."""
        p = self.part
        if p.chunks != 1:
            return None
        width = f + (h + 3) // 4 * 4
        slot = self.buffer(width, device)[0, p.rank, :p.rows]
        return slot[:, :f], slot[:, f:f + h]

    # -- collective -------------------------------------------------------------------------------
    def gather(self, width: int, device, overlap: bool = False):
        """All-gather every chunk in place.  ``overlap=False``: the current stream waits for all of
        them, returns the table.  ``overlap=True``: the transfers run on the communication stream
        and ``(table, [event per chunk])`` is returned; consumers wait per chunk."""
        p = self.part
        buf = self.buffer(width, device)
        table = buf.view(-1, buf.shape[-1])[:, :width]
        if p.world == 1:
            return (table, [None] * p.chunks) if overlap else table
        if not overlap:
            for q in range(p.chunks):
                dist.all_gather_into_tensor(buf[q].view(-1, buf.shape[-1]), buf[q, p.rank], group=self.group)
            return table
        cur = torch.cuda.current_stream()
        if self._comm is None:
            self._comm = torch.cuda.Stream()
            self._events = [torch.cuda.Event() for _ in range(p.chunks)]
        self._comm.wait_stream(cur)                 # the local slots were written on the compute stream
        with torch.cuda.stream(self._comm):
            for q in range(p.chunks):
                dist.all_gather_into_tensor(buf[q].view(-1, buf.shape[-1]), buf[q, p.rank], group=self.group)
                self._events[q].record(self._comm)
        return table, list(self._events)

    @kernels._timed("nccl_all_gather")
    def gather_pair(self, z: torch.Tensor, er: torch.Tensor, overlap: bool = False):
        """One gathered table for the two source-side tensors of a GAT layer: every source row is
        ``[z (F) | er (H)]``.  Returns ``(z_view, er_view, events | None)`` (strided views)."""
        f, h = int(z.shape[1]), int(er.shape[1])
        width = f + (h + 3) // 4 * 4
        views = self.local_views(f, h, z.device)
        in_place = views is not None and views[0].data_ptr() == z.data_ptr() and views[1].data_ptr() == er.data_ptr()
        if not in_place:             # the producer did not write into the slot: copy
            self.store_local(z, width, 0)
            self.store_local(er, width, f)
        if overlap and self.part.chunks > 1:
            full, events = self.gather(width, z.device, overlap=True)
        else:
            full, events = self.gather(width, z.device), None
        return full[:, :f], full[:, f:f + h], events

    @kernels._timed("nccl_all_gather")
    def gather_one(self, t: torch.Tensor, overlap: bool = False):
        """``(table, events | None)`` for a single source-side tensor (GCN: Z)."""
        width = int(t.shape[1])
        self.store_local(t, width, 0)
        if overlap and self.part.chunks > 1:
            return self.gather(width, t.device, overlap=True)
        return self.gather(width, t.device), None

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        """Executor ``source_table`` hook for consumers that need the whole table at once."""
        if t.shape[0] != self.part.rows:
            return t            # already a gathered table
        return self.gather_one(t)[0]


# ---- peer-to-peer exchange on the copy engines --------------------------------------------------

class _IpcBuffer:
    """fp32 buffer cudaMalloc'ed by the library (so its CUDA IPC handle names a base pointer), exposed
    to torch through ``__cuda_array_interface__``."""

    def __init__(self, shape):
        import ctypes as C
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
        import ctypes as C
        buf = C.create_string_buffer(64)
        _cabi.check(self.lib.gta_ipc_export(self.ptr, buf), "gta_ipc_export")
        return buf.raw

    def close(self):
        if self.ptr:
            self.lib.gta_ipc_free(self.ptr)
            self.ptr = 0


class PeerExchange(SourceExchange):
    """Same contract as :class:`SourceExchange`, but the transfer is a set of device-to-device copies
    from the peers' IPC-mapped slot buffers on a copy stream (copy engines over NVLink, no SMs), so it
    really overlaps the aggregation kernel.  Per step: a one-element NCCL all-reduce on the compute
    stream (every rank's slot is written), then ``chunks`` groups of ``world`` copies, one event per
    chunk.  The slot buffers are double-buffered: a peer may still be pulling step i while this rank
    already writes step i+1.  Raises ``RuntimeError`` on every rank if any rank cannot map its peers
    (the caller falls back to the NCCL all-gather)."""

    def __init__(self, part: Partition, group=None):
        super().__init__(part, group)
        self._state = {}

    def _setup(self, width: int, device):
        import ctypes as C
        key = (width, torch.device(device))
        if key in self._state:
            return self._state[key]
        p = self.part
        lib = _cabi.load()
        ld = (width + 3) // 4 * 4
        cs = p.chunk_rows
        st = {"step": 0, "ld": ld}
        ok = 1
        err = ""
        try:
            st["mine"] = [_IpcBuffer((p.chunks, cs, ld)) for _ in range(2)]
            st["mine_t"] = [m.tensor() for m in st["mine"]]
            handles = [m.handle() for m in st["mine"]]
        except Exception as exc:          # keep going to the collective below so every rank agrees
            ok, err, handles = 0, str(exc), [b"", b""]
        gathered = [None] * p.world
        dist.all_gather_object(gathered, handles, group=self.group)
        peer = [[0, 0] for _ in range(p.world)]
        if ok:
            try:
                for r in range(p.world):
                    for b in range(2):
                        if r == p.rank:
                            peer[r][b] = st["mine"][b].ptr
                        else:
                            mapped = C.c_void_p()
                            _cabi.check(lib.gta_ipc_open(gathered[r][b], C.byref(mapped)), "gta_ipc_open")
                            peer[r][b] = int(mapped.value)
            except Exception as exc:
                ok, err = 0, str(exc)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            raise RuntimeError("peer-to-peer exchange unavailable on at least one rank" + (": " + err if err else ""))
        table = self.buffer(width, device)
        chunk_bytes = cs * ld * 4
        plans = []
        for b in range(2):
            per_chunk = []
            for q in range(p.chunks):
                dsts, srcs, sizes = [], [], []
                for k in range(p.world):
                    r = (p.rank + k) % p.world            # own slot first, then round the ring
                    rows_r = p.bounds[r + 1] - p.bounds[r]
                    valid = max(0, min(cs, rows_r - q * cs))
                    if valid == 0:
                        continue
                    dsts.append(table.data_ptr() + (q * p.world + r) * chunk_bytes)
                    srcs.append(peer[r][b] + q * chunk_bytes)
                    sizes.append(valid * ld * 4)
                n = len(dsts)
                per_chunk.append(((C.c_void_p * n)(*dsts), (C.c_void_p * n)(*srcs), (C.c_int64 * n)(*sizes), n))
            plans.append(per_chunk)
        st.update(peer=peer, plans=plans, sync=torch.zeros(1, dtype=torch.float32, device=device),
                  stream=torch.cuda.Stream(), events=[torch.cuda.Event() for _ in range(p.chunks)])
        self._state[key] = st
        return st

    def local_views(self, f: int, h: int, device):
        p = self.part
        if p.chunks != 1:
            return None
        st = self._setup(f + (h + 3) // 4 * 4, device)
        slot = st["mine_t"][st["step"] % 2][0, :p.rows]
        return slot[:, :f], slot[:, f:f + h]

    def _store(self, st, t: torch.Tensor, col: int):
        p = self.part
        mine = st["mine_t"][st["step"] % 2]
        cs, w = p.chunk_rows, int(t.shape[1])
        for q in range(p.chunks):
            lo, hi = q * cs, min((q + 1) * cs, p.rows)
            if hi > lo:
                mine[q, :hi - lo, col:col + w].copy_(t[lo:hi])

    def _pull(self, st, width: int, device, overlap: bool):
        lib = _cabi.load()
        p = self.part
        cur = torch.cuda.current_stream()
        dist.all_reduce(st["sync"], group=self.group)       # on the compute stream: every slot is written
        copy = st["stream"]
        copy.wait_stream(cur)
        b = st["step"] % 2
        for q in range(p.chunks):
            dsts, srcs, sizes, n = st["plans"][b][q]
            _cabi.check(lib.gta_copy_many(dsts, srcs, sizes, n, copy.cuda_stream), "gta_copy_many")
            st["events"][q].record(copy)
        st["step"] += 1
        buf = self.buffer(width, device)
        table = buf.view(-1, buf.shape[-1])[:, :width]
        if overlap and p.chunks > 1:
            return table, list(st["events"])
        cur.wait_event(st["events"][-1])
        return table, None

    @kernels._timed("p2p_gather")
    def gather_pair(self, z: torch.Tensor, er: torch.Tensor, overlap: bool = False):
        f, h = int(z.shape[1]), int(er.shape[1])
        width = f + (h + 3) // 4 * 4
        st = self._setup(width, z.device)
        views = self.local_views(f, h, z.device)
        in_place = views is not None and views[0].data_ptr() == z.data_ptr() and views[1].data_ptr() == er.data_ptr()
        if not in_place:
            self._store(st, z, 0)
            self._store(st, er, f)
        full, events = self._pull(st, width, z.device, overlap)
        return full[:, :f], full[:, f:f + h], events

    @kernels._timed("p2p_gather")
    def gather_one(self, t: torch.Tensor, overlap: bool = False):
        width = int(t.shape[1])
        st = self._setup(width, t.device)
        self._store(st, t, 0)
        return self._pull(st, width, t.device, overlap)
