"""Destination-range partitioning over the GPUs of one box.

One process per GPU (torchrun).  Rank p owns CSR rows ``[bounds[p], bounds[p+1])`` -- bounds
balanced by edge count (gta_partition) -- and the matching rows of every node tensor.  The only
exchange per layer is an all-gather of the source-side tables (``Z`` and ``er`` for GAT, ``Z``
for GCN) over NVLink (NCCL); destination rows are disjoint, so there is no reduction.

The gathered table is ``[parts, stride, F]`` with every rank's rows padded to ``stride`` =
max rows per rank; local source ids are remapped once at setup (gta_remap_sources).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _cabi
from .graph import DeviceGraph, _stream, partition_bounds, slice_rows


@dataclass
class Partition:
    rank: int
    world: int
    bounds: list          # python ints, len world+1
    stride: int           # padded rows per rank in gathered tables
    local: DeviceGraph    # zero-based local CSR, sources remapped into the gathered table
    num_nodes: int

    @property
    def row_begin(self) -> int:
        return self.bounds[self.rank]

    @property
    def row_end(self) -> int:
        return self.bounds[self.rank + 1]

    @property
    def rows(self) -> int:
        return self.row_end - self.row_begin


def make_partition(full: DeviceGraph, rank: int, world: int) -> Partition:
    """Cut a (replicated) full graph into this rank's destination range."""
    lib = _cabi.load()
    b = partition_bounds(full, world)
    bounds = [int(v) for v in b.cpu().tolist()]
    stride = max(bounds[p + 1] - bounds[p] for p in range(world))
    stride = (stride + 3) // 4 * 4
    local = slice_rows(full, bounds[rank], bounds[rank + 1])
    remapped = torch.empty_like(local.indices)
    _cabi.check(lib.gta_remap_sources(_cabi.ptr(local.indices), local.num_edges, _cabi.ptr(b), world, stride,
                                      _cabi.ptr(remapped), _stream()), "gta_remap_sources")
    local.indices = remapped
    local.num_sources = world * stride
    return Partition(rank, world, bounds, stride, local, full.num_nodes)


class SourceExchange:
    """All-gather of a local ``[rows, F]`` table into the padded ``[world*stride, F]`` table.

    Buffers are cached per (width) so the steady state allocates nothing; the local rows are
    produced directly into this rank's slot (``local_slot``) so the collective runs in place."""

    def __init__(self, part: Partition, group=None):
        self.part = part
        self.group = group
        self._buf = {}

    def buffer(self, width: int, device) -> torch.Tensor:
        key = (width, device)
        if key not in self._buf:
            ld = (width + 3) // 4 * 4
            self._buf[key] = torch.zeros((self.part.world * self.part.stride, ld), dtype=torch.float32,
                                         device=device)
        return self._buf[key]

    def local_slot(self, width: int, device) -> torch.Tensor:
        """View of this rank's rows inside the gathered buffer ([rows, width])."""
        p = self.part
        buf = self.buffer(width, device)
        return buf[p.rank * p.stride: p.rank * p.stride + p.rows, :width]

    def gather(self, width: int, device) -> torch.Tensor:
        """In-place all-gather; returns the full ``[world*stride, width]`` view."""
        p = self.part
        buf = self.buffer(width, device)
        if p.world > 1:
            mine = buf[p.rank * p.stride:(p.rank + 1) * p.stride]
            dist.all_gather_into_tensor(buf, mine, group=self.group)
        return buf[:, :width]

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        """Generic path (executor ``source_table`` hook): copy a local table in, gather."""
        if t.shape[0] != self.part.rows:
            return t            # already a full table
        width = int(t.shape[1])
        slot = self.local_slot(width, t.device)
        if slot.data_ptr() != t.data_ptr():
            slot.copy_(t)
        return self.gather(width, t.device)
