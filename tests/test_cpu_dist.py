"""Multi-rank host logic on CPU: world_size 2 over gloo (no GPU).  The exchange of source-side
tables (dist.SourceExchange) must place rank p's rows at p*stride of the gathered table -- the
layout gta_remap_sources rewrites source ids for."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gta_oracle as O
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import dist as gdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, bounds, stride, width, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = gdist.Partition(rank, world, bounds, stride, None, bounds[-1])
    ex = gdist.SourceExchange(part)
    rows = part.rows
    local = torch.arange(rows * width, dtype=torch.float32).reshape(rows, width) + 1000.0 * (rank + 1)
    full = ex(local)                       # executor hook: copy into the slot, all-gather in place
    again = ex(full)                       # a full table passes through untouched
    assert again.data_ptr() == full.data_ptr()
    np.save(os.path.join(out_dir, f"full_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("width", [4, 6])
def test_source_exchange_world2_gloo(tmp_path, width):
    bounds = [0, 5, 8]                      # uneven destination ranges
    stride = 8                              # max rows rounded up to a multiple of 4
    port = _free_port()
    mp.spawn(_worker, args=(2, port, bounds, stride, width, str(tmp_path)), nprocs=2, join=True)
    f0 = np.load(tmp_path / "full_0.npy")
    f1 = np.load(tmp_path / "full_1.npy")
    assert f0.shape == (16, width) and np.array_equal(f0, f1)
    for p in range(2):
        rows = bounds[p + 1] - bounds[p]
        want = np.arange(rows * width, dtype=np.float32).reshape(rows, width) + 1000.0 * (p + 1)
        assert np.array_equal(f0[p * stride: p * stride + rows], want)
        assert np.all(f0[p * stride + rows:(p + 1) * stride] == 0)      # padding rows stay zero


def test_remap_formula_matches_partition_layout():
    """Pure-numpy statement of gta_remap_sources against the oracle's partition bounds: the
    remapped CSR gathers exactly the rows the global CSR gathers."""
    g = synthetic.powerlaw_graph(1000, 20000, seed=5, i0=6.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, 1000)
    world = 4
    b = O.partition_bounds(indptr, world)
    stride = int(-(-np.diff(b).max() // 4) * 4)
    owner = np.searchsorted(b, indices, side="right") - 1
    remapped = owner * stride + (indices - b[owner])
    x = np.random.default_rng(0).standard_normal((1000, 8))
    table = np.zeros((world * stride, 8))
    for p in range(world):
        table[p * stride: p * stride + (b[p + 1] - b[p])] = x[b[p]:b[p + 1]]
    assert np.array_equal(table[remapped], x[indices])
    # monotonic: ascending-source order inside every row is preserved
    rows = O.row_ids(indptr)
    same_row = rows[1:] == rows[:-1]
    assert np.all(np.diff(remapped)[same_row] > 0)


def _worker_chunked(rank, world, port, bounds, stride, chunks, width, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = gdist.Partition(rank, world, bounds, stride, None, bounds[-1], chunks)
    ex = gdist.SourceExchange(part)
    local = torch.arange(part.rows * width, dtype=torch.float32).reshape(part.rows, width) + 1000.0 * (rank + 1)
    full = ex(local)
    np.save(os.path.join(out_dir, f"chunked_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


def test_chunked_exchange_layout_world2_gloo(tmp_path):
    """chunks = 2: the gathered table is [chunks, world, stride/chunks, F]; a source with owner p and
    local offset o sits at q*(world*cs) + p*cs + (o - q*cs) -- the formula gta_remap_sources applies."""
    bounds, stride, chunks, width = [0, 5, 8], 8, 2, 4
    port = _free_port()
    mp.spawn(_worker_chunked, args=(2, port, bounds, stride, chunks, width, str(tmp_path)), nprocs=2, join=True)
    f0 = np.load(tmp_path / "chunked_0.npy")
    assert np.array_equal(f0, np.load(tmp_path / "chunked_1.npy")) and f0.shape == (16, width)
    cs = stride // chunks
    for p in range(2):
        rows = bounds[p + 1] - bounds[p]
        want = np.arange(rows * width, dtype=np.float32).reshape(rows, width) + 1000.0 * (p + 1)
        for o in range(rows):
            q = o // cs
            assert np.array_equal(f0[q * (2 * cs) + p * cs + (o - q * cs)], want[o])
