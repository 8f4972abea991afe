"""Multi-rank host logic on CPU: world_size 2 over gloo (no GPU).  The all-gather exchange
(dist.SourceExchange) must place rank p's rows at p*stride of the gathered table, the fused exchange's
whole-table fallback at ((p - rank) mod world)*stride -- the two layouts gta_remap_sources rewrites source
ids for -- and a partitioned execute() must reproduce the unpartitioned rows."""
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
    part = gdist.Partition(rank, world, bounds, stride, None, bounds[-1], rotate=False)
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
    stride = 8                              # max rows rounded up to a multiple of 8 (whole 128-byte lines per slot)
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


def test_rotated_remap_formula():
    """rotate = r (fused exchange): slot k of rank r holds rank (r + k) mod world, the rank's own rows first.
    Numpy statement of gta_remap_sources(rotate=r): the remapped ids gather the same rows from the rolled table,
    and re-sorting a row by them is a cyclic shift of its ascending source list."""
    g = synthetic.powerlaw_graph(1000, 20000, seed=5, i0=6.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, 1000)
    world = 4
    b = O.partition_bounds(indptr, world)
    stride = int(-(-np.diff(b).max() // 8) * 8)
    owner = np.searchsorted(b, indices, side="right") - 1
    x = np.random.default_rng(0).standard_normal((1000, 8))
    for r in range(world):
        part = gdist.Partition(r, world, [int(v) for v in b], stride, None, 1000, rotate=True)
        assert [part.slot_of(part.owner_of(k)) for k in range(world)] == list(range(world)) and part.slot_of(r) == 0
        remapped = ((owner - r) % world) * stride + (indices - b[owner])
        table = np.zeros((world * stride, 8))
        for p in range(world):
            k = part.slot_of(p)
            table[k * stride: k * stride + (b[p + 1] - b[p])] = x[b[p]:b[p + 1]]
        assert np.array_equal(table[remapped], x[indices])
        lo, hi = indptr[b[r]], indptr[b[r] + 1]          # first row of the rank: sources >= b[r] come first
        order = np.argsort(remapped[lo:hi], kind="stable")
        srcs = indices[lo:hi][order]
        cut = int(np.sum(indices[lo:hi] >= b[r]))
        assert np.array_equal(srcs, np.concatenate([indices[lo:hi][-cut:] if cut else [], indices[lo:hi][:hi - lo - cut]]))


def _worker_rolled(rank, world, port, bounds, stride, width, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = gdist.Partition(rank, world, bounds, stride, None, bounds[-1], rotate=True)
    ex = gdist.FusedExchange(part)
    local = torch.arange(part.rows * width, dtype=torch.float32).reshape(part.rows, width) + 1000.0 * (rank + 1)
    full = ex(local)                       # whole-table fallback: all-gather in rank order, rolled to slot order
    np.save(os.path.join(out_dir, f"rolled_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


def test_fused_exchange_whole_table_fallback_world2_gloo(tmp_path):
    bounds, stride, width = [0, 5, 8], 8, 4
    port = _free_port()
    mp.spawn(_worker_rolled, args=(2, port, bounds, stride, width, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        f = np.load(tmp_path / f"rolled_{rank}.npy")
        assert f.shape == (16, width)
        for k in range(2):
            p = (rank + k) % 2
            rows = bounds[p + 1] - bounds[p]
            want = np.arange(rows * width, dtype=np.float32).reshape(rows, width) + 1000.0 * (p + 1)
            assert np.array_equal(f[k * stride: k * stride + rows], want)
            assert np.all(f[k * stride + rows:(k + 1) * stride] == 0)


# ---- end to end: a partitioned execute() over gloo, kernels replaced by the CPU test double -------------
def _cpu_partition(indptr, indices, rank, world, n, rotate):
    """Host restatement of dist.make_partition (gta_partition + slice + gta_remap_sources + per-row re-sort)."""
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph
    b = O.partition_bounds(indptr, world)
    bounds = [int(v) for v in b]
    stride = int(-(-max(np.diff(b)) // 8) * 8)
    full = graph.DeviceGraph(n, int(indptr[-1]), torch.from_numpy(indptr), torch.from_numpy(indices.astype(np.int32)),
                             num_sources=n)
    local = graph.slice_rows(full, bounds[rank], bounds[rank + 1])
    src = local.indices.numpy().astype(np.int64)
    owner = np.searchsorted(b, src, side="right") - 1
    slot = (owner - rank) % world if rotate else owner
    remapped = slot * stride + (src - b[owner])
    rows = O.row_ids(local.indptr.numpy())
    perm = np.lexsort((remapped, rows))          # what gta_csr_build does with (local row, remapped source)
    local.indices = torch.from_numpy(remapped[perm].astype(np.int32))
    local.perm = torch.from_numpy(perm.astype(np.int64))
    local.num_sources = world * stride
    return gdist.Partition(rank, world, bounds, stride, local, n, rotate), full


class _WholeTable:
    """The fused exchange's whole-table fallback only (its in-kernel path needs CUDA IPC): rotated layout."""

    def __init__(self, part):
        self.part, self._ex = part, gdist.FusedExchange(part)

    def __call__(self, t):
        return self._ex(t)


def _worker_execute(rank, world, port, prog_file, op_file, network, reorder, rotate, out_dir):
    import sys
    import yaml
    tests_dir = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, tests_dir)
    import host_kernels
    import test_gpu_executor as shared
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    executor.kernels = host_kernels
    graph.DeviceGraph.schedule = lambda self, *a, **k: None
    n, e = 600, 7000
    g = synthetic.powerlaw_graph(n, e, seed=9, i0=12.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    op_info = yaml.safe_load(open(op_file))
    records = yaml.safe_load(open(prog_file))
    node_inputs, weights, edge_inputs = shared._inputs(op_info, n, g.num_edges)
    part, full = _cpu_partition(indptr, indices, rank, world, n, rotate)
    r0, r1 = part.row_begin, part.row_end
    e0, e1 = int(indptr[r0]), int(indptr[r1])
    t = torch.from_numpy
    # edge inputs come in the global CSR order of the rank's rows: permute_edges puts them in the local order
    out = executor.execute(records, op_info, part.local, {k: t(v[r0:r1]) for k, v in node_inputs.items()},
                           {k: t(v) for k, v in weights.items()},
                           {k: part.permute_edges(t(v[e0:e1])) for k, v in edge_inputs.items()},
                           network=network, is_reorder=reorder, check_shapes=False,
                           source_table=_WholeTable(part) if rotate else gdist.SourceExchange(part))
    (p, y), = out.items()
    np.save(os.path.join(out_dir, f"y_{rank}.npy"), y.numpy())
    if rank == 0:       # the same program, unpartitioned, same test double
        whole = executor.execute(records, op_info, full, {k: t(v) for k, v in node_inputs.items()},
                                 {k: t(v) for k, v in weights.items()}, {k: t(v) for k, v in edge_inputs.items()},
                                 network=network, is_reorder=reorder, check_shapes=False)
        np.save(os.path.join(out_dir, "y_whole.npy"), whole[p].numpy())
        sem = O.NETWORK_SEMANTICS.get((network, reorder), {})
        ref = O.run_opgraph(op_info, indptr, indices, node_inputs, weights, edge_inputs, semantics=sem, stabilize=True)
        np.save(os.path.join(out_dir, "y_ref.npy"), ref[p])
        np.save(os.path.join(out_dir, "bounds.npy"), np.asarray(part.bounds))
    dist.destroy_process_group()


@pytest.mark.parametrize("rotate", [False, True], ids=["allgather-layout", "rotated-layout"])
@pytest.mark.parametrize("name,network,reorder", [
    ("GAT-cora-layer1-original__0-1-2_4-5-6-7-8_3-9-10-11-12-13", "GAT", False),
    ("GCN-cora-layer1-trans__0_1-2-3", "GCN", True),
    # the linear DGN edge phase (degree x own rows uses the LOCAL degrees) and PNA's one-pass edge sum (gathered rows
    # through the exchange, row term and edge features local) in a partitioned run
    ("DGN-cora-layer2-original__0-1-2-3-4-5-6-7-8-9-10", "DGN", False),
    ("PNA-cora-layer2-original__0-1-2-3-4-5-6-7-8-9-10", "PNA", False),
])
def test_partitioned_execute_world2_gloo(tmp_path, name, network, reorder, rotate):
    """Destination-range partition, one replication of the source-side table per layer: the rows two ranks
    compute (host logic of dist.py + executor.py, kernels = CPU test double) are the rows one process computes --
    bit for bit in the all-gather layout (same reduction order), within tolerance in the rotated one."""
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    mode = "trans" if reorder else "original"
    port = _free_port()
    mp.spawn(_worker_execute, args=(2, port, os.path.join(golden, "isa", name + ".yaml"),
                                    os.path.join(golden, "opgraph", f"{'-'.join(name.split('-')[:3])}-{mode}.yaml"), network, reorder,
                                    rotate, str(tmp_path)), nprocs=2, join=True)
    bounds = np.load(tmp_path / "bounds.npy")
    whole, ref = np.load(tmp_path / "y_whole.npy"), np.load(tmp_path / "y_ref.npy")
    parts = [np.load(tmp_path / f"y_{r}.npy") for r in range(2)]
    assert [p.shape[0] for p in parts] == list(np.diff(bounds)) and 0 < bounds[1] < bounds[2]
    got = np.concatenate(parts)
    if not rotate:
        assert np.array_equal(got, whole)                   # same reduction order -> same bits
    np.testing.assert_allclose(got, whole, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-5 * np.abs(ref).max())
