#!/usr/bin/env python
"""Device graph preprocessing timings (GPU box): COO->CSR, work list, partition, reorder and the
reference's tile tables (calculate_sparsity / cal_min_sparsity over a whole size list)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph, synthetic


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r


for shape, sizes in (("cora", graph.gen_size(16, 2720)),
                     ("reddit", [16, 64, 256, 1024, 1600, 2048, 2400, 3072, 3200, 4000, 4096, 4800, 5120, 5600, 6144, 6400,
                                 7168, 7200, 8192])):       # the authors' Reddit list, FinalVersion For Paper/compiler.py:95
    coo = synthetic.shape_graph(shape)
    n = coo.num_nodes
    dst, src = torch.from_numpy(coo.dst).cuda(), torch.from_numpy(coo.src).cuda()
    ms_csr, g = timed(lambda: graph.csr_from_coo(dst, src, n))
    ms_sched, _ = timed(lambda: graph.build_schedule(g.indptr, g.indices, 0, n, g.num_edges, n, 1024, 0))
    ms_part, _ = timed(lambda: graph.partition_bounds(g, 8))
    ms_reord, _ = timed(lambda: graph.degree_reorder(g))
    t0 = time.perf_counter()
    maxlist = [graph.cal_min_sparsity(g, s, workspace_bytes=1 << 30) for s in sizes]
    torch.cuda.synchronize()
    ms_tiles = (time.perf_counter() - t0) * 1e3
    print(f"{shape}: N={n} E={g.num_edges}  csr {ms_csr:.2f} ms  work list {ms_sched:.2f} ms  partition {ms_part:.3f} ms  "
          f"reorder {ms_reord:.2f} ms  maxlist over {len(sizes)} tile sizes {ms_tiles:.1f} ms -> {maxlist[:6]}...{maxlist[-3:]}")
