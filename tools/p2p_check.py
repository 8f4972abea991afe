#!/usr/bin/env python
"""torchrun check of the peer-to-peer exchange: the GAT layer through dist.PeerExchange (copy engines)
must equal the NCCL all-gather path -- bit for bit at chunks = 1, and for chunks > 1 equal to the NCCL
path with the same chunked layout.  Run:  torchrun --nproc-per-node 2 tools/p2p_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph, kernels, synthetic
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import dist as gdist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n, e, fin, f, h = 20000, 2000000, 96, 128, 4
g = synthetic.powerlaw_graph(n, e, seed=3, i0=20.0)
full = graph.csr_from_coo(g.dst, g.src, n)
x, w, al, ar = synthetic.gat_tensors(n, fin, f, h, seed=1)
ok = True
for chunks in (1, 4):
    part = gdist.make_partition(full, rank, world, chunks=chunks)
    xd = kernels.to_table(torch.from_numpy(x[part.row_begin:part.row_end]).to(dev))
    wd, ald, ard = (torch.from_numpy(a).to(dev) for a in (w, al, ar))
    outs = {}
    for name, ex in (("nccl", gdist.SourceExchange(part)), ("p2p", gdist.PeerExchange(part))):
        res = []
        for step in range(4):                       # several steps: exercises the double buffering
            z, el, er = kernels.gemm(xd * (1.0 + step), wd, ald, ard)
            zf, erf, events = ex.gather_pair(z, er, overlap=True)
            sched = part.local.schedule(col_block=part.col_block) if events is not None else None
            res.append(kernels.gat_aggregate(part.local, el, erf, zf, sched=sched, block_events=events).clone())
        torch.cuda.synchronize()
        outs[name] = res
    same = all(torch.equal(a, b) for a, b in zip(outs["nccl"], outs["p2p"]))
    finite = all(bool(torch.isfinite(a).all()) for a in outs["p2p"])
    t = torch.tensor([int(same and finite)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"chunks={chunks}: p2p == nccl on every rank and step: {bool(t.item())}")
    ok = ok and bool(t.item())
dist.barrier()
if rank == 0:
    print("P2P CHECK", "PASSED" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
