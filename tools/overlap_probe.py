#!/usr/bin/env python
"""Timeline of copy/compute overlap (GPU box): H2D of a Reddit-size feature table on one stream
while an aggregation kernel runs on another."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph, kernels, pipeline
from tools.agg_probe import device_powerlaw

n, e, f, h, fin = 232965, 114615892, 128, 4, 602
dst, src = device_powerlaw(n, e)
g = graph.csr_from_coo(dst, src, n); del dst, src
z = kernels.alloc_table(n, f, "cuda"); z.normal_()
el = torch.randn(n, h, device="cuda"); er = torch.randn(n, h, device="cuda")
x_pin = pipeline.pinned_table(n, fin); x_dev = kernels.alloc_table(n, fin, "cuda")
y_pin = torch.empty((n, f)).pin_memory()
flat = lambda t: torch.as_strided(t, (t.shape[0] * t.stride(0),), (1,))
print("pinned:", x_pin.is_pinned(), flat(x_pin).is_pinned(), y_pin.is_pinned())
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
for mode in ("copy only", "kernel only", "both", "both+d2h"):
    torch.cuda.synchronize()
    a0, a1, b0, b1, c0, c1 = ev(), ev(), ev(), ev(), ev(), ev()
    t0 = ev(); t0.record(); 
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream()); s3.wait_stream(torch.cuda.current_stream())
    if mode != "kernel only":
        with torch.cuda.stream(s1):
            a0.record(s1); flat(x_dev).copy_(flat(x_pin), non_blocking=True); a1.record(s1)
    if mode != "copy only":
        with torch.cuda.stream(s2):
            b0.record(s2)
            for _ in range(2): y = kernels.gat_aggregate(g, el, er, z)
            b1.record(s2)
    if mode == "both+d2h":
        with torch.cuda.stream(s3):
            c0.record(s3); y_pin.copy_(y, non_blocking=True); c1.record(s3)
    torch.cuda.synchronize()
    msg = mode + ":"
    if mode != "kernel only": msg += f" H2D [{t0.elapsed_time(a0):.2f}, {t0.elapsed_time(a1):.2f}]"
    if mode != "copy only": msg += f" kernels x2 [{t0.elapsed_time(b0):.2f}, {t0.elapsed_time(b1):.2f}]"
    if mode == "both+d2h": msg += f" D2H [{t0.elapsed_time(c0):.2f}, {t0.elapsed_time(c1):.2f}]"
    print(msg)
