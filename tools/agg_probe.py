#!/usr/bin/env python
"""Micro-benchmark of the aggregation kernels on synthetic graphs (GPU box):
    python tools/agg_probe.py [--cases name:N:E:F:H ...] [--chunk 1024] [--iters 10]
Prints ms, GTEPS and algorithmic GB/s per case; used to separate L2-residency effects
(halve N at fixed E) from instruction/latency effects."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph, kernels  # noqa: E402


def device_powerlaw(n, e, seed=0):
    """Fast on-device generator (probe only): power-law endpoints, duplicates kept."""
    g = torch.Generator(device="cuda").manual_seed(seed)

    def ends(m):
        u = torch.rand(m, device="cuda", generator=g, dtype=torch.float64)
        a, b = (n + 100.0) ** 0.5, 100.0 ** 0.5
        return ((u * (a - b) + b) ** 2 - 100.0).floor().clamp_(0, n - 1).to(torch.int32)
    perm = torch.randperm(n, device="cuda", generator=g).to(torch.int32)
    a, b = perm[ends(e // 2).long()], perm[ends(e // 2).long()]
    return torch.cat([a, b]), torch.cat([b, a])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="*", default=["reddit:232965:114615892:128:4", "half:116482:114615892:128:4",
                                                   "quarter:58241:114615892:128:4", "f64:232965:114615892:64:4"])
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--kinds", nargs="*", default=["gat", "spmm"])
    ap.add_argument("--col-blocks", nargs="*", type=int, default=[1], help="column blocks to try (1 = none)")
    ap.add_argument("--split-launch", action="store_true", help="gate per column block (block 0, then the rest)")
    args = ap.parse_args()
    for case in args.cases:
        name, n, e, f, h = case.split(":")
        n, e, f, h = int(n), int(e), int(f), int(h)
        if name.startswith("rmat"):      # rmat<scale>:0:0:F:H -- the bench's device RMAT generator
            from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic
            dst, src, n = synthetic.rmat_graph_device(int(name[4:]))
            e = int(dst.shape[0])
        else:
            dst, src = device_powerlaw(n, e)
        g = graph.csr_from_coo(dst, src, n)
        del dst, src
        z = kernels.alloc_table(n, f, "cuda")
        z.normal_()
        el = torch.randn(n, h, device="cuda")
        er = torch.randn(n, h, device="cuda")
        w = torch.rand(e, 1, device="cuda")
        for kind, ncb in [(k, c) for k in args.kinds for c in args.col_blocks]:
            sched = g.schedule(args.chunk, 0 if ncb <= 1 else -(-n // ncb))
            ev = [None] * sched.num_blocks if args.split_launch else None
            # "gat": softmax shifted by the per-block bound (default); "gato": online softmax with a running maximum;
            # a trailing "b": the table stored in bf16
            zz = kernels.to_table(z.to(torch.bfloat16)) if kind.endswith("b") else z
            base = kind[:-1] if kind.endswith("b") else kind
            fn = (lambda: kernels.gat_aggregate(g, el, er, zz, sched=sched, block_events=ev)) if base == "gat" else \
                 (lambda: kernels.gat_aggregate(g, el, er, zz, sched=sched, block_events=ev, bounded=False)) if base == "gato" else \
                 (lambda: kernels.aggregate(g, zz, w, sched=sched, block_events=ev))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(args.iters):
                fn()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / args.iters
            byt = e * (4 + f * zz.element_size() + (h * 4 if kind.startswith("gat") else 4)) + n * (f * 4 + 8)
            print(f"{name:8s} {kind:5s} cb={ncb} chunk={args.chunk} N={n} E={e} F={f} H={h} items={sched.num_items} slots={sched.num_slots} "
                  f"{ms:8.3f} ms  {e / ms / 1e6:7.2f} GTEPS  {byt / ms / 1e6:8.1f} GB/s algorithmic", flush=True)
        del g, z, el, er, w
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
