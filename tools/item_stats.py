#!/usr/bin/env python
"""Work-list statistics on the CPU (no GPU): how the edges of a shape spread over the items of the aggregation
work list -- at most `chunk` consecutive edges of one destination row inside one column block -- and how many of the
kernel's 8-edge batches are partly empty.  Decides whether packing several short rows into one item is worth building.

    python tools/item_stats.py [--shape reddit] [--blocks 3] [--chunk 1024] [--parts 1]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--blocks", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--parts", type=int, default=1, help="destination-range partitions (stats for partition 0)")
    args = ap.parse_args()
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic
    n, e, _ = synthetic.SHAPES[args.shape]
    g = synthetic.shape_graph(args.shape)
    dst, src = g.dst.astype(np.int64), g.src.astype(np.int64)
    if args.parts > 1:
        deg = np.bincount(dst, minlength=n)
        indptr = np.concatenate([[0], np.cumsum(deg)])
        b = np.searchsorted(indptr, (np.arange(args.parts + 1) * int(indptr[-1])) // args.parts)   # edge-balanced bounds
        b[0], b[-1] = 0, n
        keep = dst < b[1]
        dst, src = dst[keep], src[keep]
    col = -(-n // args.blocks)
    key = dst * args.blocks + src // col                    # (row, column block) segment id
    seg = np.bincount(key, minlength=n * args.blocks)
    seg = seg[seg > 0]
    full, rest = np.divmod(seg, args.chunk)
    items = np.concatenate([np.full(int(full.sum()), args.chunk), rest[rest > 0]])
    batches = -(-items // args.batch)
    slots = batches * args.batch
    print(f"shape {args.shape}: {len(dst)} edges, {args.blocks} column blocks, chunk {args.chunk}, partition 1/{args.parts}")
    print(f"items {len(items)}  mean edges/item {items.mean():.1f}  median {np.median(items):.0f}")
    for t in (1, 2, 4, 8, 16, 32, 64):
        m = items <= t
        print(f"  items with <= {t:3d} edges: {m.mean() * 100:5.1f} % of items, {items[m].sum() / items.sum() * 100:5.2f} % of edges")
    print(f"batch slots (x{args.batch}) issued {slots.sum()}  filled {items.sum()}  -> {items.sum() / slots.sum() * 100:.1f} % full")
    rows_touched = len(np.unique(dst))
    print(f"rows with edges {rows_touched}; items per such row {len(items) / rows_touched:.2f} "
          f"(rows needing the combine pass: {(np.bincount((key // args.blocks))[np.unique(dst)] > 0).sum()})")


if __name__ == "__main__":
    main()
