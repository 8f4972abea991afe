#!/usr/bin/env python
"""Random op graphs under random plans through execute() on the GPU, against the oracle: the device-side run of
tests/test_cpu_executor_fuzz.py (which swaps the kernels for a CPU test double).  tests/test_gpu_z_widen.py runs the
first 60 cases; as a script it runs as many as asked (`tools/gpu.sh -- 'python tests/device_fuzz.py --cases 300'`).
Round 1: 150 cases on a B200, 0 failed, 0 unsupported.

Prints one line per failing case (seed, op, error, kernel log) and a summary; exit code 1 on any failure.
"""
from __future__ import annotations

import argparse
import os
import random
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))      # lives under tests/: it uses the oracle as its checker


def run_cases(first: int, cases: int):
    """Returns (failed, unsupported, messages)."""
    import torch

    import test_cpu_executor_fuzz as F
    from oracle import gta_oracle as O
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import _cabi, executor, graph, isa, lowering, synthetic

    g = synthetic.powerlaw_graph(F.N, F.E, seed=3, i0=6.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, F.N)
    dg = graph.csr_from_coo(g.dst, g.src, F.N)
    max_deg = int(max(np.diff(indptr).max(), np.bincount(indices, minlength=F.N).max()))
    up = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    bad = unsupported = 0
    messages = []
    for seed in range(first, first + cases):
        rng = random.Random(1000 + seed)
        op_info, sem_x, sem_o = F._random_graph(rng, F.N, g.num_edges, max_deg)
        n_ops = len(op_info)
        plan, tiles = F._random_plan(rng, n_ops)
        try:
            records = lowering.lower(op_info, plan, tiles, F.N)
            isa.Program.from_records(records).block_ops(op_info)
        except (lowering.LoweringError, isa.IsaError):
            plan, tiles = [[i] for i in range(n_ops)], [[32, 1]] * n_ops
            records = lowering.lower(op_info, plan, tiles, F.N)
        data = np.random.default_rng(seed)
        ni, w, ei = {}, {}, {}
        for pos, op in enumerate(op_info):
            win = op["INPUT"]["size_per_feature"][0] // 4
            if op["COMP_TYPE"] == "MM":
                w[pos] = data.uniform(-0.3, 0.3, size=(win, op["OUTPUT"]["size_per_feature"] // 4)).astype(np.float32)
            if not op["INPUT"]["input_g_list"]:
                ni[pos] = data.uniform(-0.5, 0.5, size=(F.N, win)).astype(np.float32)
            if -1 in op["INPUT"]["input_g_list"]:
                ei[pos] = data.uniform(0.1, 1.0, size=(g.num_edges, 1)).astype(np.float32)
        ref, ref_scale = O.run_opgraph(op_info, indptr, indices, ni, w, ei, semantics=sem_o, stabilize=False,
                                       fix_gat_op10=False, return_scale=True)
        try:
            out, log = executor.execute(records, op_info, dg, up(ni), up(w), up(ei), semantics=sem_x, stabilize=False,
                                        fuse_across_blocks=bool(seed % 2), return_log=True)
        except _cabi.GtaUnsupported as ex:
            unsupported += 1
            messages.append(f"seed {seed}: unsupported: {ex}")
            continue
        except Exception as ex:
            bad += 1
            messages.append(f"seed {seed}: {type(ex).__name__}: {ex}")
            continue
        for p, y in out.items():
            want = ref[p] if ref[p].ndim == 2 else ref[p][:, None]
            got = y.cpu().numpy()
            sc = ref_scale[p] if ref_scale[p].ndim == 2 else ref_scale[p][:, None]
            # the stated fp32 tolerance: 1e-5 |y64| + 1e-5 rowscale (first-order error scale of the op graph)
            worst = float(np.max(np.abs(got - want) / (1e-5 * np.abs(want) + 1e-5 * sc + 1e-30))) \
                if got.shape == want.shape and np.all(np.isfinite(got)) else float("inf")
            if not worst <= 1.0:
                bad += 1
                messages.append(f"seed {seed}: op {p} error {worst:.2f}x the tolerance, plan {plan} kernels {log}")
                break
    return bad, unsupported, messages


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--first", type=int, default=0)
    args = ap.parse_args()
    bad, unsupported, messages = run_cases(args.first, args.cases)
    print("\n".join(messages))
    print(f"{args.cases} cases: {bad} failed, {unsupported} unsupported")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
