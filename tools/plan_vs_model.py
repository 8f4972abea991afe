#!/usr/bin/env python
"""Measured B200 time of every golden Cora program beside the reference simulator's modelled latency
(any legal plan executes, and its measured time can be set beside
``simulate()``'s prediction).

    python tools/plan_vs_model.py [--out gpurun_out/plan_vs_model.json]

The modelled side is tests/golden/model_times.json (oracle/gen_model_times.py: the unmodified reference
simulator, GTA architecture, cycles at 1 GHz).  The measured side runs the same ISA program through
``execute()`` on the same synthetic Cora-shape graph: median of CUDA-event timings, once honouring every
STORE_* of the plan (what the plan literally says) and once with dead stores fused away (the default).
The two columns answer different questions -- an ASIC with a 2 MB buffer vs a GPU with 126 MB of L2 --
so the interesting output is the RANKING: which plans the model prefers and which the GPU prefers.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "plan_vs_model.json"))
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import torch
    import yaml

    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph, synthetic

    golden = os.path.join(REPO, "tests", "golden")
    with open(os.path.join(golden, "manifest.json")) as f:
        by_file = {p["file"]: p for p in json.load(f)["programs"]}
    with open(os.path.join(golden, "model_times.json")) as f:
        model = json.load(f)
    n, e, _ = synthetic.SHAPES["cora"]
    coo = synthetic.shape_graph("cora")
    dg = graph.csr_from_coo(coo.dst, coo.src, n)
    dev = lambda d: {k: ([torch.from_numpy(a).cuda() for a in v] if isinstance(v, list) else torch.from_numpy(v).cuda())
                     for k, v in d.items()}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    rows = []
    for m in model["programs"]:
        if "cycles" not in m:
            continue
        p = by_file[m["file"]]
        with open(os.path.join(golden, p["opgraph"])) as f:
            op_info = yaml.safe_load(f)
        with open(os.path.join(golden, p["file"])) as f:
            records = yaml.safe_load(f)
        ni, w, ei = (dev(d) for d in synthetic.opgraph_inputs(op_info, n, e))
        row = {"program": os.path.basename(m["file"])[:-5], "blocks": len(p["op_array"]),
               "model_us": m["cycles"] / 1e3, "model_rw_mb": m["rw_bytes"] / 1e6}
        for label, fuse in (("stores_honoured_us", False), ("fused_us", True)):
            def run():
                return executor.execute(records, op_info, dg, ni, w, ei, network=p["network"], is_reorder=p["reorder"],
                                        fuse_across_blocks=fuse)
            eager = timed(run) * 1e3
            try:        # launch-bound at this size: the headline column is a CUDA-graph replay
                row[label] = timed(executor.GraphedExecution(run).replay) * 1e3
            except Exception as ex:
                print("  (no CUDA graph for %s: %s)" % (row["program"], str(ex).splitlines()[0][:120]))
                torch.cuda.synchronize()
                row[label] = eager
            row[label.replace("_us", "_eager_us")] = eager
        rows.append(row)
        print("%-62s model %10.1f us   B200 stores honoured %8.1f us   fused %8.1f us" %
              (row["program"], row["model_us"], row["stores_honoured_us"], row["fused_us"]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"graph": "synthetic Cora shape N=%d E=%d" % (n, e), "gpu": torch.cuda.get_device_name(0),
                   "timing": "median of %d CUDA-event timings; *_us = CUDA-graph replay, *_eager_us = Python-issued" % args.iters,
                   "rows": rows}, f, indent=1)
    print("written", args.out)


if __name__ == "__main__":
    main()
