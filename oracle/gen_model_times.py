#!/usr/bin/env python
"""Modelled latency of the golden Cora programs: the UNMODIFIED reference simulator run here.

    python oracle/gen_model_times.py [--ref /root/reference]      # -> tests/golden/model_times.json

For every Cora-shape ISA program of tests/golden/manifest.json this drives the reference's own flow in a
scratch directory (same harness as gen_golden.py): its ``preprocessing.calculate_sparsity`` on the dense
adjacency of the synthetic Cora-shape graph (synthetic.shape_graph("cora"), the graph the GPU runs), its
``interpret`` for the plan, then its ``simulate(tile_size_list, 'cora', network, layer, isReorder, False,
True, 'GTA')`` (vTCAD/code/simulator.py:423) -- a per-cycle Python loop, tens of seconds per plan.  Only
the numbers it returns, ``(cycles, rw)``, are committed; tools/plan_vs_model.py sets the time the B200
measures for the same program beside them.  Test infrastructure, like the
rest of oracle/.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import gen_golden  # noqa: E402
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden", "model_times.json"))
    ap.add_argument("--max-programs", type=int, default=64)
    ap.add_argument("--timeout", type=int, default=420, help="seconds per program (the simulator can spin forever)")
    ap.add_argument("--only", type=int, default=None, help="(internal) simulate program number ONLY in this process")
    args = ap.parse_args()
    with open(os.path.join(REPO, "tests", "golden", "manifest.json")) as f:
        programs = [p for p in json.load(f)["programs"] if p["dataset"] == "cora"][:args.max_programs]
    if args.only is None:
        # one child process per program: the reference's simulate() is a `while True` over cycles and
        # does not terminate on every program its own interpreter emits
        out = []
        for i, p in enumerate(programs):
            part = args.out + f".part{i}"
            try:
                subprocess.run([sys.executable, os.path.abspath(__file__), "--ref", args.ref, "--out", part,
                                "--only", str(i), "--max-programs", str(args.max_programs)], timeout=args.timeout, check=True)
                with open(part) as f:
                    out.extend(json.load(f)["programs"])
            except subprocess.TimeoutExpired:
                print(f"{p['file']}: reference simulate() did not finish within {args.timeout} s", flush=True)
                out.append({"file": p["file"], "error": f"simulate() did not finish within {args.timeout} s"})
            except subprocess.CalledProcessError as ex:
                out.append({"file": p["file"], "error": f"child exited {ex.returncode}"})
            finally:
                if os.path.exists(part):
                    os.remove(part)
        _write(args.out, out)
        return
    programs = [programs[args.only]]

    root = gen_golden.build_harness(args.ref)
    os.chdir(root)
    sys.path.insert(0, os.path.join(root, "code"))
    load = gen_golden._load
    gen = load(os.path.join(root, "code", "genGraphOP.py"), "ref_genGraphOP")
    interp = load(os.path.join(root, "code", "interpreter.py"), "ref_interpreter")
    prep = load(os.path.join(root, "code", "preprocessing.py"), "ref_preprocessing")
    sim = load(os.path.join(root, "code", "simulator.py"), "ref_simulator")

    n, e, f = synthetic.SHAPES["cora"]
    g = synthetic.shape_graph("cora")
    dense = np.zeros((n, n), dtype=np.float32)
    dense[g.dst, g.src] = 1.0
    os.makedirs("dataset/cora", exist_ok=True)
    np.save("dataset/cora/adj_cora.npy", dense)
    _cnz = np.count_nonzero
    np.count_nonzero = lambda *a, **k: int(_cnz(*a, **k))      # NumPy>=2 YAML shim
    have = set()
    out = []
    for p in programs:
        network, layer, reorder = p["network"], p["layer"], p["reorder"]
        path = gen_golden.net_path(network, "cora", layer, reorder)
        gen.gen_yaml(path, n, e, f, network, layer, reorder)
        if network == "GCN" and reorder:
            gen_golden.fix_gcn_trans(path)
        for sr, _ in p["tile_size_list"]:
            if sr not in have:
                prep.save(prep.calculate_sparsity(sr, 1, "dataset/cora/adj_cora.npy"), f"dataset/cora/adj_cora_{sr}_1.yaml")
                have.add(sr)
        t0 = time.perf_counter()
        sink = io.StringIO()
        try:
            with contextlib.redirect_stdout(sink):
                interp.interpret("cora", network, reorder, f"layer{layer}", p["op_array"], p["tile_size_list"])
                cycles, rw = sim.simulate(p["tile_size_list"], "cora", network, f"layer{layer}", reorder, False, True, "GTA")
        except Exception as ex:      # the simulator is not total over its own ISA
            print(f"{p['file']}: reference simulate() failed: {type(ex).__name__}: {ex}", flush=True)
            out.append({"file": p["file"], "error": f"{type(ex).__name__}: {ex}"})
            continue
        wall = time.perf_counter() - t0
        print(f"{p['file']}: {cycles} cycles, rw {rw}, {wall:.1f} s of simulation", flush=True)
        out.append({"file": p["file"], "network": network, "layer": layer, "reorder": reorder, "cycles": int(cycles),
                    "rw_bytes": int(rw), "simulate_wall_s": round(wall, 2)})
    np.count_nonzero = _cnz
    _write(args.out, out)
    import shutil
    shutil.rmtree(root, ignore_errors=True)


def _write(path, out):
    with open(path, "w") as f_:
        json.dump({"graph": "synthetic.shape_graph('cora')", "architecture": "GTA, isFlexibleHardware=True",
                   "note": "cycles at the modelled 1 GHz clock = ns", "programs": out}, f_, indent=1)
    print("written", path, flush=True)


if __name__ == "__main__":
    main()
