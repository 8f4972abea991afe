#!/usr/bin/env python
"""The reference's own pipeline, timed as shipped: compile -> interpret -> simulate
(vTCAD/code/test.py:10-15) on the Cora-shape GCN layer 1, the flow BASELINE.md section 4 item 1 asks to be
timed on the GPU box's host beside the GPU numbers.

    python tools/reference_pipeline.py --install     # here: copy the needed reference files to baseline/_ref
    python tools/reference_pipeline.py               # anywhere baseline/_ref exists: time the pipeline, print JSON

``baseline/_ref`` is git-ignored (nothing of the reference enters the history) but travels to the GPU box with
the gpurun snapshot.  The reference is run UNMODIFIED, in a scratch working directory laid out the way its
CWD-relative paths expect: ``code/`` = vTCAD/code + genGraphOP.py, ``Network/``,
``dataset/``, ``Results/``.  ``compile`` reads maxlist/sizelist tables of all 170 tile sizes, which the
reference's preprocessing needs 141 s (dense N^2) to produce; here they are written flat (every tile >= N makes
the plan independent of them, as in oracle/gen_golden.py) and only the ONE adjacency table the simulated plan
needs is produced by the reference's own ``calculate_sparsity`` (timed separately).  Single Python thread: the
reference cannot use more.  ``simulate`` is a per-cycle Python loop that does not terminate on every program;
the pipeline runs in a child process with a time limit.
"""
from __future__ import annotations

import argparse
import contextlib
import importlib.util
import io
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
DEFAULT_DST = os.path.join(REPO, "baseline", "_ref")
PLAN = {"network": "GCN", "layer": 1, "reorder": False, "op_array": [[0], [1, 2, 3]], "tile_size_list": [[2720, 1], [48, 1]]}


def install(ref: str = "/root/reference", dst: str = DEFAULT_DST) -> str:
    """Copy what the pipeline imports (and nothing else) from the reference checkout into ``dst``."""
    code = os.path.join(dst, "code")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(os.path.join(ref, "vTCAD", "code"), code, ignore=shutil.ignore_patterns("__pycache__", "*.csv"))
    shutil.copy(os.path.join(ref, "vTCAD", "GraphOP", "genGraphOP.py"), code)
    shutil.copy(os.path.join(ref, "code", "preprocessing.py"), os.path.join(code, "preprocessing.py"))
    with open(os.path.join(dst, "README"), "w") as f:
        f.write("Unmodified files of the reference (vTCAD/code, vTCAD/GraphOP/genGraphOP.py, code/preprocessing.py), copied by\n"
                "tools/reference_pipeline.py --install for the reference-pipeline timing of bench.py.  Git-ignored.\n")
    return dst


def _load(path, alias):
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(ref_dir: str) -> dict:
    """Child process body: build the scratch layout, run the three stages, return the timings."""
    import numpy as np
    import yaml
    sys.path.insert(0, REPO)
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic
    root = tempfile.mkdtemp(prefix="gta_ref_pipeline_")
    try:
        shutil.copytree(os.path.join(ref_dir, "code"), os.path.join(root, "code"))
        os.chdir(root)
        sys.path.insert(0, os.path.join(root, "code"))
        gen = _load("code/genGraphOP.py", "ref_genGraphOP")
        comp = _load("code/compiler.py", "ref_compiler")
        interp = _load("code/interpreter.py", "ref_interpreter")
        prep = _load("code/preprocessing.py", "ref_preprocessing")
        sim = _load("code/simulator.py", "ref_simulator")
        n, e, f = synthetic.SHAPES["cora"]
        g = synthetic.shape_graph("cora")
        network, layer, reorder = PLAN["network"], PLAN["layer"], PLAN["reorder"]
        m = "trans" if reorder else "original"
        path = f"Network/{network}/{network}-cora/{network}-{m}/{network}-layer{layer}-{m}.yaml"
        gen.gen_yaml(path, n, e, f, network, layer, reorder)
        os.makedirs("dataset/cora", exist_ok=True)
        sizes = prep.gen_size(16, 2720)
        with open("dataset/cora/sizelist_cora.yaml", "w") as fh:
            yaml.dump(sizes, fh)
        with open("dataset/cora/maxlist_cora.yaml", "w") as fh:
            yaml.dump([min(s, 168) for s in sizes], fh)
        dense = np.zeros((n, n), dtype=np.float32)
        dense[g.dst, g.src] = 1.0
        np.save("dataset/cora/adj_cora.npy", dense)
        _cnz = np.count_nonzero
        np.count_nonzero = lambda *a, **k: int(_cnz(*a, **k))      # NumPy >= 2 YAML shim
        sink = io.StringIO()
        out = {"plan": PLAN, "graph": "synthetic.shape_graph('cora')", "threads": 1}
        with contextlib.redirect_stdout(sink):
            t0 = time.perf_counter()
            for sr in sorted({t[0] for t in PLAN["tile_size_list"]}):
                prep.save(prep.calculate_sparsity(sr, 1, "dataset/cora/adj_cora.npy"), f"dataset/cora/adj_cora_{sr}_1.yaml")
            out["preprocess_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            res = comp.compile("cora", network, f"layer{layer}", reorder, False, True, False)
            out["compile_s"] = time.perf_counter() - t0
            out["compile_plans"] = len(res[0])
            t0 = time.perf_counter()
            interp.interpret("cora", network, reorder, f"layer{layer}", PLAN["op_array"], PLAN["tile_size_list"])
            out["interpret_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            cycles, rw = sim.simulate(PLAN["tile_size_list"], "cora", network, f"layer{layer}", reorder, False, True, "GTA")
            out["simulate_s"] = time.perf_counter() - t0
        np.count_nonzero = _cnz
        out["modelled_cycles"] = int(cycles)
        out["modelled_rw_bytes"] = int(rw)
        out["modelled_ms_at_1GHz"] = cycles / 1e6
        out["note"] = ("unmodified reference, one Python thread; preprocess_s covers only the %d adjacency tables this plan "
                       "needs (all 170 tile sizes: 141 s, SURVEY.md section 3.1)" % len({t[0] for t in PLAN["tile_size_list"]}))
        return out
    finally:
        os.chdir(REPO)
        shutil.rmtree(root, ignore_errors=True)


def time_pipeline(ref_dir: str = DEFAULT_DST, timeout: int = 240) -> dict:
    """Run the pipeline in a child process (the simulator may spin) and return its timing record."""
    if not os.path.isdir(os.path.join(ref_dir, "code")):
        raise FileNotFoundError(f"{ref_dir} is missing: run `python tools/reference_pipeline.py --install` where /root/reference exists")
    proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", ref_dir], capture_output=True, text=True,
                          timeout=timeout)
    if proc.returncode != 0:
        raise RuntimeError("reference pipeline failed: " + (proc.stderr.strip().splitlines() or ["?"])[-1])
    return json.loads(proc.stdout.strip().splitlines()[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--install", action="store_true")
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--dst", default=DEFAULT_DST)
    ap.add_argument("--child", default=None, help="(internal) run the stages in this process")
    args = ap.parse_args()
    if args.child:
        print(json.dumps(_run(args.child)))
        return
    if args.install:
        print("installed", install(args.ref, args.dst))
        return
    print(json.dumps(time_pipeline(args.dst), indent=1))


if __name__ == "__main__":
    main()
