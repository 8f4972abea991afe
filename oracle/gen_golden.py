#!/usr/bin/env python
"""Generate tests/golden/ by running the UNMODIFIED reference in a scratch directory.

Run HERE (the container that has /root/reference); the GPU box only sees the committed
fixtures.  Nothing from the reference is copied into the repository: the scratch tree
lives under /tmp and only the reference's OUTPUTS (op-graph YAMLs, ISA programs, tile
tables, compile tuples) are written to tests/golden/.

    python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden]

Harness recipe = SURVEY.md Appendix C: ``code/`` holds vTCAD/code + genGraphOP.py +
changeyaml.py + code/preprocessing.py; CWD-relative ``Network/``, ``dataset/``,
``Results/``; NumPy>=2 shim around ``np.count_nonzero`` for the YAML round trip.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic  # noqa: E402


def _load(path, alias):
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_harness(ref: str) -> str:
    root = tempfile.mkdtemp(prefix="gta_ref_harness_")
    code = os.path.join(root, "code")
    shutil.copytree(os.path.join(ref, "vTCAD", "code"), code, ignore=shutil.ignore_patterns("__pycache__"))
    shutil.copy(os.path.join(ref, "vTCAD", "GraphOP", "genGraphOP.py"), code)
    shutil.copy(os.path.join(ref, "FinalVersion For Paper", "changeyaml.py"), code)
    shutil.copy(os.path.join(ref, "code", "preprocessing.py"), os.path.join(code, "preprocessing.py"))
    shutil.copy(os.path.join(ref, "V2", "GAT_Cora.yaml"), os.path.join(root, "GAT_Cora.yaml"))
    return root


def net_path(network, ds, layer, reorder):
    m = "trans" if reorder else "original"
    return f"Network/{network}/{network}-{ds}/{network}-{m}/{network}-layer{layer}-{m}.yaml"


def opgraph_name(network, ds, layer, reorder):
    return f"{network}-{ds}-layer{layer}-{'trans' if reorder else 'original'}.yaml"


def fix_gcn_trans(path):
    """genGraphOP emits GCN-trans with off-by-one output_lists and a
    1-entry feature_number on a 2-input op; the DATA is corrected (schema unchanged)."""
    with open(path) as f:
        data = yaml.safe_load(f)
    for pos, outs in enumerate([[1], [2], [3], []]):
        data[pos]["OUTPUT"]["output_list"] = outs
    e = data[2]["INPUT"]["feature_number"][0]
    data[2]["INPUT"]["feature_number"] = [e, e]
    with open(path, "w") as f:
        yaml.safe_dump(data, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    os.makedirs(out, exist_ok=True)
    for sub in ("opgraph", "isa", "tiles"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)

    root = build_harness(args.ref)
    os.chdir(root)
    sys.path.insert(0, os.path.join(root, "code"))
    gen = _load(os.path.join(root, "code", "genGraphOP.py"), "ref_genGraphOP")
    chg = _load(os.path.join(root, "code", "changeyaml.py"), "ref_changeyaml")
    interp = _load(os.path.join(root, "code", "interpreter.py"), "ref_interpreter")
    comp = _load(os.path.join(root, "code", "compiler.py"), "ref_compiler")
    prep = _load(os.path.join(root, "code", "preprocessing.py"), "ref_preprocessing")

    manifest = {"programs": [], "tiles": [], "compile": []}

    # ---- 1. op graphs (gen_yaml, unmodified) ------------------------------------------
    shapes = {k: synthetic.SHAPES[k] for k in ("cora", "flickr", "reddit")}
    for ds, (n, e, f) in shapes.items():
        for network in ("GCN", "GAT", "SGC", "GraphSAGE", "GIN", "DGN", "PNA"):
            for layer in (1, 2, 3):
                for reorder in (False, True):
                    p = net_path(network, ds, layer, reorder)
                    gen.gen_yaml(p, n, e, f, network, layer, reorder)
                    if network == "GCN" and reorder:
                        fix_gcn_trans(p)
                    if network in ("GCN", "GAT") or ds == "cora":
                        shutil.copy(p, os.path.join(out, "opgraph", opgraph_name(network, ds, layer, reorder)))
    # re-stamped V2/GAT_Cora.yaml (changeyaml.modify_yaml, H = 4)
    for ds in ("cora", "reddit"):
        n, e, f = shapes[ds]
        p = os.path.join(root, f"GAT_restamped_{ds}.yaml")
        shutil.copy(os.path.join(root, "GAT_Cora.yaml"), p)
        chg.modify_yaml(p, n, e, f, [])
        shutil.copy(p, os.path.join(out, "opgraph", f"GAT-{ds}-restamped-h4.yaml"))

    # BASELINE config 1: the reference's own V1/V2-era fixture, verbatim (9 ops, Cora shape, no COMP_TYPE;
    # V2/simpletest.yaml:14-209).  The executor takes it with the by-position COMP_TYPE list of SURVEY App. A.
    shutil.copy(os.path.join(args.ref, "V2", "simpletest.yaml"), os.path.join(out, "opgraph", "simpletest.yaml"))

    # ---- 2. ISA programs (interpret, unmodified) --------------------------------------
    programs = [
        # (network, ds, layer, reorder, op_array, tile_size_list)
        ("GCN", "cora", 1, False, [[0], [1, 2, 3]], [[2720, 1], [48, 1]]),          # SURVEY 8a example
        ("GCN", "cora", 1, False, [[0], [3], [1, 2]], [[2720, 1], [96, 1], [64, 1]]),  # simulator.py:657-659
        ("GCN", "cora", 1, True, [[0], [1, 2, 3]], [[512, 1], [512, 1]]),
        ("GCN", "flickr", 1, True, [[0], [1, 2, 3]], [[512, 1], [512, 1]]),         # Appendix B1
        ("GCN", "flickr", 2, True, [[0], [1, 2, 3]], [[512, 1], [512, 1]]),
        ("GCN", "reddit", 1, True, [[0], [1, 2, 3]], [[8192, 1], [8192, 1]]),
        ("GAT", "cora", 1, False, [[0, 1, 2], [4, 5, 6, 7, 8], [3, 9, 10, 11, 12, 13]],
         [[176, 1], [2720, 1], [1200, 1]]),                                          # Appendix B3
        ("GAT", "cora", 1, False, [[i] for i in range(14)], [[2720, 1]] * 14),       # no fusion
        ("GAT", "cora", 1, True, [[i] for i in range(13)], [[2720, 1]] * 13),       # no fusion
        ("GAT", "flickr", 1, False, [[0, 1, 2], [4, 5, 6, 7, 8], [3, 9, 10, 11, 12, 13]],
         [[512, 1], [8192, 1], [2048, 1]]),
        ("GAT", "reddit", 1, False, [[13], [0, 1, 2], [3, 9, 10, 11, 12], [4, 5, 6, 7, 8]],
         [[8192, 1], [1024, 1], [1024, 1], [8192, 1]]),                              # Appendix D best plan
        ("GAT", "reddit", 1, False, [[0, 1, 2], [4, 5, 6, 7, 8], [3, 9, 10, 11, 12, 13]],
         [[1024, 1], [8192, 1], [1024, 1]]),
    ]
    # GAT-trans: take the reference compiler's own best plan (hand-written plans can be illegal)
    os.makedirs("dataset/cora", exist_ok=True)
    sizes = prep.gen_size(16, 2720)
    with open("dataset/cora/sizelist_cora.yaml", "w") as f:
        yaml.dump(sizes, f)
    with open("dataset/cora/maxlist_cora.yaml", "w") as f:
        yaml.dump([min(s, 168) for s in sizes], f)
    # (some compile() plans crash the reference's own interpret(); take the first that lowers)
    for network, reorder, layer in (("GAT", True, 1), ("GAT", False, 1), ("SGC", False, 1), ("GraphSAGE", False, 1),
                                    ("GIN", False, 1), ("GAT", False, 2), ("GAT", False, 3), ("GCN", True, 2),
                                    ("GCN", True, 3)):
        for rank, cand in enumerate(comp.compile("cora", network, f"layer{layer}", reorder, False, True, False)[0]):
            try:
                interp.interpret("cora", network, reorder, f"layer{layer}", cand[0], cand[1])
            except Exception:
                continue
            programs.append((network, "cora", layer, reorder, [list(b) for b in cand[0]], [list(t) for t in cand[1]]))
            print("%s cora layer%d reorder=%s: compile() plan rank %d lowers: %s" % (network, layer, reorder, rank, cand[0]))
            break
    # DGN / PNA (COMP_MM on edges): hand-written plans, kept only where the reference's interpret() lowers them
    for network, reorder, n_ops in (("DGN", False, 11), ("PNA", False, 11), ("PNA", True, 11)):
        for plan in ([list(range(n_ops))], [[i] for i in range(n_ops)],
                     [list(range(n_ops // 2)), list(range(n_ops // 2, n_ops))]):
            tiles = [[64, 1]] * len(plan)
            try:
                interp.interpret("cora", network, reorder, "layer2", plan, tiles)
            except Exception as ex:
                print("%s reorder=%s plan of %d blocks: the reference does not lower it (%s)" % (network, reorder, len(plan), type(ex).__name__))
                continue
            programs.append((network, "cora", 2, reorder, plan, tiles))
    for network, ds, layer, reorder, op_array, tiles in programs:
        interp.interpret(ds, network, reorder, f"layer{layer}", op_array, tiles)
        m = "trans" if reorder else "original"
        src = f"Results/Insts/{network}-{ds}-layer{layer}-{m}.yaml"
        tag = "_".join("-".join(map(str, b)) for b in op_array)
        name = f"{network}-{ds}-layer{layer}-{m}__{tag}.yaml"
        shutil.copy(src, os.path.join(out, "isa", name))
        manifest["programs"].append({"file": "isa/" + name, "network": network, "dataset": ds,
                                     "layer": layer, "reorder": reorder, "op_array": op_array,
                                     "tile_size_list": tiles,
                                     "opgraph": "opgraph/" + opgraph_name(network, ds, layer, reorder)})

    # ---- 3. tile tables (calculate_sparsity, unmodified) ------------------------------
    _cnz = np.count_nonzero
    np.count_nonzero = lambda *a, **k: int(_cnz(*a, **k))    # NumPy>=2 YAML shim (Appendix C-1)
    for tag, n, e, seed, sizes in (("g300", 300, 2400, 3, [16, 48, 64, 304]),
                                   ("g97", 97, 600, 5, [1, 7, 32, 97, 128])):
        g = synthetic.powerlaw_graph(n, e, seed=seed, i0=10.0)
        dense = np.zeros((n, n), dtype=np.float32)
        dense[g.dst, g.src] = 1.0
        # sprinkle self loops: the reference zeroes the diagonal (preprocessing.py:17)
        loops = np.arange(0, n, 7)
        dense[loops, loops] = 1.0
        os.makedirs(f"dataset/{tag}", exist_ok=True)
        npy = f"dataset/{tag}/adj_{tag}.npy"
        np.save(npy, dense)
        tables = {}
        for sr in sizes:
            tables[str(sr)] = prep.calculate_sparsity(sr, 1, npy)
        dstc = np.concatenate([g.dst, loops.astype(np.int32)])
        srcc = np.concatenate([g.src, loops.astype(np.int32)])
        np.savez_compressed(os.path.join(out, "tiles", f"{tag}.npz"), num_nodes=n, dst=dstc, src=srcc,
                            **{f"table_{k}": np.asarray(v, dtype=np.int64) for k, v in tables.items()})
        # maxlist / sizelist through the reference's own YAML round trip
        for sr in sizes:
            prep.save(tables[str(sr)], f"dataset/{tag}/adj_{tag}_{sr}_1.yaml")
        maxlist = [prep.cal_min_sparsity(tag, sr) for sr in sizes]
        # the reference's own on-disk files for one graph (byte-for-byte targets of graph.write_tile_tables)
        if tag == "g97":
            prep.save(sizes, f"dataset/{tag}/sizelist_{tag}.yaml")
            prep.save(maxlist, f"dataset/{tag}/maxlist_{tag}.yaml")
            for name in [f"adj_{tag}_{sr}_1.yaml" for sr in sizes] + [f"sizelist_{tag}.yaml", f"maxlist_{tag}.yaml"]:
                shutil.copy(f"dataset/{tag}/{name}", os.path.join(out, "tiles", name))
        manifest["tiles"].append({"file": f"tiles/{tag}.npz", "sizes": sizes, "maxlist": [int(v) for v in maxlist],
                                  "gen_size_16_100": prep.gen_size(16, 100)})
    np.count_nonzero = _cnz

    # ---- 4. compile anchors -----------------------------------------------------------
    # GAT/Cora layer 1, no fusion: rw = 58 978 768 (code/genetic_algorithm.py:68); adjacency
    # independent when every tile >= N, so any maxlist works.  Use a flat table.
    for network, reorder in (("GAT", False), ("GAT", True), ("GCN", False), ("GCN", True)):
        res = comp.compile("cora", network, "layer1", reorder, False, True, False)[0]
        keep = [res[0], res[-1]] + [r for r in res if set(r[3]) == {"0"}]
        manifest["compile"].append({
            "network": network, "reorder": reorder, "dataset": "cora", "layer": 1, "num_plans": len(res),
            "plans": [{"fused_array": r[0], "tile_sizes": [list(t) for t in r[1]], "rw": int(r[2]), "pattern": r[3]}
                      for r in keep]})

    with open(os.path.join(out, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("golden fixtures written to", out)
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
