"""Op graph + plan -> message-passing ISA program: host mirror of the reference's ``interpret()``.

The reference's emitter (vTCAD/code/interpreter.py:809-853) stays the front-end of record; this module
restates it from scratch so a box that has only this package can go ``op graph -> ISA -> execute()``,
and so the instruction format is pinned by tests: for every plan in ``tests/golden/isa`` the YAML this
module writes is byte-identical to what the unmodified reference wrote (tests/test_cpu_lowering.py).

Same entry points, argument meaning and failure modes:

* ``interpret(data_set, network, isReorder, layer, op_array, tile_size_list)`` reads
  ``Network/<net>/<net>-<ds>/<net>-<map>/<net>-<layer>-<map>.yaml`` and ``code/hardware_info.yaml``
  relative to the CWD and writes ``Results/Insts/<net>-<ds>-<layer>-<map>.yaml`` (interpreter.py:809-853);
  only ``cora | pubmed | flickr | reddit`` have a node count (interpreter.py:811-819).
* ``lower(op_info, op_array, tile_size_list, node_num, fusable)`` is the same pipeline on in-memory data.
* ``gen_inst(op_info, op_id, fused_array, TR, TC, SR, SC) -> [loads, comp, store]`` (interpreter.py:313-479).

What the pipeline does (per fused block, ``SR x SC`` tiles, ``TR = ceil(N/SR)``, ``TC = ceil(N/SC)``):
every op becomes LOAD_* / COMP_* (FETCH for a scatter) / STORE_* records with RAW/WAR token ratios
(``TILE_RULES``, ``RATE_RULES`` below); compute pairs listed as fusable in ``hardware_info.yaml`` merge into
``COMP_x_COMP_y`` records; FETCH records are removed and their producers spliced to their consumers.
Quirks that shape the output are kept on purpose and marked ``reference quirk``.
"""
from __future__ import annotations

import math
import os

import yaml

NODE_COUNT = {"cora": 2708, "pubmed": 19717, "flickr": 89250, "reddit": 232965}     # interpreter.py:811-819
UNIT_OF = {"ADD": "VEC_ALU", "SF": "SF_ALU", "MUL": "VEC_ALU", "MM": "MM"}         # interpreter.py:5-14
MEMORY_UNIT = "Memory_Access_Unit"
VIRTUAL_UNIT = "Virtual_Loader"

#: fusable (producer kind, consumer kind, producer COMP, consumer COMP) of the shipped hardware_info.yaml:11-68
DEFAULT_FUSABLE = {
    (("scatter", "gather"), ("NONE", "ADD")), (("scatter", "applyedge"), ("NONE", "MM")),
    (("scatter", "applyedge"), ("NONE", "ADD")), (("applyedge", "gather"), ("MM", "ADD")),
    (("applyedge", "gather"), ("MUL", "ADD")),
}


class LoweringError(ValueError):
    """The plan cannot be lowered (the reference raises TypeError / IndexError at the same points)."""


class _NoAliasDumper(yaml.SafeDumper):
    def ignore_aliases(self, data):
        return True


def load_fusable(path: str) -> set:
    """``Inst_fused`` entries with ``Is_Fused: True`` (interpreter.py:508-521)."""
    with open(path) as f:
        data = yaml.load(f, Loader=yaml.FullLoader)
    out = set()
    for item in data:
        for entry in item.get("Inst_fused", []) if isinstance(item, dict) else []:
            if entry["Is_Fused"]:
                out.add((tuple(entry["Pattern"]), tuple(entry["Compute_Type"])))
    return out


# ---- tile and rate rules --------------------------------------------------------------------------

def _tiles(role: str, kind: str, order: str, TR: int, TC: int, SR: int, SC: int):
    """(Tile_Times, Tile_Size) of a load / comp / store record (interpreter.py:55-129)."""
    edge = (TR * TC, SR * SC)
    node = (TR, SR) if order == "R" else (TC, SC)
    if role == "comp":
        return {"scatter": edge, "gather": edge, "applyedge": edge, "applynode": node}[kind]
    if role == "load":
        if kind == "scatter":
            return (TR, SR) if order == "R" else (TR * TC, SC)
        return {"gather": edge, "applyedge": edge, "applynode": node}[kind]
    if kind in ("gather", "applynode"):                     # store
        return (TR, SR) if order == "R" else (TR * TC, SC)
    return edge


def _rate(kind: str, order: str, consumer: str, TR: int, TC: int):
    """Tokens a producer of ``kind`` emits per tokens its ``consumer`` takes (interpreter.py:165-194)."""
    if kind == "scatter" and consumer in ("gather", "applyedge"):
        return [1, 1]
    if kind == "gather" and consumer == "applynode":
        return [TC, 1] if order == "R" else [TR, 1]
    if kind == "applyedge" and consumer in ("gather", "applyedge"):
        return [1, 1]
    if kind == "applynode" and consumer == "scatter":
        return [1, TC] if order == "R" else [1, TR]
    if kind == "applynode" and consumer == "applynode":
        return [1, 1]
    raise LoweringError(f"no token rate from a {kind} to a {consumer} inside one block")


def _dep(inst: dict, times) -> dict:
    return {"TYPE": inst["TYPE"], "ID": inst["ID"], "Times": list(times)}


def _record(typ, ident, unit, tiles, length, weight=None) -> dict:
    rec = {"TYPE": typ, "ID": ident, "Hardware_Unit": unit, "Tile_Times": tiles[0], "Tile_Size": tiles[1],
           "Feature_Length": length}
    if weight is not None:
        rec["Weight_Size"] = weight
    rec["Dependency"] = {"RAW": [], "WAR": []}
    rec["Enable"] = {"RAW": [], "WAR": []}
    return rec


def _comp_record(op_info, pos, TR, TC, SR, SC) -> dict:
    op = op_info[pos]
    kind, comp = op["TYPE"], op["COMP_TYPE"]
    typ = "FETCH" if kind == "scatter" else "COMP_" + comp
    return _record(typ, f"{pos}_{kind}_0", UNIT_OF.get(comp), _tiles("comp", kind, op["ORDER"], TR, TC, SR, SC),
                   op["INPUT"]["size_per_feature"][0], weight=0)


def _load_record(op_info, pos, slot, typ, TR, TC, SR, SC) -> dict:
    op = op_info[pos]
    kind, order = op["TYPE"], op["ORDER"]
    times, size = _tiles("load", kind, order, TR, TC, SR, SC)
    unit = MEMORY_UNIT
    g_num = op["INPUT"]["input_g_num"]
    if typ == "LOAD_W":
        length, ident, times, size = op["INPUT"]["input_size"][0], f"{pos}_{kind}_{g_num}", 1, 1
    elif typ == "LOAD_N" and kind == "gather":              # the accumulator of a gather
        length, ident = op["OUTPUT"]["size_per_feature"], f"{pos}_{kind}_{g_num}"
        if order == "R":
            unit, times = VIRTUAL_UNIT, TR
        else:
            times = TC
    else:
        length, ident = op["INPUT"]["size_per_feature"][slot], f"{pos}_{kind}_{slot}"
    return _record(typ, ident, unit, (times, size), length)


def _store_record(op_info, pos, TR, TC, SR, SC) -> dict:
    op = op_info[pos]
    kind = op["TYPE"]
    typ = "STORE_E" if kind in ("scatter", "applyedge") else "STORE_N"
    return _record(typ, f"{pos}_{kind}_0", MEMORY_UNIT, _tiles("store", kind, op["ORDER"], TR, TC, SR, SC),
                   op["OUTPUT"]["size_per_feature"])


def _link(first: dict, second: dict, times) -> None:
    """``first`` feeds ``second``: WAR on the producer, RAW on the consumer, mirrored into Enable."""
    t = list(times)
    first["Dependency"]["WAR"].append(_dep(second, t))
    first["Enable"]["RAW"].append(_dep(second, t))
    second["Dependency"]["RAW"].append(_dep(first, [t[1], t[0]]))
    second["Enable"]["WAR"].append(_dep(first, [t[1], t[0]]))


def _letter(kind: str) -> str:
    return kind[5].upper()        # apply(e)dge -> E, apply(n)ode -> N


def gen_inst(op_info, op_id, fused_array, TR, TC, SR, SC):
    """Records of one op inside a fused block: ``[loads, comp, store or []]`` (interpreter.py:313-479)."""
    op = op_info[op_id]
    kind, order, comp_t = op["TYPE"], op["ORDER"], op["COMP_TYPE"]
    inputs, g_num, outputs = op["INPUT"]["input_g_list"], op["INPUT"]["input_g_num"], op["OUTPUT"]["output_list"]
    comp = _comp_record(op_info, op_id, TR, TC, SR, SC)
    loads = []
    per_node = [1, TR] if order == "R" else [1, TC]

    # the op's own side input: accumulator of a gather, weights of an MM, an undeclared external operand
    own = None
    if kind == "gather":
        own = ("LOAD_N", [1, TC] if order == "R" else [1, 1])
    elif kind == "applyedge" and comp_t == "MM":
        own = ("LOAD_W", [1, TR * TC])
    elif kind == "applynode" and comp_t == "MM":
        own = ("LOAD_W", per_node)
    elif kind == "applyedge" and len(inputs) != g_num:
        own = ("LOAD_E", [1, 1])
    elif kind == "applynode" and len(inputs) != g_num:
        own = ("LOAD_N", per_node)
    if own is not None:
        ld = _load_record(op_info, op_id, g_num - 1, own[0], TR, TC, SR, SC)
        loads.append(ld)
        if own[0] == "LOAD_W":
            comp["Weight_Size"] = ld["Feature_Length"]
        _link(ld, comp, own[1])

    def external_load(slot, n_declared):
        if kind == "scatter":
            typ, times = "LOAD_N", ([1, TC] if order == "R" else [1, 1])
        elif kind == "gather":
            typ, times = "LOAD_E", [1, 1]
        else:
            times = [1, 1]
            kinds = ["LOAD_" + _letter(kind)] * (1 if comp_t == "MM" else n_declared)
            if slot >= len(kinds):
                raise LoweringError(f"op {op_id}: no load type for input slot {slot}")
            typ = kinds[slot]
        ld = _load_record(op_info, op_id, slot, typ, TR, TC, SR, SC)
        loads.append(ld)
        _link(ld, comp, times)

    if not inputs:
        # reference quirk: a non-MM apply op without declared inputs gets no load at all
        if kind in ("scatter", "gather") or comp_t == "MM":
            external_load(0, 1)
    else:
        for slot, producer in enumerate(inputs):
            if producer in fused_array:
                p = op_info[producer]
                t = _rate(p["TYPE"], p["ORDER"], kind, TR, TC)
                pc = _comp_record(op_info, producer, TR, TC, SR, SC)
                comp["Dependency"]["RAW"].append(_dep(pc, [t[1], t[0]]))
                comp["Enable"]["WAR"].append(_dep(pc, [t[1], t[0]]))
            else:
                external_load(slot, len(inputs))

    store = []

    def add_store():
        st = _store_record(op_info, op_id, TR, TC, SR, SC)
        times = [1, TC] if (kind == "gather" and order == "R") else [1, 1]
        st["Dependency"]["RAW"].append(_dep(comp, times))
        st["Enable"]["WAR"].append(_dep(comp, times))
        comp["Dependency"]["WAR"].append(_dep(st, [times[1], times[0]]))
        comp["Enable"]["RAW"].append(_dep(st, [times[1], times[0]]))
        return st

    if not outputs:
        store = add_store()
    else:
        for consumer in outputs:
            if consumer in fused_array:
                t = _rate(kind, order, op_info[consumer]["TYPE"], TR, TC)
                cc = _comp_record(op_info, consumer, TR, TC, SR, SC)
                comp["Dependency"]["WAR"].append(_dep(cc, t))
                comp["Enable"]["RAW"].append(_dep(cc, t))
            elif store == []:
                store = add_store()
    return [loads, comp, store]


# ---- instruction fusion and FETCH elimination --------------------------------------------------------

def _find(blocks, typ, ident):
    """Index INSIDE ITS BLOCK of the first record with this TYPE and ID (interpreter.py:481-485)."""
    for block in blocks:
        for j, inst in enumerate(block):
            if inst["TYPE"] == typ and inst["ID"] == ident:
                return j
    return None


def _second_field(text: str):
    parts = text.split("_")
    return parts[1] if len(parts) > 1 else None


def _pair_is_fusable(a: dict, b: dict, fusable: set) -> bool:
    if not (a["TYPE"].startswith("COMP") and b["TYPE"].startswith("COMP")):
        return False
    key = ((_second_field(a["ID"]), _second_field(b["ID"])), (_second_field(a["TYPE"]), _second_field(b["TYPE"])))
    return key in fusable


def _fuse_pairs(blocks, fusable):
    successors = [[[_find(blocks, d["TYPE"], d["ID"]) for d in inst["Dependency"]["WAR"]] for inst in block]
                  for block in blocks]
    pairs = []
    for block, succ in zip(blocks, successors):
        found = []
        for j, inst in enumerate(block):
            if not inst["TYPE"].startswith("COMP_") or len(succ[j]) != 1:
                continue
            nxt = succ[j][0]
            if _pair_is_fusable(inst, block[nxt], fusable):
                if len(succ[nxt]) == 1 and _pair_is_fusable(block[nxt], block[succ[nxt][0]], fusable):
                    raise LoweringError("three-way instruction fusion is not emitted by the reference (interpreter.py:567)")
                found.append((j, nxt))
        pairs.append(found)
    return pairs


def _rename_into(block, fused: dict) -> None:
    """Point every dependency on a merged record at the fused one.  reference quirk: the match is by
    SUBSTRING of ID and TYPE (interpreter.py:721-738), so '1_...' also matches inside '11_...'."""
    for inst in block:
        for group in (inst["Dependency"], inst["Enable"]):
            for deps in (group["RAW"], group["WAR"]):
                for d in deps:
                    if d["ID"] in fused["ID"] and d["TYPE"] in fused["TYPE"]:
                        d["ID"], d["TYPE"] = fused["ID"], fused["TYPE"]


def _merge(block, i: int, j: int) -> dict:
    a, b = block[i], block[j]
    raw = list(a["Dependency"]["RAW"]) + [d for d in b["Dependency"]["RAW"]
                                          if not (d["TYPE"] == a["TYPE"] and d["ID"] == a["ID"])]
    war = [d for d in a["Dependency"]["WAR"] if not (d["TYPE"] == b["TYPE"] and d["ID"] == b["ID"])] \
        + list(b["Dependency"]["WAR"])
    unit = "MM" if "COMP_MM" in (a["TYPE"], b["TYPE"]) else "VEC_ALU"
    fused = {"TYPE": a["TYPE"] + "_" + b["TYPE"], "ID": a["ID"] + "_" + b["ID"], "Hardware_Unit": unit,
             "Tile_Times": a["Tile_Times"], "Tile_Size": a["Tile_Size"], "Feature_Length": a["Feature_Length"],
             "Dependency": {"RAW": raw, "WAR": war},
             "Enable": {"RAW": war, "WAR": raw}}          # the SAME lists: later edits show on both sides
    _rename_into(block, fused)
    return fused


def _apply_fusion(blocks, pairs) -> None:
    merged = [[_merge(block, i, j) for i, j in found] for block, found in zip(blocks, pairs)]
    for block, found, new in zip(blocks, pairs, merged):
        for k in range(len(found) - 1, -1, -1):
            for idx in sorted(found[k], reverse=True):
                del block[idx]
            block.append(new[k])


def _position(inst: dict, deps) -> int:
    for k, d in enumerate(deps):
        if d["ID"] == inst["ID"] and d["TYPE"] == inst["TYPE"]:
            return k
    raise LoweringError(f"{inst['ID']} is not among the dependencies of the FETCH it points to")


def _drop_fetches(blocks) -> None:
    """Remove every FETCH and splice its producers to its consumers (interpreter.py:768-806)."""
    doomed = []
    for bi, block in enumerate(blocks):
        for fi, fetch in enumerate(block):
            if fetch["TYPE"] != "FETCH":
                continue
            doomed.append((bi, fi))
            raw, war = fetch["Dependency"]["RAW"], fetch["Dependency"]["WAR"]
            for inst in block:
                for k, d in enumerate(inst["Dependency"]["RAW"]):
                    if d["ID"] == fetch["ID"] and d["TYPE"] == "FETCH":
                        src_k = _position(inst, war)
                        if src_k >= len(raw):
                            raise LoweringError(f"{fetch['ID']} feeds more consumers than it has producers")
                        mirror = inst["Enable"]["WAR"][k]
                        for tgt in (d, mirror):
                            tgt["ID"], tgt["TYPE"], tgt["Times"] = raw[src_k]["ID"], raw[src_k]["TYPE"], raw[src_k]["Times"]
                for k, d in enumerate(inst["Dependency"]["WAR"]):
                    if d["ID"] == fetch["ID"] and d["TYPE"] == "FETCH":
                        dst_k = _position(inst, raw)
                        if dst_k >= len(war):
                            raise LoweringError(f"{fetch['ID']} has more producers than consumers")
                        mirror = inst["Enable"]["RAW"][k]
                        for tgt in (d, mirror):
                            tgt["ID"], tgt["TYPE"] = war[dst_k]["ID"], war[dst_k]["TYPE"]
    for bi, fi in sorted(doomed, reverse=True):
        del blocks[bi][fi]


# ---- entry points --------------------------------------------------------------------------------------

def lower(op_info, op_array, tile_size_list, node_num: int, fusable=None) -> list:
    """The ISA program (list of blocks of instruction records) for a plan: ``op_array`` lists the op
    positions of every fused block, ``tile_size_list`` its ``[SR, SC]``."""
    fusable = DEFAULT_FUSABLE if fusable is None else fusable
    for pos, op in enumerate(op_info):
        if "COMP_TYPE" not in op:
            raise KeyError("COMP_TYPE")          # V1/V2-era YAML, as in the reference (interpreter.py:135)
    blocks = []
    for ops, (SR, SC) in zip(op_array, tile_size_list):
        TR, TC = math.ceil(node_num / SR), math.ceil(node_num / SC)
        block = []
        for pos in ops:
            loads, comp, store = gen_inst(op_info, pos, ops, TR, TC, SR, SC)
            block.extend(loads)
            block.append(comp)
            if store != []:
                block.append(store)
        blocks.append(block)
    _apply_fusion(blocks, _fuse_pairs(blocks, fusable))
    _drop_fetches(blocks)
    return blocks


def dump(blocks, path: str) -> None:
    """Write a program the way the reference does (``yaml.dump`` with aliases suppressed, :33-47)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        yaml.dump(blocks, f, Dumper=_NoAliasDumper)


def dumps(blocks) -> str:
    return yaml.dump(blocks, Dumper=_NoAliasDumper)


def unfused_plan(op_info, tile_rows: int = 512):
    """``(op_array, tile_size_list)`` with one op per block: the plan that always lowers (no instruction fusion, every
    tensor stored), for running the chain without the reference's ``compile()``.  ``execute()`` fuses dead stores away
    again (``fuse_across_blocks``), so on the GPU it costs nothing against a compiler-chosen plan
    (profiles/r01_plan_vs_model.json: every GAT layer-1 plan runs in the same 75 us)."""
    n = len(op_info)
    return [[i] for i in range(n)], [[int(tile_rows), 1] for _ in range(n)]


def interpret(data_set, network, isReorder, layer, op_array, tile_size_list):
    """Drop-in for the reference's ``interpret`` (same arguments, same CWD-relative files, returns None)."""
    node_num = NODE_COUNT.get(data_set, 0)
    op_map = "trans" if isReorder else "original"
    with open(f"Network/{network}/{network}-{data_set}/{network}-{op_map}/{network}-{layer}-{op_map}.yaml") as f:
        op_info = yaml.load(f, Loader=yaml.FullLoader)
    # the reference requires its fusion table next to the code; without it the shipped table's five
    # fusable patterns (hardware_info.yaml:11-68) apply
    fusable = load_fusable("code/hardware_info.yaml") if os.path.exists("code/hardware_info.yaml") else DEFAULT_FUSABLE
    blocks = lower(op_info, op_array, tile_size_list, node_num, fusable)
    dump(blocks, os.path.join("Results/Insts", f"{network}-{data_set}-{layer}-{op_map}.yaml"))
