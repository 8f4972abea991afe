"""Loader / validator for the message-passing ISA programs the reference interpreter emits.

Input format = the YAML ``interpret()`` writes to ``Results/Insts/<net>-<ds>-<layer>-<map>.yaml``
(vTCAD/code/interpreter.py:809-853): a list of fused blocks, each a list of instruction
records ``{TYPE, ID, Hardware_Unit, Tile_Times, Tile_Size, Feature_Length, [Weight_Size],
Dependency{RAW,WAR}, Enable{RAW,WAR}}`` (gen_comp_inst :145-161, gen_load_inst :244-259,
gen_store_inst :281-296).  ``Enable`` mirrors ``Dependency`` and no consumer reads it
(SURVEY.md Appendix E), so it is ignored here.

The program is treated as READ-ONLY (the reference's fusion pass shares dict objects
between ``Dependency`` and ``Enable``, interpreter.py:627-632).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import yaml

COMP_TYPES = ("MM", "ADD", "MUL", "SF", "NONE")
OP_TYPES = ("applynode", "applyedge", "scatter", "gather")
LOAD_TYPES = ("LOAD_N", "LOAD_E", "LOAD_W")
STORE_TYPES = ("STORE_N", "STORE_E")

_ID_PART = re.compile(r"(\d+)_(applynode|applyedge|scatter|gather)_(\d+)")


class IsaError(ValueError):
    """Malformed or unknown instruction (the executor rejects, never crashes)."""


def _int_list(v, what, pos):
    if not isinstance(v, list) or any(isinstance(x, bool) or not isinstance(x, int) for x in v):
        raise IsaError(f"op {pos}: {what} must be a list of integers, got {v!r}")
    return v


#: COMP_TYPE by list position for the reference's V1/V2-era fixture ``V2/simpletest.yaml:14-209`` (the GAT
#: attention half: XW, el, er, two scatters, +, exp(leaky_relu), row sum, activation), which predates the
#: COMP_TYPE field; the V3+ interpreter raises KeyError on it (vTCAD/code/interpreter.py:135).  SURVEY App. A.
LEGACY_SIMPLETEST_COMP_TYPES = ("MM", "MM", "MM", "NONE", "NONE", "ADD", "SF", "ADD", "SF")


def stamp_comp_types(op_info, comp_types):
    """Copy of a V1/V2-era op graph with ``COMP_TYPE`` supplied by list position (records that already carry
    one keep it).  This is what ``changeyaml.modify_yaml`` does for the 14-op GAT (changeyaml.py:18-114),
    generalised: the caller names the arithmetic, nothing is guessed."""
    import copy
    if not isinstance(op_info, list):
        raise IsaError("an op graph is a non-empty list of op records")
    comp_types = list(comp_types)
    if len(comp_types) != len(op_info):
        raise IsaError(f"legacy_comp_types has {len(comp_types)} entries, the op graph has {len(op_info)} ops")
    out = copy.deepcopy(op_info)
    for pos, (op, ct) in enumerate(zip(out, comp_types)):
        if ct not in COMP_TYPES:
            raise IsaError(f"legacy_comp_types[{pos}] = {ct!r} is not one of {COMP_TYPES}")
        if isinstance(op, dict):
            op.setdefault("COMP_TYPE", ct)
    return out


def validate_op_graph(op_info) -> None:
    """Schema check of an op-graph list (template/op_template.yaml:1-19 plus ``COMP_TYPE``): the fields the
    executor reads exist and have the types it assumes.  Raises :class:`IsaError` naming the op."""
    if not isinstance(op_info, list) or not op_info:
        raise IsaError("an op graph is a non-empty list of op records")
    n = len(op_info)
    for pos, op in enumerate(op_info):
        if not isinstance(op, dict):
            raise IsaError(f"op {pos}: an op record is a mapping, got {type(op).__name__}")
        if "COMP_TYPE" not in op:
            raise IsaError(f"op {pos} has no COMP_TYPE (V1/V2-era YAML; re-stamp it, changeyaml.py:18-114, or pass "
                           f"legacy_comp_types= to execute(), e.g. isa.LEGACY_SIMPLETEST_COMP_TYPES)")
        if op.get("TYPE") not in OP_TYPES:
            raise IsaError(f"op {pos}: unknown TYPE {op.get('TYPE')!r}")
        if op["COMP_TYPE"] not in COMP_TYPES:
            raise IsaError(f"op {pos}: unknown COMP_TYPE {op['COMP_TYPE']!r}")
        if op.get("ORDER") not in ("R", "C"):
            raise IsaError(f"op {pos}: ORDER {op.get('ORDER')!r}")
        inp, out = op.get("INPUT"), op.get("OUTPUT")
        if not isinstance(inp, dict) or not isinstance(out, dict):
            raise IsaError(f"op {pos}: INPUT and OUTPUT must be mappings")
        for key in ("input_g_list", "size_per_feature", "feature_number"):
            if key not in inp:
                raise IsaError(f"op {pos}: INPUT.{key} is missing")
            _int_list(inp[key], f"INPUT.{key}", pos)
        for q in inp["input_g_list"]:
            if q != -1 and not 0 <= q < n:
                raise IsaError(f"op {pos}: producer {q} is not an op of this graph")
        if len(inp["size_per_feature"]) < max(len(inp["input_g_list"]), 1):
            raise IsaError(f"op {pos}: INPUT.size_per_feature has fewer entries than the op has inputs")
        if "output_list" not in out:
            raise IsaError(f"op {pos}: OUTPUT.output_list is missing")
        _int_list(out["output_list"], "OUTPUT.output_list", pos)
        width = out.get("size_per_feature")
        if isinstance(width, bool) or not isinstance(width, int) or width <= 0:
            raise IsaError(f"op {pos}: OUTPUT.size_per_feature must be a positive integer, got {width!r}")
        if any(isinstance(w, bool) or w <= 0 for w in inp["size_per_feature"]):
            raise IsaError(f"op {pos}: INPUT.size_per_feature must be positive")


@dataclass(frozen=True)
class OpRef:
    op: int        # position in the op-graph list (the reference indexes op_info by position)
    kind: str      # applynode | applyedge | scatter | gather
    slot: int      # input slot for LOAD_*, 0 for COMP / STORE


@dataclass
class Instruction:
    type: str
    id: str
    unit: str
    tile_times: int
    tile_size: int
    feature_length: int
    weight_size: int | None
    raw: list = field(default_factory=list)   # [(TYPE, ID, [a, b])]
    war: list = field(default_factory=list)
    refs: tuple = ()

    @property
    def is_load(self) -> bool:
        return self.type in LOAD_TYPES

    @property
    def is_store(self) -> bool:
        return self.type in STORE_TYPES

    @property
    def is_comp(self) -> bool:
        return self.type.startswith("COMP_")

    @property
    def comp_types(self) -> tuple:
        """('MUL', 'ADD') for COMP_MUL_COMP_ADD, ('MM',) for COMP_MM."""
        if not self.is_comp:
            return ()
        return tuple(p for p in self.type.split("_") if p != "COMP")


def parse_id(inst_id: str) -> tuple:
    if not isinstance(inst_id, str):
        raise IsaError(f"instruction ID must be a string, got {type(inst_id).__name__}")
    refs = tuple(OpRef(int(m.group(1)), m.group(2), int(m.group(3))) for m in _ID_PART.finditer(inst_id))
    rebuilt = "_".join(f"{r.op}_{r.kind}_{r.slot}" for r in refs)
    if not refs or rebuilt != inst_id:
        raise IsaError(f"unparseable instruction ID {inst_id!r}")
    return refs


def _deps(rec: dict, kind: str) -> list:
    out = []
    for d in rec["Dependency"][kind]:
        times = list(d["Times"])
        if not isinstance(d["TYPE"], str) or not isinstance(d["ID"], str) or len(times) != 2:
            raise TypeError(f"{kind} dependency {d!r}")
        out.append((d["TYPE"], d["ID"], [int(t) for t in times]))
    return out


def _parse_instruction(rec: dict) -> Instruction:
    if not isinstance(rec, dict):
        raise IsaError(f"an instruction is a mapping, got {type(rec).__name__}")
    try:
        typ = rec["TYPE"]
        if not isinstance(typ, str):
            raise TypeError(f"TYPE {typ!r}")
        inst = Instruction(
            type=typ, id=rec["ID"], unit=rec.get("Hardware_Unit", ""),
            tile_times=int(rec["Tile_Times"]), tile_size=int(rec["Tile_Size"]),
            feature_length=int(rec["Feature_Length"]),
            weight_size=rec.get("Weight_Size"),
            raw=_deps(rec, "RAW"), war=_deps(rec, "WAR"),
        )
    except (KeyError, TypeError, ValueError) as exc:
        raise IsaError(f"malformed instruction record: {exc!r}") from exc
    inst.refs = parse_id(inst.id)
    if typ == "FETCH":
        raise IsaError("FETCH must have been removed by fuse_fetch (interpreter.py:768-806)")
    if inst.is_load or inst.is_store:
        if len(inst.refs) != 1:
            raise IsaError(f"{typ} {inst.id}: load/store names exactly one op")
    elif inst.is_comp:
        ct = inst.comp_types
        if len(ct) != len(inst.refs) or any(c not in COMP_TYPES for c in ct):
            raise IsaError(f"unknown compute instruction {typ} {inst.id}")
    else:
        raise IsaError(f"unknown instruction TYPE {typ!r}")
    return inst


@dataclass
class Program:
    blocks: list        # list[list[Instruction]]

    @classmethod
    def from_records(cls, records) -> "Program":
        if not isinstance(records, list) or any(not isinstance(b, list) for b in records):
            raise IsaError("an ISA program is a list of blocks, each a list of instructions")
        return cls([[_parse_instruction(r) for r in block] for block in records])

    @classmethod
    def load(cls, path: str) -> "Program":
        with open(path) as f:
            return cls.from_records(yaml.safe_load(f))

    def block_ops(self, op_info) -> list:
        """Op positions of every block, scatters included.

        Compute ops are named by their COMP instruction.  A scatter leaves no COMP after
        fuse_fetch; it is attributed to the block that loads its input or stores its output,
        else (producer and consumers all inside one block) to its producer's block."""
        owner = {}
        for b, block in enumerate(self.blocks):
            for inst in block:
                for r in inst.refs:
                    if inst.is_comp or r.kind == "scatter":
                        owner.setdefault(r.op, b)
        for pos, op in enumerate(op_info):
            if pos in owner or op["TYPE"] != "scatter":
                continue
            prods = [p for p in op["INPUT"]["input_g_list"] if p != -1]
            if prods and prods[0] in owner:
                owner[pos] = owner[prods[0]]
        beyond = sorted(p for p in owner if p >= len(op_info))
        if beyond:
            raise IsaError(f"the program names ops {beyond}, the op graph has {len(op_info)}")
        missing = [p for p in range(len(op_info)) if p not in owner]
        if missing:
            raise IsaError(f"ops {missing} of the op graph appear in no block of the program")
        out = [[] for _ in self.blocks]
        for pos in sorted(owner):
            out[owner[pos]].append(pos)
        return out

    def stored_ops(self) -> list:
        """Per block: op positions whose output a STORE_* materialises."""
        return [[inst.refs[0].op for inst in block if inst.is_store] for block in self.blocks]

    def summary(self) -> list:
        return [[(i.type, i.id, i.tile_times, i.tile_size, i.feature_length) for i in b] for b in self.blocks]
