"""Op-graph generator: the host-side mirror of the reference's ``gen_yaml``.

The reference describes one GNN layer as a YAML list of op records (SURVEY.md §8 a1/a2;
``vTCAD/GraphOP/genGraphOP.py:4-25`` builds one record, ``:27-154`` lists the records of each
network).  This module produces the same files from a compact table: an op is

    (kind.order, computation, producers, consumers, input widths, output width)

with widths as symbols -- ``I`` the layer's input width, ``O`` its output width, ``A`` the width of
the attention projection -- and everything else DERIVED: which of N/E counts its rows, the declared
input count, and the weight bytes of an ``MM`` (= input width x output width x 4).  Files written by
``gen_yaml`` are byte-identical to the reference's for every (network, layer, reorder) it knows
(tests/test_cpu_lowering.py compares them with the fixtures the unmodified reference wrote).

The reference's published quirks are reproduced on purpose so that downstream tools see the same
input: GAT's second projection carries ``OP_NO`` 1, and the reordered GCN is
emitted with consumer lists shifted by one and a one-entry ``feature_number`` on a two-input op.
``repair=True`` applies the data fix of Appendix C-4 to that last case (this is what the golden
fixtures and the executor tests use); the default keeps the file as the reference writes it.
"""
from __future__ import annotations

import os

import yaml

# widths per layer, in fp32 elements (reference: genGraphOP.py:31-32)
_LAYER_OUT = {1: 128, 2: 64, 3: 16}
_ATTN = 16

_EDGE_ROWS = {"applyedge", "gather"}   # ops whose INPUT rows are edges
_EDGE_OUT = {"scatter", "applyedge"}   # ops whose OUTPUT rows are edges


class OpGraphError(ValueError):
    pass


def _mp(src_width):
    """scatter(C) -> x edge weight -> gather: the 3-op neighbourhood sum used by four networks"""
    return lambda first, src: [
        ("scatter.C", "NONE", src, [first + 1], [src_width], src_width),
        ("applyedge.R", "MUL", [first, -1], [first + 2], [src_width, src_width], src_width),
        ("gather.R", "ADD", [first + 1], [first + 3], [src_width], src_width),
    ]


def _gat_front():
    return [
        ("applynode.R", "MM", [], [1, 2, 3], ["I"], "O"),
        ("applynode.R", "MM", [0], [4], ["O"], "A"),
        ("applynode.R", "MM", [0], [5], ["O"], "A", {"op_no": 1}),
        ("scatter.C", "NONE", [0], [11], ["O"], "O"),
        ("scatter.R", "NONE", [1], [6], ["A"], "A"),
        ("scatter.C", "NONE", [2], [6], ["A"], "A"),
        ("applyedge.R", "ADD", [4, 5], [7], ["A", "A"], "A"),
    ]


def _tail_pna():
    return [
        ("applyedge.R", "ADD", [3, 4], [6], ["O", "O"], "O"),
        ("applyedge.R", "ADD", [2, 5], [7], ["O", "O"], "O"),
        ("applyedge.R", "SF", [6], [8], ["O"], "O"),
        ("gather.R", "ADD", [7], [9], ["O"], "O"),
        ("applynode.R", "MUL", [8], [10], ["O"], "O"),
        ("applynode.R", "MM", [9], [], ["O"], "O"),
    ]


def _table(network, reorder):
    mp = _mp("I")
    if network == "GCN" and not reorder:
        return mp(0, []) + [("applynode.R", "MM", [2], [], ["I"], "O")]
    if network == "GCN":
        return [
            ("applynode.R", "MM", [], [1], ["I"], "O"),
            ("scatter.C", "NONE", [0], [1], ["O"], "O"),
            ("applyedge.R", "MUL", [1, -1], [2], ["O", "O"], "O", {"rows": 1}),
            ("gather.R", "ADD", [2], [], ["O"], "O"),
        ]
    if network == "GAT" and not reorder:
        return _gat_front() + [
            ("applyedge.R", "SF", [6], [8, 9], ["A"], "A"),
            ("gather.R", "ADD", [7], [10], ["A"], "A"),
            ("applyedge.R", "MUL", [7, 10], [11], ["A", "A"], "A"),
            ("scatter.R", "NONE", [7], [9], ["A"], "A"),
            ("applyedge.R", "MUL", [3, 9], [12], ["O", "A"], "O"),
            ("gather.R", "ADD", [11], [13], ["O"], "O"),
            ("applynode.R", "SF", [12], [], ["O"], "O"),
        ]
    if network == "GAT":
        return _gat_front() + [
            ("applyedge.R", "MUL", [3, 8], [10], ["O", "A"], "O"),
            ("applyedge.R", "SF", [6], [9], ["A"], "A"),
            ("gather.R", "ADD", [8], [11], ["A"], "A"),
            ("gather.R", "ADD", [7], [11], ["O"], "O"),
            ("applynode.R", "MUL", [9, 10], [12], ["A", "O"], "O"),
            ("applynode.R", "SF", [11], [], ["O"], "O"),
        ]
    if network == "SGC":
        return mp(0, []) + mp(3, [2]) + [("applynode.R", "MM", [5], [], ["I"], "O")]
    if network == "GraphSAGE":
        return mp(0, []) + [
            ("applynode.R", "MM", [2], [5], ["I"], "O"),
            ("applynode.R", "MM", [], [5], ["I"], "O"),
            ("applynode.R", "ADD", [3, 4], [6], ["O", "O"], "O"),
            ("applynode.R", "SF", [5], [], ["O"], "O"),
        ]
    if network == "GIN":
        return mp(0, []) + [
            ("applynode.R", "MUL", [-1, -1], [4], ["I", 1], "I"),
            ("applynode.R", "ADD", [2, 3], [5], ["I", "I"], "I"),
            ("applynode.R", "MM", [4], [6], ["I"], "O"),
            ("applynode.R", "SF", [5], [7], ["O"], "O"),
            ("applynode.R", "MM", [6], [8], ["O"], "O"),
            ("applynode.R", "SF", [7], [], ["O"], "O"),
        ]
    if network == "DGN":
        return [
            ("scatter.C", "NONE", [], [2], ["I"], "I"),
            ("scatter.R", "NONE", [], [2], ["I"], "I"),
            ("applyedge.R", "ADD", [0, 1], [3], ["I", "I"], "I"),
            ("applyedge.R", "MM", [2], [7], ["I"], "O"),
            ("scatter.C", "NONE", [], [6], ["O"], "O"),
            ("scatter.R", "NONE", [], [6], ["O"], "O"),
            ("applyedge.R", "ADD", [4, 5], [7], ["O", "O"], "O"),
            ("applyedge.R", "ADD", [3, 6], [8], ["O", "O"], "O"),
            ("gather.R", "ADD", [7], [9], ["O"], "O"),
            ("applynode.R", "MUL", [8], [10], ["O"], "O"),
            ("applynode.R", "SF", [9], [], ["O"], "O"),
        ]
    if network == "PNA" and not reorder:
        return [
            ("scatter.C", "NONE", [], [3], ["I"], "I"),
            ("scatter.R", "NONE", [], [4], ["I"], "I"),
            ("applyedge.R", "MM", [], [6], ["I"], "O"),
            ("applyedge.R", "MM", [0], [5], ["I"], "O"),
            ("applyedge.R", "MM", [1], [5], ["I"], "O"),
        ] + _tail_pna()
    if network == "PNA":
        return [
            ("applynode.R", "MM", [0], [3], ["I"], "O"),
            ("applynode.R", "MM", [1], [4], ["I"], "O"),
            ("applyedge.R", "MM", [], [6], ["I"], "O"),
            ("scatter.C", "NONE", [], [5], ["O"], "O"),
            ("scatter.R", "NONE", [], [5], ["O"], "O"),
        ] + _tail_pna()
    raise OpGraphError(f"no such network: {network!r}")


NETWORKS = ("GCN", "GAT", "SGC", "GraphSAGE", "GIN", "DGN", "PNA")


def gen_one_op(op_no, comp_type, type, order, feature_number, input_g_list, input_g_num,
               input_nong_num, input_nong_list, input_size, input_size_per_feature,
               output_list, output_number, output_size_per_feature):
    """One op record, same positional signature as the reference (genGraphOP.py:4)."""
    return {
        "OP_NO": op_no, "COMP_TYPE": comp_type, "TYPE": type, "ORDER": order,
        "INPUT": {
            "input_g_list": list(input_g_list), "input_g_num": input_g_num,
            "input_nong_num": input_nong_num, "input_nong_list": list(input_nong_list),
            "input_size": list(input_size), "feature_number": list(feature_number),
            "size_per_feature": list(input_size_per_feature),
        },
        "OUTPUT": {
            "output_list": list(output_list), "output_number": output_number,
            "size_per_feature": output_size_per_feature,
        },
    }


def build(node_num, edge_num, size_per_feature, network, layer, isReorder, repair=False, attn=_ATTN):
    """The op records of one layer as a list of dicts (what ``gen_yaml`` dumps).  ``attn`` is the
    width of GAT's attention projection: 16 in ``gen_yaml``, 4 in the re-stamped ``GAT_Cora.yaml``."""
    if layer not in _LAYER_OUT:
        raise OpGraphError(f"layer must be 1, 2 or 3, got {layer!r}")
    widths = {"I": [size_per_feature, 128, 64][layer - 1], "O": _LAYER_OUT[layer], "A": attn}

    def w(sym):
        return int(widths.get(sym, sym))

    data = []
    for pos, spec in enumerate(_table(network, bool(isReorder))):
        kind_order, comp, producers, consumers, in_w, out_w = spec[:6]
        quirk = spec[6] if len(spec) > 6 else {}
        kind, order = kind_order.split(".")
        rows_in = edge_num if kind in _EDGE_ROWS else node_num
        rows_out = edge_num if kind in _EDGE_OUT else node_num
        is_mm = comp == "MM"
        data.append(gen_one_op(
            quirk.get("op_no", pos), comp, kind, order,
            [rows_in] * quirk.get("rows", len(in_w)),
            producers, len(in_w), int(is_mm), [],
            [w(in_w[0]) * w(out_w) * 4] if is_mm else [],
            [w(s) * 4 for s in in_w], consumers, rows_out, w(out_w) * 4))
    if repair and network == "GCN" and isReorder:
        repair_gcn_trans(data)
    return data


def repair_gcn_trans(data):
    """consumer lists off by one, one-entry feature_number on op 2."""
    for rec, consumers in zip(data, ([1], [2], [3], [])):
        rec["OUTPUT"]["output_list"] = consumers
    rows = data[2]["INPUT"]["feature_number"][0]
    data[2]["INPUT"]["feature_number"] = [rows, rows]
    return data


def dumps(data):
    return yaml.safe_dump(data)


def gen_yaml(path, node_num, edge_num, size_per_feature, network, layer, isReorder, repair=False):
    """Drop-in for the reference's ``gen_yaml(path, N, E, F, network, layer, isReorder)``
    (genGraphOP.py:27): writes the layer's op graph to ``path``, creating directories."""
    data = build(node_num, edge_num, size_per_feature, network, layer, isReorder, repair=repair)
    folder = os.path.dirname(path)
    if folder:
        os.makedirs(folder, exist_ok=True)
    with open(path, "w") as f:
        f.write(dumps(data))
    return data


def modify_yaml(path, node_num, edge_num, size_per_feature, isSharedBuffer=None):
    """Drop-in for the reference's re-stamper (``FinalVersion For Paper/changeyaml.py:3-201``):
    rewrites, in place, the sizes of the shipped 14-op GAT file (``V2/GAT_Cora.yaml``, which carries
    CiteSeer shapes and no ``COMP_TYPE``) to the graph ``(N, E, F)`` -- layer 1, attention width 4 --
    and adds ``COMP_TYPE``.  Wiring (``OP_NO``, ``TYPE``, ``ORDER``, producer/consumer lists) is kept
    from the file.  ``isSharedBuffer`` is accepted and ignored, as in the reference."""
    with open(path) as f:
        data = yaml.safe_load(f)
    stamp = build(node_num, edge_num, size_per_feature, "GAT", 1, False, attn=4)
    # the re-stamper lists op 11's operand widths as (alpha, Z); gen_yaml lists them as (Z, alpha)
    stamp[11]["INPUT"]["size_per_feature"].reverse()
    if not isinstance(data, list) or len(data) < len(stamp):
        raise OpGraphError(f"{path}: expected the {len(stamp)}-op GAT layer, found "
                           f"{len(data) if isinstance(data, list) else type(data).__name__}")
    for rec, new in zip(data, stamp):
        for key in ("size_per_feature", "feature_number", "input_size"):
            rec["INPUT"][key] = new["INPUT"][key]
        for key in ("size_per_feature", "output_number"):
            rec["OUTPUT"][key] = new["OUTPUT"][key]
        rec["COMP_TYPE"] = new["COMP_TYPE"]
    with open(path, "w") as f:
        f.write(dumps(data))
    return data


def generate_connections(yaml_file):
    """[producer OP_NO, consumer] pairs of an op-graph file (genGraphOP.py:156-169)."""
    with open(yaml_file) as f:
        data = yaml.safe_load(f)
    return [[rec["OP_NO"], c] for rec in data for c in rec["OUTPUT"]["output_list"]]


def network_path(network, dataset, layer, isReorder, root="Network"):
    """The reference's file layout: Network/<net>/<net>-<ds>/<net>-<mode>/<net>-layer<k>-<mode>.yaml"""
    mode = "trans" if isReorder else "original"
    return os.path.join(root, network, f"{network}-{dataset}", f"{network}-{mode}",
                        f"{network}-layer{layer}-{mode}.yaml")
