"""lowering.py (host mirror of the reference's interpret()/gen_inst()) against the ISA programs the
UNMODIFIED reference emitted for the same op graphs and plans (tests/golden/isa): byte-identical YAML."""
import json
import os

import pytest
import yaml

from gta_graph_tensor_acclelrator_for_general_gnn_b200 import isa, lowering

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "manifest.json")) as _f:
    MANIFEST = json.load(_f)


def _load(rel):
    with open(os.path.join(GOLDEN, rel)) as f:
        return yaml.safe_load(f)


@pytest.mark.parametrize("prog", MANIFEST["programs"], ids=[p["file"].split("/")[-1][:-5] for p in MANIFEST["programs"]])
def test_program_is_byte_identical_to_the_reference(prog):
    op_info = _load(prog["opgraph"])
    blocks = lowering.lower(op_info, prog["op_array"], prog["tile_size_list"], lowering.NODE_COUNT[prog["dataset"]])
    want = open(os.path.join(GOLDEN, prog["file"])).read()
    assert lowering.dumps(blocks) == want
    # and it parses as a program whose blocks are the plan
    program = isa.Program.from_records(blocks)
    assert [sorted(b) for b in program.block_ops(op_info)] == [sorted(b) for b in prog["op_array"]]


def test_interpret_drop_in_writes_the_same_file(tmp_path, monkeypatch):
    """Same positional arguments and CWD-relative files as the reference's interpret()."""
    prog = next(p for p in MANIFEST["programs"] if p["file"].endswith("GCN-flickr-layer1-trans__0_1-2-3.yaml"))
    net = tmp_path / "Network" / "GCN" / "GCN-flickr" / "GCN-trans"
    net.mkdir(parents=True)
    (net / "GCN-layer1-trans.yaml").write_text(open(os.path.join(GOLDEN, prog["opgraph"])).read())
    (tmp_path / "code").mkdir()
    table = [{"Hardware": {"Buffer_Size": 2}},
             {"Inst_fused": [{"Pattern": list(p), "Compute_Type": list(c), "Buffer_Type": "Edge", "Is_Fused": True}
                             for p, c in sorted(lowering.DEFAULT_FUSABLE)]
              + [{"Pattern": ["gather", "scatter"], "Compute_Type": ["ADD", "NONE"], "Buffer_Type": "Node", "Is_Fused": False}]}]
    (tmp_path / "code" / "hardware_info.yaml").write_text(yaml.safe_dump(table))
    monkeypatch.chdir(tmp_path)
    assert lowering.interpret("flickr", "GCN", True, "layer1", prog["op_array"], prog["tile_size_list"]) is None
    got = (tmp_path / "Results" / "Insts" / "GCN-flickr-layer1-trans.yaml").read_text()
    assert got == open(os.path.join(GOLDEN, prog["file"])).read()
    assert lowering.load_fusable("code/hardware_info.yaml") == lowering.DEFAULT_FUSABLE


def test_gen_inst_shape_of_the_fused_spmm_core():
    """SURVEY section 8a example: GCN-cora-original block [1,2,3], tile 48."""
    op_info = _load("opgraph/GCN-cora-layer1-original.yaml")
    loads, comp, store = lowering.gen_inst(op_info, 2, [1, 2, 3], 57, 2708, 48, 1)
    assert [(l["TYPE"], l["ID"], l["Hardware_Unit"], l["Tile_Times"]) for l in loads] == \
        [("LOAD_N", "2_gather_1", "Virtual_Loader", 57)]
    assert comp["TYPE"] == "COMP_ADD" and comp["Tile_Times"] == 57 * 2708 and comp["Tile_Size"] == 48
    assert comp["Dependency"]["RAW"][1] == {"TYPE": "COMP_MUL", "ID": "1_applyedge_0", "Times": [1, 1]}
    assert comp["Dependency"]["WAR"] == [{"TYPE": "COMP_MM", "ID": "3_applynode_0", "Times": [2708, 1]}]
    assert store == []


def test_plans_the_reference_cannot_lower_are_rejected():
    """compile()'s best GAT-trans plan crashes the reference's interpret() (gather -> scatter inside a
    block has no token rate, interpreter.py:165-194,203); here it is a LoweringError, not a crash."""
    gat_t = next(c for c in MANIFEST["compile"] if c["network"] == "GAT" and c["reorder"])
    best = gat_t["plans"][0]
    op_info = _load("opgraph/GAT-cora-layer1-trans.yaml")
    with pytest.raises(lowering.LoweringError):
        lowering.lower(op_info, best["fused_array"], best["tile_sizes"], 2708)
    legacy = [dict(op) for op in op_info]
    del legacy[0]["COMP_TYPE"]
    with pytest.raises(KeyError, match="COMP_TYPE"):
        lowering.lower(legacy, [[0]], [[16, 1]], 2708)


# ---- opgraph.py: gen_yaml / modify_yaml mirrors ------------------------------------------------
import re
import shutil

from gta_graph_tensor_acclelrator_for_general_gnn_b200 import opgraph, synthetic

_OPGRAPHS = sorted(f for f in os.listdir(os.path.join(GOLDEN, "opgraph")) if "-layer" in f)


@pytest.mark.parametrize("name", _OPGRAPHS, ids=[f[:-5] for f in _OPGRAPHS])
def test_gen_yaml_is_byte_identical_to_the_reference(name, tmp_path):
    net, ds, layer, mode = re.match(r"(\w+)-(\w+)-layer(\d)-(original|trans)\.yaml", name).groups()
    n, e, f = synthetic.SHAPES[ds]
    path = tmp_path / opgraph.network_path(net, ds, int(layer), mode == "trans")
    opgraph.gen_yaml(str(path), n, e, f, net, int(layer), mode == "trans", repair=True)
    assert path.read_text() == open(os.path.join(GOLDEN, "opgraph", name)).read()


def test_gen_yaml_keeps_the_published_gcn_trans_form_by_default():
    """Without repair=True the reordered GCN is what the reference writes:
    consumers shifted by one and a 1-entry feature_number on the 2-input edge op."""
    raw = opgraph.build(2708, 10556, 1433, "GCN", 1, True)
    assert [r["OUTPUT"]["output_list"] for r in raw] == [[1], [1], [2], []]
    assert raw[2]["INPUT"]["feature_number"] == [10556] and raw[2]["INPUT"]["input_g_num"] == 2
    fixed = opgraph.repair_gcn_trans(opgraph.build(2708, 10556, 1433, "GCN", 1, True))
    assert fixed == _load("opgraph/GCN-cora-layer1-trans.yaml")


def test_gen_yaml_rejects_unknown_network_and_layer():
    with pytest.raises(opgraph.OpGraphError):
        opgraph.build(10, 20, 8, "GraphConv", 1, False)
    with pytest.raises(opgraph.OpGraphError):
        opgraph.build(10, 20, 8, "GCN", 4, False)


@pytest.mark.parametrize("src,dst", [("cora", "reddit"), ("reddit", "cora")])
def test_modify_yaml_restamps_like_the_reference(src, dst, tmp_path):
    """The re-stamper overwrites every size field, so re-stamping one golden file (made by the
    unmodified reference from V2/GAT_Cora.yaml) to the other's shape must reproduce the other."""
    p = tmp_path / "GAT_Cora.yaml"
    shutil.copy(os.path.join(GOLDEN, "opgraph", f"GAT-{src}-restamped-h4.yaml"), p)
    opgraph.modify_yaml(str(p), *synthetic.SHAPES[dst], [])
    assert p.read_text() == open(os.path.join(GOLDEN, "opgraph", f"GAT-{dst}-restamped-h4.yaml")).read()


def test_modify_yaml_rejects_a_short_file(tmp_path):
    p = tmp_path / "short.yaml"
    p.write_text(opgraph.dumps(opgraph.build(10, 20, 8, "GCN", 1, False)))
    with pytest.raises(opgraph.OpGraphError):
        opgraph.modify_yaml(str(p), 10, 20, 8, [])


def test_generate_connections_lists_op_edges(tmp_path):
    p = tmp_path / "g.yaml"
    opgraph.gen_yaml(str(p), 10, 20, 8, "GCN", 1, False)
    assert opgraph.generate_connections(str(p)) == [[0, 1], [1, 2], [2, 3]]


def test_generated_graph_lowers_to_the_golden_program():
    """gen_yaml -> lower, no reference file on the way: still the program the reference emitted."""
    prog = next(p for p in MANIFEST["programs"] if p["file"].endswith("GCN-flickr-layer1-trans__0_1-2-3.yaml"))
    n, e, f = synthetic.SHAPES["flickr"]
    op_info = opgraph.build(n, e, f, "GCN", 1, True, repair=True)
    blocks = lowering.lower(op_info, prog["op_array"], prog["tile_size_list"], n)
    assert lowering.dumps(blocks) == open(os.path.join(GOLDEN, prog["file"])).read()


@pytest.mark.parametrize("network", opgraph.NETWORKS)
@pytest.mark.parametrize("reorder", [False, True])
def test_unfused_plan_lowers_for_every_network(network, reorder):
    op_info = opgraph.build(*synthetic.SHAPES["cora"], network, 1, reorder, repair=True)
    plan, tiles = lowering.unfused_plan(op_info)
    program = isa.Program.from_records(lowering.lower(op_info, plan, tiles, synthetic.SHAPES["cora"][0]))
    assert program.block_ops(op_info) == plan
