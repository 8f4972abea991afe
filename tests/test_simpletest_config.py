"""BASELINE config 1: the reference's own ``V2/simpletest.yaml`` (9 ops, Cora shape, no COMP_TYPE field;
tests/golden/opgraph/simpletest.yaml is the reference's file, copied by oracle/gen_golden.py) runs through
lowering -> execute() once COMP_TYPE is supplied by position, on the CPU test double and on
the GPU, against the op-by-op oracle."""
import os

import numpy as np
import pytest
import yaml

from conftest import assert_close_rowscale
from oracle import gta_oracle as O
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import isa, lowering, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLANS = {"unfused": None, "three-blocks": [[0, 1, 2], [3, 4, 5, 6], [7, 8]]}


def _case():
    with open(os.path.join(GOLDEN, "opgraph", "simpletest.yaml")) as f:
        raw = yaml.safe_load(f)
    assert len(raw) == 9 and all("COMP_TYPE" not in op for op in raw)
    n, e, fin = synthetic.SHAPES["cora"]
    assert raw[0]["INPUT"]["feature_number"] == [n] and raw[5]["INPUT"]["feature_number"] == [e, e]
    assert raw[0]["INPUT"]["size_per_feature"] == [fin * 4]
    g = synthetic.shape_graph("cora")
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    x, w, al, ar = synthetic.gat_tensors(n, fin, 128, 4, seed=0)
    return raw, g, indptr, indices, {0: x}, {0: w, 1: al, 2: ar}


def _program(stamped, plan, n):
    if plan is None:
        plan, tiles = lowering.unfused_plan(stamped)
    else:
        tiles = [[512, 1]] * len(plan)
    return lowering.lower(stamped, plan, tiles, n)


def test_unstamped_file_is_refused_with_a_pointer_to_the_fix():
    raw = _case()[0]
    with pytest.raises(isa.IsaError, match="legacy_comp_types"):
        isa.validate_op_graph(raw)
    with pytest.raises(isa.IsaError):
        isa.stamp_comp_types(raw, ["MM"] * 8)
    with pytest.raises(isa.IsaError):
        isa.stamp_comp_types(raw, ["MM"] * 8 + ["XX"])
    stamped = isa.stamp_comp_types(raw, isa.LEGACY_SIMPLETEST_COMP_TYPES)
    assert [op["COMP_TYPE"] for op in stamped] == list(isa.LEGACY_SIMPLETEST_COMP_TYPES)
    assert all("COMP_TYPE" not in op for op in raw)          # the caller's list is not touched
    isa.validate_op_graph(stamped)


@pytest.mark.parametrize("plan", list(PLANS), ids=list(PLANS))
def test_simpletest_on_the_host_double(plan, monkeypatch):
    import torch
    import host_kernels
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph
    raw, g, indptr, indices, node_inputs, weights = _case()
    stamped = isa.stamp_comp_types(raw, isa.LEGACY_SIMPLETEST_COMP_TYPES)
    records = _program(stamped, PLANS[plan], g.num_nodes)
    ref, scale = O.run_opgraph(stamped, indptr, indices, node_inputs, weights, stabilize=False, return_scale=True)
    # the executor wired to the CPU test double (tests/host_kernels.py), as in tests/test_cpu_executor.py
    monkeypatch.setattr(executor, "kernels", host_kernels)
    monkeypatch.setattr(graph.DeviceGraph, "schedule", lambda self, *a, **k: None)
    dg = graph.DeviceGraph(g.num_nodes, g.num_edges, torch.from_numpy(indptr), torch.from_numpy(indices.astype(np.int32)),
                           num_sources=g.num_nodes)
    t = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
    out = executor.execute(records, raw, dg, t(node_inputs), t(weights), legacy_comp_types=isa.LEGACY_SIMPLETEST_COMP_TYPES,
                           stabilize=False)
    assert sorted(out) == [8]
    assert_close_rowscale(out[8].numpy(), ref[8], scale[8], what=plan)


@pytest.mark.gpu
@pytest.mark.parametrize("plan", list(PLANS), ids=list(PLANS))
def test_simpletest_on_the_gpu(plan):
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph
    raw, g, indptr, indices, node_inputs, weights = _case()
    stamped = isa.stamp_comp_types(raw, isa.LEGACY_SIMPLETEST_COMP_TYPES)
    records = _program(stamped, PLANS[plan], g.num_nodes)
    ref, scale = O.run_opgraph(stamped, indptr, indices, node_inputs, weights, stabilize=False, return_scale=True)
    dg = graph.csr_from_coo(g.dst, g.src, g.num_nodes)
    dev = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    out, log = executor.execute(records, raw, dg, dev(node_inputs), dev(weights), stabilize=False, return_log=True,
                                legacy_comp_types=isa.LEGACY_SIMPLETEST_COMP_TYPES)
    assert sorted(out) == [8]
    assert_close_rowscale(out[8].cpu().numpy(), ref[8], scale[8], what=f"{plan}: {log}")
    assert "gta_gemm_f32+el/er" in [k for k, _ in log] or plan == "unfused"
