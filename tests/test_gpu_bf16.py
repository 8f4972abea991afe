"""bf16 STORAGE MODE: the gathered table Z in bf16, fp32
accumulation.  Kernel level: on the SAME bf16-rounded inputs the kernels are held to the fp32 tolerance (only the
summation differs).  Layer level: against the fp64 oracle of the unrounded layer, the mode's own tolerance
rtol 2e-2, atol 1e-2 * rowscale."""
import os

import numpy as np
import pytest
import yaml

from conftest import assert_close_rowscale
from oracle import gta_oracle as O, parity as P
from gta_graph_tensor_acclelrator_for_general_gnn_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BF16_RTOL, BF16_ATOL = 2e-2, 1e-2


def _bf16_round(a: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bf16 -> fp32, in numpy."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(a.shape)


def _assert_bf16_close(y, y64, rowscale, what=""):
    y = np.asarray(y, dtype=np.float64)
    bound = BF16_RTOL * np.abs(y64) + BF16_ATOL * rowscale + 1e-30
    worst = float(np.max(np.abs(y - y64) / bound))
    assert np.all(np.isfinite(y)) and worst <= 1.0, f"{what}: {worst:.3f}x the bf16 tolerance"


@pytest.mark.parametrize("f,wkind", [(128, "scalar"), (64, "scalar"), (256, "scalar"), (512, "none"), (128, "heads")])
@pytest.mark.parametrize("chunk,col_block", [(1024, 0), (64, 700)])
def test_aggregate_bf16_kernel(f, wkind, chunk, col_block):
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph, kernels
    n, e = 3000, 90000
    g = synthetic.powerlaw_graph(n, e, seed=2, i0=3.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    dg = graph.csr_from_coo(g.dst, g.src, n)
    rng = np.random.default_rng(1)
    x = _bf16_round(rng.standard_normal((n, f), dtype=np.float32))
    w = {"scalar": rng.uniform(0.1, 1.0, size=(e, 1)), "none": None, "heads": rng.uniform(0.1, 1.0, size=(e, 4))}[wkind]
    w = None if w is None else w.astype(np.float32)
    xd = kernels.to_table(torch.from_numpy(x).cuda().to(torch.bfloat16))
    assert xd.dtype == torch.bfloat16 and torch.equal(xd.float().cpu(), torch.from_numpy(x))
    wd = None if w is None else torch.from_numpy(w).cuda()
    sched = dg.schedule(chunk, col_block)
    out = kernels.aggregate(dg, xd, wd, sched=sched)
    assert out.dtype == torch.float32 and torch.equal(out, kernels.aggregate(dg, xd, wd, sched=sched))
    want = O.spmm(indptr, indices, w, x)
    scale = O.spmm(indptr, indices, None if w is None else np.abs(w), np.abs(x))
    assert_close_rowscale(out.cpu().numpy(), want, scale, what=f"bf16 aggregate f={f} {wkind}")
    # and the fp32 kernel on the same (bf16-representable) values agrees within the same bound
    out32 = kernels.aggregate(dg, kernels.to_table(torch.from_numpy(x).cuda()), wd, sched=sched)
    assert_close_rowscale(out32.cpu().numpy(), want, scale, what="fp32 aggregate, same inputs")


@pytest.mark.parametrize("f,heads", [(128, 4), (128, 1), (64, 2), (256, 4)])
@pytest.mark.parametrize("chunk,col_block", [(1024, 0), (64, 700)])
def test_gat_bf16_kernel(f, heads, chunk, col_block):
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import graph, kernels
    n, e = 3000, 90000
    g = synthetic.powerlaw_graph(n, e, seed=2, i0=3.0)
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    dg = graph.csr_from_coo(g.dst, g.src, n)
    rng = np.random.default_rng(3)
    z = _bf16_round(rng.standard_normal((n, f), dtype=np.float32))
    el = rng.standard_normal((n, heads), dtype=np.float32)
    er = rng.standard_normal((n, heads), dtype=np.float32)
    rows = O.row_ids(indptr)
    lr = O.leaky_relu(el.astype(np.float64)[rows] + er.astype(np.float64)[indices])
    mx = O.segment_max(lr, indptr)
    p = np.exp(lr - np.where(np.isfinite(mx), mx, 0)[rows])
    alpha = p / O.segment_sum(p, indptr)[rows]
    want = O.elu(O.segment_sum(O.head_broadcast(alpha, f) * z.astype(np.float64)[indices], indptr))
    scale = O.gat_rowscale(indptr, indices, z.astype(np.float64), alpha)
    zd = kernels.to_table(torch.from_numpy(z).cuda().to(torch.bfloat16))
    eld, erd = torch.from_numpy(el).cuda(), torch.from_numpy(er).cuda()
    sched = dg.schedule(chunk, col_block)
    for bounded in (True, False):
        out = kernels.gat_aggregate(dg, eld, erd, zd, sched=sched, bounded=bounded)
        assert torch.equal(out, kernels.gat_aggregate(dg, eld, erd, zd, sched=sched, bounded=bounded))
        assert_close_rowscale(out.cpu().numpy(), want, scale, what=f"bf16 GAT f={f} H={heads} bounded={bounded}")
    with pytest.raises(Exception):          # 8 heads on a bf16 table: no kernel yet, refused loudly
        kernels.gat_aggregate(dg, torch.zeros((n, 8), device="cuda"), torch.zeros((n, 8), device="cuda"),
                              kernels.to_table(torch.zeros((n, 128), device="cuda", dtype=torch.bfloat16)))


def test_gemm_bf16_output_is_the_rounded_fp32_product():
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import kernels
    n, k, f, h = 5000, 602, 128, 4
    x, w, al, ar = synthetic.gat_tensors(n, k, f, h, seed=5)
    dev = lambda a: torch.from_numpy(a).cuda()
    z32, el32, er32 = kernels.gemm(kernels.to_table(dev(x)), dev(w), dev(al), dev(ar))
    zb, elb, erb = kernels.gemm(kernels.to_table(dev(x)), dev(w), dev(al), dev(ar), z_dtype=torch.bfloat16)
    assert zb.dtype == torch.bfloat16 and zb.stride(0) % 8 == 0
    assert torch.equal(zb, z32.to(torch.bfloat16))          # one rounding, of the same fp32 accumulator
    assert torch.equal(elb, el32) and torch.equal(erb, er32)          # el / er never see the rounding


@pytest.mark.parametrize("name,network,reorder,final", [
    ("GAT-cora-layer1-original__0-1-2_4-5-6-7-8_3-9-10-11-12-13", "GAT", False, 13),
    ("GCN-cora-layer1-trans__0_1-2-3", "GCN", True, 3),
])
def test_layer_in_bf16_storage_mode(name, network, reorder, final):
    import torch
    from gta_graph_tensor_acclelrator_for_general_gnn_b200 import executor, graph
    n, e, fin = synthetic.SHAPES["cora"]
    g = synthetic.shape_graph("cora")
    indptr, indices, _ = O.csr_build(g.dst, g.src, n)
    dg = graph.csr_from_coo(g.dst, g.src, n)
    with open(os.path.join(GOLDEN, "isa", name + ".yaml")) as fh:
        prog = yaml.safe_load(fh)
    opname = "GAT-cora-restamped-h4.yaml" if network == "GAT" else "GCN-cora-layer1-trans.yaml"
    with open(os.path.join(GOLDEN, "opgraph", opname)) as fh:
        op_info = yaml.safe_load(fh)
    x, w, al, ar = synthetic.gat_tensors(n, fin, 128, 4, seed=0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if network == "GAT":
        out, log = executor.execute(prog, op_info, dg, {0: dev(x)}, {0: dev(w), 1: dev(al), 2: dev(ar)}, network="GAT",
                                    feature_dtype=torch.bfloat16, return_log=True)
        z, zabs, el, er = P.host_tables(x, w, al, ar)
        ref = O.gat_layer(indptr, indices, x, w, al, ar)
        scale = O.segment_sum(O.head_broadcast(ref["alpha"], 128) * zabs[indices], indptr)
        want = ref["Y"]
    else:
        ew = synthetic.gcn_edge_norm(indptr, indices)
        out, log = executor.execute(prog, op_info, dg, {0: dev(x)}, {0: dev(w)}, {2: dev(ew)[:, None]}, network="GCN",
                                    is_reorder=True, feature_dtype=torch.bfloat16, return_log=True)
        z, zabs, _, _ = P.host_tables(x, w)
        want = O.spmm(indptr, indices, ew, z)
        scale = O.spmm(indptr, indices, np.abs(ew), zabs)
    assert any(k.endswith("zbf16") or k == "gta_gemm_f32+el/er" for k, _ in log), log
    _assert_bf16_close(out[final].cpu().numpy(), want, scale, what=f"{network} layer, bf16 storage")
    # the mode really stored bf16: the fp32 run differs from it
    kw = dict(network=network, is_reorder=reorder)
    ins = ({0: dev(x)}, {0: dev(w), 1: dev(al), 2: dev(ar)}) if network == "GAT" else ({0: dev(x)}, {0: dev(w)})
    ei = None if network == "GAT" else {2: dev(synthetic.gcn_edge_norm(indptr, indices))[:, None]}
    y32 = executor.execute(prog, op_info, dg, ins[0], ins[1], ei, **kw)[final]
    assert not torch.equal(y32, out[final])
