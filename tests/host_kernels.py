"""TEST DOUBLE for ``gta_graph_tensor_acclelrator_for_general_gnn_b200.kernels`` -- never shipped, never timed.

The executor's host logic (lazy values, pattern matching onto fused kernels, block ordering, dead
stores, refusals) is plain Python and deserves coverage in the CPU suite, where no kernel can run.
This module restates what every wrapper in ``kernels.py`` is documented to compute, with torch CPU
ops in fp32, so ``tests/test_cpu_executor.py`` can run whole ISA programs through ``execute()`` and
compare them with the oracle.  It says nothing about the CUDA kernels: those are checked against the
oracle on the GPU (tests/test_gpu_*.py).  The product never imports this file.
"""
import torch

from gta_graph_tensor_acclelrator_for_general_gnn_b200 import _cabi

LEAKY_SLOPE = 0.2
EVENT_LOG = None


def pad4(n):
    return (n + 3) // 4 * 4


def alloc_table(rows, width, device, zero=False):
    return (torch.zeros if zero else torch.empty)((rows, width), dtype=torch.float32, device=device)


def to_table(x):
    if x.dtype != torch.float32:
        raise TypeError("fp32 tables only")
    return x if x.dim() == 2 else x[:, None]


def _rows(g):
    counts = (g.indptr[1:] - g.indptr[:-1]).long()
    return torch.repeat_interleave(torch.arange(g.num_rows), counts)


def _spread(t, width):
    """head broadcast: column c of the result reads column c // (width / w) of t"""
    t = t if t.dim() == 2 else t[:, None]
    if width % t.shape[1]:
        raise ValueError(f"operand width {t.shape[1]} must divide {width}")
    return t.repeat_interleave(width // t.shape[1], dim=1)


def _segment_sum(e, g):
    out = torch.zeros((g.num_rows, e.shape[1]), dtype=torch.float32)
    out.index_add_(0, _rows(g), e)
    return out


def _segment_max(e, g):
    out = torch.full((g.num_rows, e.shape[1]), float("-inf"))
    return out.scatter_reduce(0, _rows(g)[:, None].expand_as(e), e, reduce="amax", include_self=True)


def _epilogue(x, code):
    if code == _cabi.EPI_ELU:
        return torch.nn.functional.elu(x)
    if code == _cabi.EPI_RELU:
        return torch.relu(x)
    return x


def _leaky(x, slope):
    return torch.where(x > 0, x, x * slope)


def gemm(x, w, al=None, ar=None, out=None, er_out=None, z_dtype=torch.float32):
    z = x @ w
    if out is not None:
        out.copy_(z)
        z = out
    if al is None and ar is None:
        return z
    el = z @ al if al is not None else None
    er = z @ ar if ar is not None else None
    if er_out is not None and er is not None:
        er_out.copy_(er)
        er = er_out
    return z, el, er


def aggregate(g, x, w=None, rowden=None, epilogue=_cabi.EPI_NONE, sched=None, out=None, block_events=None, exchange=None):
    src = g.indices.long()
    e = x[src]
    if w is not None:
        w = w if w.dim() == 2 else w[:, None]
        if rowden is not None:
            w = w / rowden[_rows(g)]
        e = e * _spread(w, x.shape[1])
    res = _epilogue(_segment_sum(e, g), epilogue)
    if out is not None:
        out.copy_(res)
        return out
    return res


def gat_logits(g, el, er, slope=LEAKY_SLOPE, stabilize=True):
    rows = _rows(g)
    s = _leaky(el[rows] + er[g.indices.long()], slope)
    rowmax = _segment_max(s, g)
    if stabilize:
        s = s - torch.where(torch.isfinite(rowmax), rowmax, torch.zeros_like(rowmax))[rows]
    p = torch.exp(s)
    return p, rowmax, _segment_sum(p, g)


def aggregate_edge_sum(g, edge=None, x=None, rowterm=None, unary=_cabi.UN_COPY, slope=LEAKY_SLOPE,
                       epilogue=_cabi.EPI_NONE, sched=None):
    given = [t for t in (edge, x, rowterm) if t is not None]
    f = int(given[0].shape[1])
    if any(int(t.shape[1]) != f or t.dtype != torch.float32 for t in given):
        raise ValueError("aggregate_edge_sum: fp32 operands of one width")
    v = torch.zeros((g.num_edges, f), dtype=torch.float32)
    if edge is not None:
        v = v + edge
    if x is not None:
        v = v + x[g.indices.long()]
    if rowterm is not None:
        v = v + rowterm[_rows(g)]
    return _epilogue(_segment_sum(_unary(unary, v, slope), g), epilogue)


def gat_aggregate(g, el, er, z, slope=LEAKY_SLOPE, epilogue=_cabi.EPI_ELU, sched=None, out=None,
                  want_stats=False, block_events=None, bounded=True, exchange=None):
    # the shapes gta_gat_aggregate_f32 takes (csrc/gat_aggregate.cu): the double must refuse what the library refuses, or
    # the CPU fuzz cannot see an executor that routes an unsupported shape to the fused kernel
    f, heads = int(z.shape[1]), int(el.shape[1])
    if f % 4 or f % heads or not ((f // heads) % 4 == 0 or f // heads in (1, 2)):
        raise _cabi.GtaError(f"gta_gat_aggregate_f32 (test double): f={f} heads={heads} has no kernel")
    p, rowmax, rowsum = gat_logits(g, el, er, slope, True)
    alpha = p / rowsum[_rows(g)]
    res = _epilogue(_segment_sum(z[g.indices.long()] * _spread(alpha, z.shape[1]), g), epilogue)
    if out is not None:
        out.copy_(res)
        res = out
    return (res, rowmax, rowsum) if want_stats else res


def _edge_operand(g, t, kind):
    t = t if t.dim() == 2 else t[:, None]
    if kind == _cabi.OPND_DST:
        return t[_rows(g)]
    if kind == _cabi.OPND_SRC:
        return t[g.indices.long()]
    return t


def _binary(op, a, b):
    wo = max(a.shape[1], b.shape[1])
    a, b = _spread(a, wo), _spread(b, wo)
    if op == _cabi.BIN_ADD:
        return a + b
    if op == _cabi.BIN_MUL:
        return a * b
    return torch.where(b != 0, a / torch.where(b != 0, b, torch.ones_like(b)), torch.zeros_like(a))


def _unary(op, a, slope):
    if op == _cabi.UN_EXP_LEAKY_RELU:
        return torch.exp(_leaky(a, slope))
    if op == _cabi.UN_ELU:
        return torch.nn.functional.elu(a)
    if op == _cabi.UN_RELU:
        return torch.relu(a)
    return a.clone()


def edge_binary(g, op, a, kind_a, b, kind_b):
    return _binary(op, _edge_operand(g, a, kind_a), _edge_operand(g, b, kind_b))


def edge_unary(g, op, a, kind_a, slope=LEAKY_SLOPE):
    return _unary(op, _edge_operand(g, a, kind_a), slope)


def node_binary(op, a, b):
    return _binary(op, to_table(a), to_table(b))


def node_unary(op, a, slope=LEAKY_SLOPE):
    return _unary(op, to_table(a), slope)
