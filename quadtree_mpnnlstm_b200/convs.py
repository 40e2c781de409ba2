"""Graph convolutions with PyG 2.2.0's parameter names and semantics, on the qmp_b200 kernels.

The reference takes these classes from ``torch_geometric.nn`` (model/model.py:10, 39-57).  State-dict
keys are kept (``lin.weight``/``bias``; ``lins.<k>.weight``/``bias``;
``lin_{key,query,value,skip}.{weight,bias}``/``lin_edge.weight``) so checkpoints are interchangeable
(SURVEY.md section 8b).  ``forward(x, edge_index, edge_attr_or_weight)`` keeps PyG's positional order.

Parameter packing: the kernels consume a few packed / folded weight tensors per group of convs
(see ``pack_tconv``).  Packing is a handful of tiny tensor ops under autograd, done once per forward
pass of the driver (not per timestep) and cached for its duration.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .graph_csr import get_csr
from .ops import NodeLinearFn, SpmmFn, TConvFn, next_seed


class Linear(nn.Module):
    """Parameter holder matching ``torch_geometric.nn.dense.linear.Linear`` (weight [out, in], optional
    bias; default init = torch's nn.Linear, 'glorot' on request)."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels, self.weight_initializer = in_channels, out_channels, weight_initializer
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            a = math.sqrt(6.0 / (self.weight.size(-2) + self.weight.size(-1)))
            self.weight.data.uniform_(-a, a)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)


def _n_nodes(x):
    return x.shape[0]


# ----------------------------------------------------------------------------- TransformerConv
def pack_tconv(convs):
    """Fold the PyG parameters of G TransformerConvs (same in/out sizes) into the kernel layout.

    logit_ij * sqrt(C) = q_i . (Wk x_j + bk + We e_ij); the bk term is constant over j and cancels in
    the softmax, so  u_i = Wk^T (Wq x_i + bq),  w_i = We^T (Wq x_i + bq):
        [W1 | b1] = [Wk | We]^T [Wq | bq] / sqrt(C)          -> [G, D+2, D+1]
        W2 = [Wv | We | bv]                                   -> [G, C, D+3]
        W3 = Ws, b3 = bs
    """
    C = convs[0].out_channels
    st = torch.stack
    Wq, bq = st([c.lin_query.weight for c in convs]), st([c.lin_query.bias for c in convs])
    Wk, bk = st([c.lin_key.weight for c in convs]), st([c.lin_key.bias for c in convs])
    Wv, bv = st([c.lin_value.weight for c in convs]), st([c.lin_value.bias for c in convs])
    We = st([c.lin_edge.weight for c in convs])
    Ws, bs = st([c.lin_skip.weight for c in convs]), st([c.lin_skip.bias for c in convs])
    KE = torch.cat([Wk, We], dim=2)                                   # [G, C, D+2]
    QB = torch.cat([Wq, bq.unsqueeze(-1)], dim=2)                     # [G, C, D+1]
    W1b = torch.bmm(KE.transpose(1, 2), QB) / math.sqrt(C)            # [G, D+2, D+1]
    D = Wq.shape[2]
    W1 = W1b[..., :D].contiguous()
    b1 = W1b[..., D].contiguous() + 0.0 * bk.sum()                    # lin_key.bias: exact zero gradient, not None
    W2 = torch.cat([Wv, We, bv.unsqueeze(-1)], dim=2).contiguous()
    return W1, b1, W2, Ws.contiguous(), bs.contiguous()


class TransformerConv(nn.Module):
    """``TransformerConv(in, out, heads=1, concat=False, beta=False, dropout=p, edge_dim=2, bias=True,
    root_weight=True)`` -- the configuration the reference selects (model/model.py:51)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0, edge_dim=None,
                 bias=True, root_weight=True, **kwargs):
        super().__init__()
        if heads != 1 or beta or not root_weight or edge_dim != 2 or not bias:
            raise NotImplementedError("TransformerConv: only heads=1, beta=False, root_weight=True, edge_dim=2, "
                                      "bias=True (the reference's CONVOLUTION_KWARGS) is implemented")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.dropout, self.edge_dim = concat, dropout, edge_dim
        self.lin_key = Linear(in_channels, heads * out_channels)
        self.lin_query = Linear(in_channels, heads * out_channels)
        self.lin_value = Linear(in_channels, heads * out_channels)
        self.lin_edge = Linear(edge_dim, heads * out_channels, bias=False)
        self.lin_skip = Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, edge_index, edge_attr=None):
        assert edge_attr is not None, "TransformerConv(edge_dim=2) needs edge attributes"
        csr = get_csr(edge_index, edge_attr, _n_nodes(x))
        p = self.dropout if self.training else 0.0
        return TConvFn.apply(x.float(), *pack_tconv([self]), csr, True, p, next_seed() if p > 0 else 0, False, None)


# ----------------------------------------------------------------------------- GCNConv
class GCNConv(nn.Module):
    """``GCNConv(in, out, add_self_loops=...)``: out_i = sum_{j->i} norm_ij W x_j + b with
    norm = deg^-1/2[j] w deg^-1/2[i], deg over incoming weights (PyG gcn_norm)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=True, normalize=True,
                 bias=True, **kwargs):
        super().__init__()
        if improved or not normalize:
            raise NotImplementedError("GCNConv: improved=False, normalize=True only")
        self.in_channels, self.out_channels, self.add_self_loops = in_channels, out_channels, add_self_loops
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x, edge_index, edge_weight=None):
        n = _n_nodes(x)
        if self.add_self_loops:
            edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, n)
        csr = get_csr(edge_index, edge_weight, n)
        agg = SpmmFn.apply(x.float(), None, csr, "gcn", 1.0, 0.0)
        b = self.bias.unsqueeze(0) if self.bias is not None else None
        return NodeLinearFn.apply(agg, self.lin.weight.unsqueeze(0), b, True)


_loop_cache = {}


def add_remaining_self_loops(edge_index, edge_weight, n):
    """PyG add_remaining_self_loops (fill 1): existing loops keep their weight.  Graph preprocessing
    (device tensor ops, cached per edge_index) -- only the legacy MPNNLSTM uses it."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), None if edge_weight is None else edge_weight.data_ptr(), n)
    hit = _loop_cache.get(key)
    if hit is not None and hit[0] is edge_index and hit[1] is edge_weight:
        return hit[2], hit[3]
    row, col = edge_index[0], edge_index[1]
    w = edge_weight if edge_weight is not None else torch.ones(row.numel(), device=row.device)
    is_loop = row == col
    loop_w = torch.ones(n, device=row.device, dtype=w.dtype)
    loop_w[row[is_loop]] = w[is_loop]
    ar = torch.arange(n, device=row.device, dtype=row.dtype)
    keep = ~is_loop
    ei = torch.stack([torch.cat([row[keep], ar]), torch.cat([col[keep], ar])]).contiguous()
    ew = torch.cat([w[keep], loop_w]).contiguous()
    if len(_loop_cache) > 8:
        _loop_cache.clear()
    _loop_cache[key] = (edge_index, edge_weight, ei, ew)
    return ei, ew


# ----------------------------------------------------------------------------- ChebConv
def cheb_basis(x, csr, K):
    """[T0 | T1 | ... ] along columns, T1 = L^ x, Tk = 2 L^ T(k-1) - T(k-2) (PyG ChebConv, sym, lambda_max=2)."""
    ts = [x]
    if K > 1:
        ts.append(SpmmFn.apply(x, None, csr, "cheb", 1.0, 0.0))
    for _ in range(2, K):
        ts.append(SpmmFn.apply(ts[-1], ts[-2], csr, "cheb", 2.0, -1.0))
    return ts


class ChebConv(nn.Module):
    """``ChebConv(in, out, K=3, normalization='sym', bias=True)`` (model/model.py:53)."""

    def __init__(self, in_channels, out_channels, K=3, normalization="sym", bias=True, **kwargs):
        super().__init__()
        assert K > 0
        if normalization != "sym":
            raise NotImplementedError("ChebConv: normalization='sym' only (lambda_max = 2)")
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.lins = nn.ModuleList([Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
                                   for _ in range(K)])
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x, edge_index, edge_weight=None):
        csr = get_csr(edge_index, edge_weight, _n_nodes(x))
        tx = torch.cat(cheb_basis(x.float(), csr, self.K), dim=1)
        W = torch.cat([lin.weight for lin in self.lins], dim=1).unsqueeze(0)
        b = self.bias.unsqueeze(0) if self.bias is not None else None
        return NodeLinearFn.apply(tx, W, b, True)


class _HeadView:
    """One attention head of a multi-head TransformerConv seen as a single-head conv (the slices of the PyG parameters that
    belong to it), in the shape ``pack_tconv`` consumes."""

    class _Lin:
        def __init__(self, weight, bias):
            self.weight, self.bias = weight, bias

    def __init__(self, conv, h):
        C = conv.out_channels
        rows = slice(h * C, (h + 1) * C)
        self.out_channels = C
        self.lin_query = self._Lin(conv.lin_query.weight[rows], conv.lin_query.bias[rows])
        self.lin_key = self._Lin(conv.lin_key.weight[rows], conv.lin_key.bias[rows])
        self.lin_value = self._Lin(conv.lin_value.weight[rows], conv.lin_value.bias[rows])
        self.lin_edge = self._Lin(conv.lin_edge.weight[rows], None)
        self.lin_skip = self._Lin(conv.lin_skip.weight[rows], conv.lin_skip.bias[rows])


class MHTransformerConv(nn.Module):
    """The reference's multi-head variant (model/model.py:26-37, CONVOLUTION_KWARGS :52: heads=3, edge_dim=2, dropout=0.1):
    PyG ``TransformerConv(in, out, heads, concat=True)`` -- every head attends with its own query / key / value / edge
    slices, the head outputs are concatenated and the root weight ``lin_skip`` (in -> heads * out) is added -- followed by
    ``lin`` (heads * out -> out).

    A head is exactly a single-head TransformerConv of width ``out`` (the 1 / sqrt(out) logit scale is per head), and the
    rows h * out .. (h + 1) * out of ``lin_skip`` are that head's skip path, so the ``heads`` heads run as ONE grouped launch
    of the TransformerConv kernels over a shared input (attn.cu: G = heads convs), and ``lin`` is one node GEMM."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0, edge_dim=None, bias=True,
                 root_weight=True, **kwargs):
        super().__init__()
        if not concat or beta or not root_weight or edge_dim != 2 or not bias:
            raise NotImplementedError("MHTransformerConv: concat=True, beta=False, root_weight=True, edge_dim=2, bias=True "
                                      "(the reference's CONVOLUTION_KWARGS) is implemented")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.dropout, self.edge_dim = concat, dropout, edge_dim
        self.lin_key = Linear(in_channels, heads * out_channels)
        self.lin_query = Linear(in_channels, heads * out_channels)
        self.lin_value = Linear(in_channels, heads * out_channels)
        self.lin_edge = Linear(edge_dim, heads * out_channels, bias=False)
        self.lin_skip = Linear(in_channels, heads * out_channels, bias=bias)
        self.lin = Linear(out_channels * heads, out_channels)

    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=None):
        assert edge_attr is not None, "MHTransformerConv(edge_dim=2) needs edge attributes"
        csr = get_csr(edge_index, edge_attr, _n_nodes(x))
        p = self.dropout if self.training else 0.0
        heads = [_HeadView(self, h) for h in range(self.heads)]
        out = TConvFn.apply(x.float(), *pack_tconv(heads), csr, True, p, next_seed() if p > 0 else 0, False, None)   # [N, heads*out]
        return NodeLinearFn.apply(out, self.lin.weight.unsqueeze(0), self.lin.bias.unsqueeze(0), True)


# ----------------------------------------------------------------------------- GATConv / GATv2Conv
_meanloop_cache = {}


def add_self_loops_mean(edge_index, edge_attr, n):
    """PyG ``remove_self_loops`` + ``add_self_loops(fill_value='mean')`` (GATConv / GATv2Conv.forward): existing self loops go,
    every node gets one whose attributes are the mean over its incoming edges (zeros without any).  Graph preprocessing on
    the device, cached per ``edge_index`` / ``edge_attr`` pair."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_attr.data_ptr(), n)
    hit = _meanloop_cache.get(key)
    if hit is not None and hit[0] is edge_index and hit[1] is edge_attr:
        return hit[2], hit[3]
    keep = edge_index[0] != edge_index[1]
    ei, ea = edge_index[:, keep], edge_attr[keep].float()
    dev = ei.device
    cnt = torch.zeros(n, device=dev).index_add_(0, ei[1], torch.ones(ei.shape[1], device=dev))
    mean = torch.zeros(n, ea.shape[1], device=dev).index_add_(0, ei[1], ea) / cnt.clamp(min=1).unsqueeze(1)
    ar = torch.arange(n, device=dev, dtype=ei.dtype)
    ei2 = torch.cat([ei, torch.stack([ar, ar])], dim=1).contiguous()
    ea2 = torch.cat([ea, mean]).contiguous()
    if len(_meanloop_cache) > 8:
        _meanloop_cache.clear()
    _meanloop_cache[key] = (edge_index, edge_attr, ei2, ea2)
    return ei2, ea2


def _glorot(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-a, a)


class _GatBase(nn.Module):
    def _check(self, name, heads, concat, dropout, add_self_loops, edge_dim, fill_value, bias):
        if heads != 1 or not concat or dropout != 0.0 or not add_self_loops or edge_dim != 2 or fill_value != "mean" or not bias:
            raise NotImplementedError(f"{name}: heads=1, concat=True, dropout=0, add_self_loops=True, edge_dim=2, fill_value='mean', "
                                      "bias=True (the reference's CONVOLUTION_KWARGS over PyG's defaults) is implemented")


class GATConv(_GatBase):
    """PyG 2.2.0 ``GATConv(in, out, heads=1, edge_dim=2)`` (model/model.py:43, 55): x' = lin_src(x) (``lin_dst`` is the same
    module), alpha_ij = softmax_j(leaky_relu(att_src . x'_j + att_dst . x'_i + att_edge . lin_edge(e_ij), 0.2)) over the
    in-edges incl. a mean-attribute self loop, out_i = sum_j alpha_ij x'_j + bias.  Node maps on qmp_gemm, the edge phase on
    qmp_gat_fwd / qmp_gat_bwd (csrc/gat.cu)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0, add_self_loops=True,
                 edge_dim=None, fill_value="mean", bias=True, **kwargs):
        super().__init__()
        self._check("GATConv", heads, concat, dropout, add_self_loops, edge_dim, fill_value, bias)
        self.in_channels, self.out_channels, self.heads, self.negative_slope = in_channels, out_channels, heads, negative_slope
        self.lin_src = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.lin_edge = Linear(edge_dim, out_channels, bias=False, weight_initializer="glorot")
        self.att_edge = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin_src.reset_parameters()
        self.lin_dst.reset_parameters()
        self.lin_edge.reset_parameters()
        for t in (self.att_src, self.att_dst, self.att_edge):
            _glorot(t)

    def forward(self, x, edge_index, edge_attr=None):
        from .ops import GatFn
        assert edge_attr is not None, "GATConv(edge_dim=2) needs edge attributes"
        n, C = _n_nodes(x), self.out_channels
        ei, ea = add_self_loops_mean(edge_index, edge_attr, n)
        csr = get_csr(ei, ea, n)
        xs = NodeLinearFn.apply(x.float(), self.lin_src.weight.unsqueeze(0), None, True)
        att = torch.cat([self.att_src.view(1, C), self.att_dst.view(1, C)]).unsqueeze(0)            # [1, 2, C]
        a = NodeLinearFn.apply(xs, att, None, True)                                                   # [N, 2] = (alpha_src, alpha_dst)
        we = (self.lin_edge.weight * self.att_edge.view(C, 1)).sum(0)                                 # att_edge . lin_edge(e) = we . e
        return GatFn.apply(xs, a[:, 0].contiguous(), a[:, 1].contiguous(), we, csr, 1, self.negative_slope) + self.bias


class GATv2Conv(_GatBase):
    """PyG 2.2.0 ``GATv2Conv(in, out, heads=1, edge_dim=2)`` (model/model.py:44, 56; share_weights=False): x_l = lin_l(x),
    x_r = lin_r(x), alpha_ij = softmax_j(att . leaky_relu(x_l[j] + x_r[i] + lin_edge(e_ij), 0.2)) over the in-edges incl. a
    mean-attribute self loop, out_i = sum_j alpha_ij x_l[j] + bias."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0, add_self_loops=True,
                 edge_dim=None, fill_value="mean", bias=True, share_weights=False, **kwargs):
        super().__init__()
        self._check("GATv2Conv", heads, concat, dropout, add_self_loops, edge_dim, fill_value, bias)
        if share_weights:
            raise NotImplementedError("GATv2Conv: share_weights=False only")
        self.in_channels, self.out_channels, self.heads, self.negative_slope = in_channels, out_channels, heads, negative_slope
        self.lin_l = Linear(in_channels, out_channels, bias=True, weight_initializer="glorot")
        self.lin_r = Linear(in_channels, out_channels, bias=True, weight_initializer="glorot")
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.lin_edge = Linear(edge_dim, out_channels, bias=False, weight_initializer="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()
        self.lin_edge.reset_parameters()
        _glorot(self.att)

    def forward(self, x, edge_index, edge_attr=None):
        from .ops import GatFn
        assert edge_attr is not None, "GATv2Conv(edge_dim=2) needs edge attributes"
        n, C = _n_nodes(x), self.out_channels
        ei, ea = add_self_loops_mean(edge_index, edge_attr, n)
        csr = get_csr(ei, ea, n)
        W = torch.stack([self.lin_l.weight, self.lin_r.weight])                                       # [2, C, in]
        b = torch.stack([self.lin_l.bias, self.lin_r.bias])
        xlr = NodeLinearFn.apply(x.float(), W, b, True)                                               # [N, 2C] = x_l | x_r
        return GatFn.apply(xlr[:, :C].contiguous(), xlr[:, C:].contiguous(), self.lin_edge.weight, self.att.view(C), csr, 2,
                           self.negative_slope) + self.bias
