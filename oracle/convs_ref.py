"""Oracle restatement of the PyG convolutions the reference can select (model/model.py:39-57).  TEST INFRASTRUCTURE.

PARITY UNPINNED: the arithmetic belongs to ``torch-geometric==2.2.0`` (reference
requirements.txt:13), which is not vendored in /root/reference and not installable here.
What follows restates PyG 2.2.0's published algorithm; call sites in the reference:
model/model.py:39-57 (CONVOLUTIONS / CONVOLUTION_KWARGS), model/model.py:72-73, :96,
model/seq2seq.py:117-121.  ``tests/test_oracle_convs.py`` validates each conv against its
dense-matrix formula.

Conventions (PyG, flow = source_to_target): ``edge_index[0]`` = source j, ``edge_index[1]`` =
target i; messages are aggregated at the target.  Parameter names match PyG's state-dict keys
 so that checkpoints are interchangeable.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def glorot(t):
    """torch_geometric.nn.inits.glorot: U(-a, a), a = sqrt(6 / (fan_in + fan_out))."""
    if t is not None:
        a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        t.data.uniform_(-a, a)


def zeros(t):
    if t is not None:
        t.data.fill_(0)


class Linear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear.  Default initialiser = torch's nn.Linear
    (kaiming_uniform(a=sqrt(5)) weight, U(+-1/sqrt(fan_in)) bias); 'glorot' on request."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer, self.bias_initializer = weight_initializer, bias_initializer
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            glorot(self.weight)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            if self.bias_initializer == "zeros":
                zeros(self.bias)
            else:
                bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
                nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


def scatter_add_rows(src, index, n):
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add(0, index, src)


def _inv_sqrt(deg):
    r = deg.pow(-0.5)
    return r.masked_fill(r == float("inf"), 0.0)


def gcn_norm(edge_index, edge_weight, n, add_self_loops):
    """torch_geometric.nn.conv.gcn_conv.gcn_norm (improved=False, source_to_target):
    deg = scatter_add(w, target); norm = deg^-1/2[src] * w * deg^-1/2[dst], inf -> 0."""
    row, col = edge_index[0], edge_index[1]
    if edge_weight is None:
        edge_weight = torch.ones(row.numel(), dtype=torch.float32)
    if add_self_loops:
        # add_remaining_self_loops: existing loops keep their weight, missing ones get 1
        loop_w = torch.ones(n, dtype=edge_weight.dtype)
        is_loop = row == col
        loop_w[row[is_loop]] = edge_weight[is_loop]
        keep = ~is_loop
        ar = torch.arange(n, dtype=row.dtype)
        row, col = torch.cat([row[keep], ar]), torch.cat([col[keep], ar])
        edge_weight = torch.cat([edge_weight[keep], loop_w])
    deg = scatter_add_rows(edge_weight, col, n)
    dis = _inv_sqrt(deg)
    return torch.stack([row, col]), dis[row] * edge_weight * dis[col]


class GCNConv(nn.Module):
    """out_i = sum_{j->i} norm_ij (W x_j) + b, W without bias (glorot), b zeros."""

    def __init__(self, in_channels, out_channels, add_self_loops=True, bias=True, **_):
        super().__init__()
        self.in_channels, self.out_channels, self.add_self_loops = in_channels, out_channels, add_self_loops
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index, edge_weight=None):
        n = x.shape[0]
        ei, norm = gcn_norm(edge_index, edge_weight, n, self.add_self_loops)
        h = self.lin(x)
        out = scatter_add_rows(norm[:, None] * h[ei[0]], ei[1], n)
        return out + self.bias if self.bias is not None else out


def cheb_norm(edge_index, edge_weight, n):
    """ChebConv.__norm__ with normalization='sym', lambda_max = 2.0: drop self-loops,
    L = I - D^-1/2 A D^-1/2 (deg = scatter_add(w, source)), scale 2L/lambda_max, then append
    self-loops filled with -1.  Entry list order: off-diagonal, +1 diagonal, -1 diagonal."""
    row, col = edge_index[0], edge_index[1]
    if edge_weight is None:
        edge_weight = torch.ones(row.numel(), dtype=torch.float32)
    keep = row != col
    row, col, w = row[keep], col[keep], edge_weight[keep]
    dis = _inv_sqrt(scatter_add_rows(w, row, n))
    w = dis[row] * w * dis[col]
    ar = torch.arange(n, dtype=row.dtype)
    row, col = torch.cat([row, ar]), torch.cat([col, ar])
    w = torch.cat([-w, torch.ones(n, dtype=w.dtype)])
    w = (2.0 * w) / torch.tensor(2.0, dtype=w.dtype)
    w = w.masked_fill(w == float("inf"), 0.0)
    row, col = torch.cat([row, ar]), torch.cat([col, ar])
    w = torch.cat([w, torch.full((n,), -1.0, dtype=w.dtype)])
    return torch.stack([row, col]), w


class ChebConv(nn.Module):
    """T0 = x, T1 = L^ x, Tk = 2 L^ T(k-1) - T(k-2); out = sum_k lins[k](Tk) + bias."""

    def __init__(self, in_channels, out_channels, K=3, normalization="sym", bias=True, **_):
        super().__init__()
        assert K > 0 and normalization == "sym"
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.lins = nn.ModuleList([Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
                                   for _ in range(K)])
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index, edge_weight=None):
        n = x.shape[0]
        ei, norm = cheb_norm(edge_index, edge_weight, n)

        def prop(z):
            return scatter_add_rows(norm[:, None] * z[ei[0]], ei[1], n)

        t0 = x
        out = self.lins[0](t0)
        if self.K > 1:
            t1 = prop(x)
            out = out + self.lins[1](t1)
        for lin in self.lins[2:]:
            t2 = 2.0 * prop(t1) - t0
            out = out + lin(t2)
            t0, t1 = t1, t2
        return out + self.bias if self.bias is not None else out


def segment_softmax(score, index, n):
    """torch_geometric.utils.softmax: subtract the per-target max, exp, divide by sum + 1e-16."""
    mx = torch.full((n,), float("-inf"), dtype=score.dtype).scatter_reduce(0, index, score.detach(), "amax",
                                                                          include_self=True)
    ex = (score - mx[index]).exp()
    den = scatter_add_rows(ex, index, n) + 1e-16
    return ex / den[index]


class TransformerConv(nn.Module):
    """PyG 2.2.0 ``TransformerConv`` with beta=False, root_weight=True (CONVOLUTION_KWARGS, model/model.py:51-52), H heads:
    q_i = Wq x_i + bq; k_ij = Wk x_j + bk + We e_ij; v_ij = Wv x_j + bv + We e_ij, each viewed as [H, C];
    alpha^h = softmax_j(q_i^h . k_ij^h / sqrt(C)); dropout(alpha); out_i^h = sum_j alpha^h v_ij^h;
    concat=True: heads side by side [H*C] + lin_skip (in -> H*C); concat=False: mean over heads + lin_skip (in -> C)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0,
                 edge_dim=None, bias=True, root_weight=True, **_):
        super().__init__()
        assert not beta and root_weight, "only the configurations the reference selects"
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.dropout, self.edge_dim = concat, dropout, edge_dim
        self.lin_key = Linear(in_channels, heads * out_channels)
        self.lin_query = Linear(in_channels, heads * out_channels)
        self.lin_value = Linear(in_channels, heads * out_channels)
        self.lin_edge = Linear(edge_dim, heads * out_channels, bias=False) if edge_dim is not None else None
        self.lin_skip = Linear(in_channels, heads * out_channels if concat else out_channels, bias=bias)

    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=None):
        assert not return_attention_weights, "attention weights are never requested on the hot path"
        n, c, h = x.shape[0], self.out_channels, self.heads
        src, dst = edge_index[0], edge_index[1]
        q = self.lin_query(x)[dst].view(-1, h, c)
        k = self.lin_key(x)[src].view(-1, h, c)
        v = self.lin_value(x)[src].view(-1, h, c)
        if self.lin_edge is not None:
            assert edge_attr is not None
            e = self.lin_edge(edge_attr).view(-1, h, c)
            k = k + e
        alpha = (q * k).sum(-1) / math.sqrt(c)                                   # [E, H]
        alpha = torch.stack([segment_softmax(alpha[:, i], dst, n) for i in range(h)], dim=1)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = v + e if self.lin_edge is not None else v
        out = scatter_add_rows(msg * alpha[:, :, None], dst, n)                  # [N, H, C]
        out = out.reshape(n, h * c) if self.concat else out.mean(dim=1)
        return out + self.lin_skip(x)


class MHTransformerConv(TransformerConv):
    """The reference's own subclass (model/model.py:26-37): the multi-head TransformerConv followed by
    ``lin`` (heads * out -> out)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0, edge_dim=None,
                 bias=True, root_weight=True, **kw):
        super().__init__(in_channels, out_channels, heads, concat, beta, dropout, edge_dim, bias, root_weight, **kw)
        self.lin = Linear(out_channels * heads, out_channels)

    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=None):
        return self.lin(super().forward(x, edge_index, edge_attr))


def add_self_loops_mean(edge_index, edge_attr, n):
    """torch_geometric.utils: remove_self_loops, then add_self_loops(fill_value='mean') -- the new loop of node i carries
    scatter(edge_attr, edge_index[1], reduce='mean')[i] (zeros for a node without in-edges).  GATConv.forward / GATv2Conv.forward."""
    keep = edge_index[0] != edge_index[1]
    ei, ea = edge_index[:, keep], edge_attr[keep]
    cnt = torch.zeros(n, dtype=ea.dtype).index_add(0, ei[1], torch.ones(ei.shape[1], dtype=ea.dtype))
    mean = scatter_add_rows(ea, ei[1], n) / cnt.clamp(min=1)[:, None]
    ar = torch.arange(n, dtype=ei.dtype)
    return torch.cat([ei, torch.stack([ar, ar])], dim=1), torch.cat([ea, mean])


class GATConv(nn.Module):
    """PyG 2.2.0 ``GATConv`` as the reference configures it (model/model.py:43, 55: heads=1, edge_dim=2; PyG defaults
    concat=True, negative_slope=0.2, dropout=0, add_self_loops=True, fill_value='mean', bias=True), H heads:
    x' = lin_src(x) viewed [H, C] (lin_dst IS lin_src for an int in_channels); alpha_src = (x' * att_src).sum(-1), alpha_dst likewise;
    self loops with mean attributes; alpha_e = alpha_src[j] + alpha_dst[i] + (lin_edge(e) * att_edge).sum(-1);
    leaky_relu; softmax over the in-edges; out_i = sum_e alpha_e x'_j, heads concatenated, + bias."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0, add_self_loops=True,
                 edge_dim=None, fill_value="mean", bias=True, **_):
        super().__init__()
        assert concat and add_self_loops and fill_value == "mean" and edge_dim is not None and bias
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.dropout = negative_slope, dropout
        self.lin_src = Linear(in_channels, heads * out_channels, bias=False, weight_initializer="glorot")
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.lin_edge = Linear(edge_dim, heads * out_channels, bias=False, weight_initializer="glorot")
        self.att_edge = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        self.lin_src.reset_parameters()
        self.lin_dst.reset_parameters()
        self.lin_edge.reset_parameters()
        glorot(self.att_src)
        glorot(self.att_dst)
        glorot(self.att_edge)
        zeros(self.bias)

    def forward(self, x, edge_index, edge_attr=None):
        n, h, c = x.shape[0], self.heads, self.out_channels
        xs = self.lin_src(x).view(-1, h, c)
        a_src, a_dst = (xs * self.att_src).sum(-1), (xs * self.att_dst).sum(-1)                 # [N, H]
        ei, ea = add_self_loops_mean(edge_index, edge_attr, n)
        src, dst = ei[0], ei[1]
        a_edge = (self.lin_edge(ea).view(-1, h, c) * self.att_edge).sum(-1)
        alpha = F.leaky_relu(a_src[src] + a_dst[dst] + a_edge, self.negative_slope)
        alpha = torch.stack([segment_softmax(alpha[:, i], dst, n) for i in range(h)], dim=1)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = scatter_add_rows(xs[src] * alpha[:, :, None], dst, n).reshape(n, h * c)
        return out + self.bias


class GATv2Conv(nn.Module):
    """PyG 2.2.0 ``GATv2Conv`` as the reference configures it (model/model.py:44, 56; share_weights=False): x_l = lin_l(x),
    x_r = lin_r(x) (both with bias, glorot weights); self loops with mean attributes; m_e = x_l[j] + x_r[i] + lin_edge(e);
    alpha_e = (leaky_relu(m_e) * att).sum(-1); softmax over the in-edges; out_i = sum_e alpha_e x_l[j], + bias."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0, add_self_loops=True,
                 edge_dim=None, fill_value="mean", bias=True, share_weights=False, **_):
        super().__init__()
        assert concat and add_self_loops and fill_value == "mean" and edge_dim is not None and bias and not share_weights
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.dropout = negative_slope, dropout
        self.lin_l = Linear(in_channels, heads * out_channels, bias=True, weight_initializer="glorot")
        self.lin_r = Linear(in_channels, heads * out_channels, bias=True, weight_initializer="glorot")
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.lin_edge = Linear(edge_dim, heads * out_channels, bias=False, weight_initializer="glorot")
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()
        self.lin_edge.reset_parameters()
        glorot(self.att)
        zeros(self.bias)

    def forward(self, x, edge_index, edge_attr=None):
        n, h, c = x.shape[0], self.heads, self.out_channels
        xl, xr = self.lin_l(x).view(-1, h, c), self.lin_r(x).view(-1, h, c)
        ei, ea = add_self_loops_mean(edge_index, edge_attr, n)
        src, dst = ei[0], ei[1]
        m = F.leaky_relu(xl[src] + xr[dst] + self.lin_edge(ea).view(-1, h, c), self.negative_slope)
        alpha = (m * self.att).sum(-1)
        alpha = torch.stack([segment_softmax(alpha[:, i], dst, n) for i in range(h)], dim=1)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = scatter_add_rows(xl[src] * alpha[:, :, None], dst, n).reshape(n, h * c)
        return out + self.bias


CONVOLUTIONS = {"GCNConv": GCNConv, "TransformerConv": TransformerConv, "ChebConv": ChebConv,
                "MHTransformerConv": MHTransformerConv, "GATConv": GATConv, "GATv2Conv": GATv2Conv}
