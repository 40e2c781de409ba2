"""CPU: the restated PyG convolutions (oracle/convs_ref.py, PARITY UNPINNED) against dense-matrix formulas."""
import math

import torch

from oracle import convs_ref as R


def _graph(n=9, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand(n, n, generator=g) < 0.4)
    a = a | a.T
    a.fill_diagonal_(False)
    a[0, 0] = a[3, 3] = True                       # self-loops exist in the reference's edge lists
    src, dst = torch.nonzero(a, as_tuple=True)
    w = torch.rand(src.numel(), generator=g) + 0.1
    w = (w + w[torch.argsort(torch.argsort(dst * n + src))][torch.argsort(torch.argsort(src * n + dst))]) / 2
    return torch.stack([src, dst]), w, n


def _dense_adj(ei, w, n):
    A = torch.zeros(n, n, dtype=torch.float64)
    A[ei[1], ei[0]] = w.double()                   # A[target, source]
    return A


def test_gcn_dense():
    ei, w, n = _graph()
    conv = R.GCNConv(5, 3, add_self_loops=False).double()
    x = torch.randn(n, 5, dtype=torch.float64)
    A = _dense_adj(ei, w, n)
    deg = A.sum(1)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    ref = (dis[:, None] * A * dis[None, :]) @ x @ conv.lin.weight.T + conv.bias
    assert torch.allclose(conv(x, ei, w.double()), ref, atol=1e-10)


def test_cheb_dense():
    ei, w, n = _graph(seed=1)
    conv = R.ChebConv(4, 6, K=3).double()
    x = torch.randn(n, 4, dtype=torch.float64)
    A = _dense_adj(ei, w, n)
    A.fill_diagonal_(0)                            # ChebConv removes self-loops
    deg = A.sum(0)                                 # by source; the graph is symmetric
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    L_hat = -(dis[:, None] * A * dis[None, :])     # 2L/lambda_max - I with lambda_max = 2, L = I - D^-1/2 A D^-1/2
    t0, t1 = x, L_hat @ x
    t2 = 2 * L_hat @ t1 - t0
    ref = t0 @ conv.lins[0].weight.T + t1 @ conv.lins[1].weight.T + t2 @ conv.lins[2].weight.T + conv.bias
    assert torch.allclose(conv(x, ei, w.double()), ref, atol=1e-10)
    assert torch.allclose(conv(x, ei), conv(x, ei, torch.ones(ei.shape[1], dtype=torch.float64)), atol=1e-12)


def test_transformer_dense():
    ei, _, n = _graph(seed=2)
    torch.manual_seed(0)
    conv = R.TransformerConv(5, 4, heads=1, concat=False, edge_dim=2, dropout=0.1).double().eval()
    x = torch.randn(n, 5, dtype=torch.float64)
    ea = torch.randn(ei.shape[1], 2, dtype=torch.float64)
    q, k, v = conv.lin_query(x), conv.lin_key(x), conv.lin_value(x)
    out = conv.lin_skip(x).clone()
    for i in range(n):
        idx = torch.nonzero(ei[1] == i).squeeze(1)
        if idx.numel() == 0:
            continue
        e = conv.lin_edge(ea[idx])
        logits = ((k[ei[0, idx]] + e) @ q[i]) / math.sqrt(4)
        alpha = torch.softmax(logits, 0)
        out[i] += (alpha[:, None] * (v[ei[0, idx]] + e)).sum(0)
    assert torch.allclose(conv(x, ei, ea), out, atol=1e-10)
    # invariances the CUDA kernel relies on: lin_key.bias does not change the output
    with torch.no_grad():
        conv.lin_key.bias.add_(3.0)
    assert torch.allclose(conv(x, ei, ea), out, atol=1e-9)


def test_multi_head_transformer_dense():
    """heads = 3, concat = True (+ the reference's MHTransformerConv output projection, model/model.py:26-37): every head is
    a softmax over the in-edges with its own C-wide slices and the 1 / sqrt(C) scale."""
    ei, _, n = _graph(seed=5)
    torch.manual_seed(1)
    H, C = 3, 4
    conv = R.MHTransformerConv(5, C, heads=H, edge_dim=2, dropout=0.1).double().eval()
    x = torch.randn(n, 5, dtype=torch.float64)
    ea = torch.randn(ei.shape[1], 2, dtype=torch.float64)
    q, k, v = conv.lin_query(x), conv.lin_key(x), conv.lin_value(x)
    cat = conv.lin_skip(x).clone()
    for i in range(n):
        idx = torch.nonzero(ei[1] == i).squeeze(1)
        if idx.numel() == 0:
            continue
        e = conv.lin_edge(ea[idx])
        for h in range(H):
            sl = slice(h * C, (h + 1) * C)
            logits = ((k[ei[0, idx]][:, sl] + e[:, sl]) @ q[i, sl]) / math.sqrt(C)
            alpha = torch.softmax(logits, 0)
            cat[i, sl] += (alpha[:, None] * (v[ei[0, idx]][:, sl] + e[:, sl])).sum(0)
    ref = cat @ conv.lin.weight.T + conv.lin.bias
    assert torch.allclose(conv(x, ei, ea), ref, atol=1e-10)
    assert conv(x, ei, ea).shape == (n, C)
    keys = set(conv.state_dict())
    assert {"lin.weight", "lin.bias", "lin_skip.weight", "lin_edge.weight", "lin_query.bias"} <= keys
    assert conv.lin_skip.weight.shape == (H * C, 5) and conv.lin.weight.shape == (C, H * C)


def test_init_matches_torch_linear_stream():
    """PyG Linear's default init consumes the RNG like nn.Linear: the product's parameter holders follow it."""
    torch.manual_seed(4)
    a = R.Linear(7, 5)
    torch.manual_seed(4)
    b = torch.nn.Linear(7, 5)
    assert torch.equal(a.weight, b.weight) and torch.equal(a.bias, b.bias)


def test_gat_dense():
    """GATConv / GATv2Conv restatements (heads=1, edge_dim=2) against per-node dense formulas: softmax over the in-edges plus one
    self loop whose attributes are the mean of the node's incoming attributes; existing self loops are replaced."""
    import torch.nn.functional as F
    ei, _, n = _graph(seed=7)
    ei = torch.cat([ei, torch.tensor([[0, 2], [0, 2]])], dim=1)          # two explicit self loops: must be dropped and re-added
    torch.manual_seed(2)
    C = 4
    x = torch.randn(n, 5, dtype=torch.float64)
    ea = torch.randn(ei.shape[1], 2, dtype=torch.float64)
    for kind in ("GATConv", "GATv2Conv"):
        conv = getattr(R, kind)(5, C, heads=1, edge_dim=2).double().eval()
        with torch.no_grad():
            conv.bias.copy_(torch.randn(C))
        out = conv.bias.expand(n, C).clone()
        for i in range(n):
            idx = torch.nonzero((ei[1] == i) & (ei[0] != i)).squeeze(1)
            src = torch.cat([ei[0, idx], torch.tensor([i])])
            attrs = torch.cat([ea[idx], ea[idx].mean(0, keepdim=True) if idx.numel() else torch.zeros(1, 2, dtype=torch.float64)])
            if kind == "GATConv":
                xs = conv.lin_src(x)
                logit = (xs[src] * conv.att_src.view(C)).sum(1) + (xs[i] * conv.att_dst.view(C)).sum() \
                    + (conv.lin_edge(attrs) * conv.att_edge.view(C)).sum(1)
                alpha = torch.softmax(F.leaky_relu(logit, 0.2), 0)
                out[i] += (alpha[:, None] * xs[src]).sum(0)
            else:
                xl, xr = conv.lin_l(x), conv.lin_r(x)
                m = F.leaky_relu(xl[src] + xr[i] + conv.lin_edge(attrs), 0.2)
                alpha = torch.softmax((m * conv.att.view(C)).sum(1), 0)
                out[i] += (alpha[:, None] * xl[src]).sum(0)
        assert torch.allclose(conv(x, ei, ea), out, atol=1e-10), kind
    keys = set(R.GATConv(5, C, heads=1, edge_dim=2).state_dict())
    assert {"att_src", "att_dst", "att_edge", "lin_src.weight", "lin_dst.weight", "lin_edge.weight", "bias"} == keys
    keys = set(R.GATv2Conv(5, C, heads=1, edge_dim=2).state_dict())
    assert {"att", "lin_l.weight", "lin_l.bias", "lin_r.weight", "lin_r.bias", "lin_edge.weight", "bias"} == keys
