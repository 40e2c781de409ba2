"""SURVEY 8(a) row `MPNNLSTM` (legacy API, model/model.py:613-684) and `MPNNLSTMI` (:727-802): the product classes against a
dense-matrix restatement written here (GCN with self loops as one normalised adjacency matrix, nn.LSTM over time), against
the oracle's GCNConv, and -- in the build container -- against the UNMODIFIED reference classes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import blob_frames, rel_err


def _graph(seed=0, H=20, W=26):
    from oracle import graph_ref as G
    rng = np.random.default_rng(seed)
    x = blob_frames(rng, 1, H, W, c=1)
    mask = rng.random((H, W)) > 0.85
    g = G.image_to_graph(G.add_positional_encoding(torch.from_numpy(x)), thresh=0.5, max_grid_size=8, mask=mask, use_edge_attrs=False)
    return g["edge_index"], g["edge_attrs"], g["data"].shape[1]


def _dense_gcn(x, W, b, ei, ew, n):
    """PyG GCNConv(add_self_loops=True): existing loops keep their weight, missing ones get 1; deg over incoming weights."""
    A = torch.zeros(n, n, dtype=x.dtype)
    A[ei[1], ei[0]] = ew.to(x.dtype)
    d = torch.diagonal(A).clone()
    has = torch.zeros(n, dtype=torch.bool)
    has[ei[0][ei[0] == ei[1]]] = True
    A[range(n), range(n)] = torch.where(has, d, torch.ones(n, dtype=x.dtype))
    dis = A.sum(1).pow(-0.5)
    dis[torch.isinf(dis)] = 0
    return (dis[:, None] * A * dis[None, :]) @ (x @ W.T) + b


def _dense_mpnnlstm(m, X, ei, ew):
    """model/model.py:640-684 with the three GCNConvs as dense matrices."""
    n = X.shape[1]
    frames = []
    for t in range(X.shape[0]):
        h = X[t]
        for conv, norm in ((m.convolution1, m.bn1), (m.convolution2, m.bn2), (m.convolution3, m.bn3)):
            h = norm(F.relu(_dense_gcn(h, conv.lin.weight, conv.bias, ei, ew, n)))
        frames.append(h)
    _, (h, _) = m.recurrents(torch.stack(frames))
    h = torch.cat([F.relu(h[-1]), X[:, :, 0].T], dim=-1)
    return torch.sigmoid(m.lin2(F.relu(m.lin1(h))))


def test_mpnnlstm_matches_dense_restatement(be, monkeypatch):
    import quadtree_mpnnlstm_b200.model as M
    # the dense nn.LSTM / nn.Linear of this legacy model are stock PyTorch modules (as in the reference); keep cuDNN / cuBLAS
    # in fp32 so the comparison sees the graph-convolution kernels, not TF32 rounding of the library LSTM
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    ei, ew, n = _graph(1)
    T, F_in, hid = 3, 4, 16
    torch.manual_seed(0)
    cpu = M.MPNNLSTM(hid, dropout=0.0, input_timesteps=T, input_features=F_in)       # parameter holder for the dense restatement
    gpu = be.dev(M.MPNNLSTM(hid, dropout=0.0, input_timesteps=T, input_features=F_in))
    gpu.load_state_dict(cpu.state_dict())
    cpu.eval()
    gpu.train()                     # dropout is 0.0, so train() changes nothing -- but cuDNN's LSTM backward insists on it
    X = torch.randn(T, n, F_in)
    xa, xb = X.clone().requires_grad_(True), be.dev(X.clone()).requires_grad_(True)
    ya = _dense_mpnnlstm(cpu, xa, ei, ew)
    yb = gpu(xb, be.dev(ei), be.dev(ew))
    assert ya.shape == yb.shape == (n, 1)
    assert rel_err(yb, ya) < 2e-5, rel_err(yb, ya)
    w = torch.randn_like(ya)
    (ya * w).sum().backward()
    (yb * be.dev(w)).sum().backward()
    assert rel_err(xb.grad, xa.grad) < 2e-4
    for (k, pa), (_, pb) in zip(cpu.named_parameters(), gpu.named_parameters()):
        diff = float((pa.grad - pb.grad.cpu()).abs().max())
        assert diff <= 1e-3 * max(float(pa.grad.abs().max()), 1e-3) + 2e-6, f"grad {k}: {diff}"
    # state-dict keys of the reference class
    assert {"convolution1.lin.weight", "convolution3.bias", "bn2.weight", "recurrents.weight_ih_l3", "lin1.weight",
            "lin2.bias"} <= set(gpu.state_dict())


def test_mpnnlstm_matches_unmodified_reference(be):
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    import quadtree_mpnnlstm_b200.model as M
    ref = load_reference()
    ei, ew, n = _graph(2)
    T, F_in, hid = 3, 4, 8
    torch.manual_seed(3)
    a = ref.model.MPNNLSTM(hid, dropout=0.0, input_timesteps=T, input_features=F_in).eval()
    b = be.dev(M.MPNNLSTM(hid, dropout=0.0, input_timesteps=T, input_features=F_in)).eval()
    b.load_state_dict(a.state_dict())
    X = torch.randn(T, n, F_in)
    ya = a(X, ei, ew)
    yb = b(be.dev(X), be.dev(ei), be.dev(ew))
    assert rel_err(yb, ya) < 2e-5


def test_mpnnlstmi_cell_stack(be):
    """MPNNLSTMI (model/model.py:727-802): GConvLSTM stack (default GCNConv cells) + BatchNorm1d + two linears; the wiring
    quirk `C=hs[1]` of the first layer (:760) is kept.  The reference's forward builds the result and returns nothing; the
    product returns it."""
    import quadtree_mpnnlstm_b200.model as M
    from oracle import cell_ref as R
    ei, ew, n = _graph(3)
    T, F_in, hid = 3, 4, 16
    torch.manual_seed(1)
    gpu = be.dev(M.MPNNLSTMI(hid, dropout=0.0, input_timesteps=T, input_features=F_in, n_layers=2)).eval()
    cells = [R.GConvLSTM(F_in, hid), R.GConvLSTM(hid, hid)]
    for c, g in zip(cells, gpu.recurrents):
        c.load_state_dict({k: v.cpu() for k, v in g.state_dict().items()})
    X = torch.randn(T, n, F_in)
    hs, cs = [None, None], [None, None]
    for x in X:
        _, h, c = cells[0](x, ei, ew, H=hs[0], C=hs[1])
        hs[0], cs[0] = h, c
        _, h, c = cells[1](hs[0], ei, ew, H=hs[1], C=cs[1])
        hs[1], cs[1] = h, c
    z = F.relu(hs[-1])
    z = F.batch_norm(z, None, None, gpu.bn1.weight.cpu(), gpu.bn1.bias.cpu(), training=True, eps=gpu.bn1.eps)
    want = torch.sigmoid(F.linear(F.relu(F.linear(z, gpu.lin1.weight.cpu(), gpu.lin1.bias.cpu())), gpu.lin2.weight.cpu(), gpu.lin2.bias.cpu()))
    got = gpu(be.dev(X), be.dev(ei), be.dev(ew))
    assert got.shape == (n, 1) and rel_err(got, want) < 5e-5, rel_err(got, want)
