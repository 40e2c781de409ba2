"""B200 mirror of the reference's ``model/model.py``: ``GraphConv``, ``GConvLSTM`` (the recurrent
cell) and the legacy ``MPNNLSTM``, with the reference's constructor arguments, forward signatures and
state-dict keys (model/model.py:59-97, 263-463, 613-684).

``GConvLSTM.forward`` runs the eight GraphConv stacks of the cell as grouped kernels: the four
``conv_x_*`` stacks share X and the four ``conv_h_*`` stacks share H, so layer 1 is two grouped
launches and deeper layers one launch over all eight; the gate math (sigmoid / tanh / peepholes) and,
when the encoder / decoder ask for it, the LayerNorms and the decoder-head input are one more kernel.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .convs import (ChebConv, GATConv, GATv2Conv, GCNConv, MHTransformerConv, TransformerConv, cheb_basis,
                    pack_tconv)
from . import fused as _fused
from .cheb_cell import ChebCellFn, ChebLstmCellFn, pack_linear_group
from .graph_csr import get_csr
from .ops import LstmGatesFn, NodeLinearFn, SpmmFn, TConvFn, next_seed

CHEB_CELL_FN = os.environ.get("QMP_CHEB_CELL_FN", "1") != "0"     # ChebConv / GCNConv cells as one autograd node (cheb_cell.py); 0: the
                                                                  # modular SpmmFn / NodeLinearFn path (the cross-check)

CONVOLUTIONS = {
    'GCNConv': GCNConv,
    'TransformerConv': TransformerConv,
    'MHTransformerConv': MHTransformerConv,
    'ChebConv': ChebConv,
    'GATConv': GATConv,
    'GATv2Conv': GATv2Conv,
    'Dummy': None
}

CONVOLUTION_KWARGS = {
    'GCNConv': dict(add_self_loops=False),
    'TransformerConv': dict(heads=1, edge_dim=2, dropout=0.1, concat=False),
    'MHTransformerConv': dict(heads=3, edge_dim=2, dropout=0.1),
    'ChebConv': dict(K=3, normalization='sym', bias=True),
    'GATConv': dict(heads=1, edge_dim=2),
    'GATv2Conv': dict(heads=1, edge_dim=2),
    'Dummy': dict(),
}

GATES = ("i", "f", "c", "o")

# Packed parameters are cached for the duration of one driver forward pass ("epoch"): the driver bumps
# the epoch at the start of every forward; stand-alone calls of a cell bump it themselves.
_epoch = [0]


def new_epoch():
    _epoch[0] += 1
    return _epoch[0]


class GraphConv(nn.Module):
    """A stack of ``n_layers`` convolutions with no nonlinearity in between (model/model.py:59-97)."""

    def __init__(self, convolution_type, in_channels, out_channels, n_layers):
        super(GraphConv, self).__init__()
        self.convolution_type = convolution_type
        self.n_layers = n_layers
        conv_func = CONVOLUTIONS[convolution_type]
        conv_kwargs = CONVOLUTION_KWARGS[convolution_type]
        if convolution_type != 'Dummy':
            self.convolutions = nn.ModuleList(
                [conv_func(in_channels, out_channels, **conv_kwargs)] +
                [conv_func(out_channels, out_channels, **conv_kwargs) for _ in range(n_layers - 1)])
        else:
            self.n_layers = 0

    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=False):
        for i in range(self.n_layers):
            x = self.convolutions[i](x, edge_index, edge_attr)
        return x


def _pack_linear_group(convs, kind):
    """[G, C, K] weight and [G, C] bias for a group of GCN (K = D) or Cheb (K = 3D) convs."""
    if kind == "GCNConv":
        W = torch.stack([c.lin.weight for c in convs])
    else:
        W = torch.stack([torch.cat([lin.weight for lin in c.lins], dim=1) for c in convs])
    b = torch.stack([c.bias for c in convs])
    return W.contiguous(), b.contiguous()


class GConvLSTM(nn.Module):
    r"""Peephole graph-convolutional LSTM cell (model/model.py:263-463).

    Args:
        in_channels (int): Number of input features.
        out_channels (int): Number of output features.
    """

    def __init__(self, in_channels, out_channels, n_conv_layers=1, convolution_type='GCNConv', name='GConvLSTM'):
        super(GConvLSTM, self).__init__()
        assert convolution_type in CONVOLUTIONS
        self.convolution_type = convolution_type
        self.n_conv_layers = n_conv_layers
        self.return_attention_weights = False
        self.name = name
        self.in_channels = in_channels
        self.out_channels = out_channels
        for g in GATES:      # creation order follows model/model.py:294-373 (same RNG stream for a seed)
            setattr(self, f"conv_x_{g}", GraphConv(convolution_type, in_channels, out_channels, n_conv_layers))
            setattr(self, f"conv_h_{g}", GraphConv(convolution_type, out_channels, out_channels, n_conv_layers))
            if g != "c":
                setattr(self, f"w_c_{g}", nn.Parameter(torch.zeros(1, out_channels)))
            setattr(self, f"b_{g}", nn.Parameter(torch.zeros(1, out_channels)))
        self._cache = {}

    # ---- packed parameters ------------------------------------------------------------------
    def _cached(self, key, epoch, build):
        hit = self._cache.get(key)
        if hit is not None and hit[0] == epoch:
            return hit[1]
        val = build()
        self._cache[key] = (epoch, val)
        return val

    def _convs(self, which, layer):
        return [getattr(self, f"conv_{which}_{g}").convolutions[layer] for g in GATES]

    def _gate_params(self, epoch, norm_h=None, norm_c=None, norm_o=None):
        def build():
            C = self.out_channels
            one = torch.ones(C, device=self.b_i.device)
            zero = torch.zeros(C, device=self.b_i.device)
            rows = [self.w_c_i.view(-1), self.w_c_f.view(-1), self.w_c_o.view(-1), self.b_i.view(-1), self.b_f.view(-1),
                    self.b_c.view(-1), self.b_o.view(-1)]
            for nm in (norm_h, norm_c, norm_o):
                rows += [nm.weight, nm.bias] if nm is not None else [one, zero]
            return _fused.RowsPackFn.apply(len(rows), *rows)     # == torch.stack(rows), copy-free parameter gradients
        if self._fusable() or self._cheb_cell():   # consumed once per timestep by one autograd node: gradients accumulate in place (fused._GradAccum)
            return self._cached(("gates", id(norm_h), id(norm_c), id(norm_o)), epoch, lambda: _fused.shared_pack(build()))
        return self._cached(("gates", id(norm_h), id(norm_c), id(norm_o)), epoch, build)

    # ---- the eight stacks -------------------------------------------------------------------
    def _pre_activations(self, X, H, csr, epoch):
        """P [N, 4C] = conv_x_*(X) + conv_h_*(H), gate order i, f, c, o."""
        S, kind, C = self.n_conv_layers, self.convolution_type, self.out_channels
        if kind == 'TransformerConv':
            p = self.conv_x_i.convolutions[0].dropout if self.training else 0.0
            seed = (lambda: next_seed()) if p > 0 else (lambda: 0)
            pk = lambda which, l: self._cached(("t", which, l), epoch, lambda: pack_tconv(self._convs(which, l)))
            px = TConvFn.apply(X, *pk("x", 0), csr, True, p, seed(), False, None)
            if S == 1:
                return TConvFn.apply(H, *pk("h", 0), csr, True, p, seed(), False, px)
            cur = torch.cat([px, TConvFn.apply(H, *pk("h", 0), csr, True, p, seed(), False, None)], dim=1)
            for l in range(1, S):
                pka = self._cached(("t", "all", l), epoch,
                                   lambda: pack_tconv(self._convs("x", l) + self._convs("h", l)))
                cur = TConvFn.apply(cur, *pka, csr, False, p, seed(), False, None)
            return cur[:, :4 * C] + cur[:, 4 * C:]
        if self._cheb_cell():
            # one autograd node for the eight stacks (cheb_cell.py): basis blocks, grouped GEMMs and the in-place weight gradients
            mode, K, packs = self._cheb_packs(epoch)
            return ChebCellFn.apply(X.contiguous(), H, csr, mode, K, S, C, *packs)
        if kind in ('GCNConv', 'ChebConv'):
            mode = "gcn" if kind == 'GCNConv' else "cheb"
            K = 1 if kind == 'GCNConv' else self.conv_x_i.convolutions[0].K

            def basis(z):                       # [N, w] -> the K propagated copies, each [N, w]
                if kind == 'GCNConv':
                    return [SpmmFn.apply(z, None, csr, "gcn", 1.0, 0.0)]
                return cheb_basis(z, csr, K)

            pk = lambda which, l: self._cached((mode, which, l), epoch,
                                               lambda: _pack_linear_group(self._convs(which, l), kind))
            outs = []
            for which, inp in (("x", X), ("h", H)):
                W, b = pk(which, 0)
                outs.append(NodeLinearFn.apply(torch.cat(basis(inp), dim=1) if K > 1 else basis(inp)[0], W, b, True))
            if S == 1:
                return outs[0] + outs[1]
            cur = torch.cat(outs, dim=1)                                        # [N, 8C], conv g owns columns g*C..
            for l in range(1, S):
                W, b = self._cached((mode, "all", l), epoch,
                                    lambda: _pack_linear_group(self._convs("x", l) + self._convs("h", l), kind))
                ts = basis(cur)                                                 # K tensors [N, 8C]
                if K > 1:   # per conv the K blocks must be adjacent: [N, 8, K, C]
                    N = cur.shape[0]
                    inp = torch.stack([t.view(N, 8, C) for t in ts], dim=2).reshape(N, 8 * K * C)
                else:
                    inp = ts[0]
                cur = NodeLinearFn.apply(inp, W, b, False)
            return cur[:, :4 * C] + cur[:, 4 * C:]
        # any other conv type (MHTransformerConv, the GAT family): the eight stacks one module call at a time -- each call runs
        # on its own qmp kernels (convs.py); only the batching of the eight stacks into grouped launches is missing
        ei, ea = csr._keepalive              # the tensors the CSR was built from: the module calls hit the CSR cache by identity
        outs = []
        for g in GATES:
            outs.append(getattr(self, f"conv_x_{g}")(X, ei, ea) + getattr(self, f"conv_h_{g}")(H, ei, ea))
        return torch.cat(outs, dim=1)

    # ---- ChebConv / GCNConv cell as one autograd node (cheb_cell.py) -----------------------------
    def _cheb_cell(self):
        from .cheb_cell import MAX_LAYERS
        return CHEB_CELL_FN and self.convolution_type in ('GCNConv', 'ChebConv') and self.n_conv_layers <= MAX_LAYERS

    def _cheb_packs(self, epoch):
        kind, S = self.convolution_type, self.n_conv_layers
        mode = "gcn" if kind == 'GCNConv' else "cheb"
        K = 1 if kind == 'GCNConv' else self.conv_x_i.convolutions[0].K
        pk = lambda key, convs: self._cached((mode, "cell") + key, epoch,
                                             lambda: _fused.shared_pack(pack_linear_group(convs(), kind)))
        packs = [pk(("x", 0), lambda: self._convs("x", 0)), pk(("h", 0), lambda: self._convs("h", 0))]
        packs += [pk(("all", l), lambda l=l: self._convs("x", l) + self._convs("h", l)) for l in range(1, S)]
        return mode, K, packs

    # ---- single-launch path (TransformerConv, hidden 32) ---------------------------------------
    def _fusable(self):
        return (_fused.ENABLED and self.convolution_type == 'TransformerConv' and self.out_channels == _fused.FC
                and self.in_channels <= 8)

    def _fused_cell(self, X, H, C, csr, params, norm_h, norm_c, norm_o, concat, want_head, eps, epoch):
        """The cell as one fused launch per conv layer (csrc/fused_fwd.inl): layer 0 reads X (4 convs) and H
        (4 convs); deeper layers read the previous layer's 8 blocks; the last layer ends in the gate epilogue."""
        S, Cw, Fin = self.n_conv_layers, self.out_channels, self.in_channels
        if Fin % 4:                  # the kernels take 16-byte rows: zero-pad X (the packed weights are zero there too)
            X = F.pad(X, (0, 4 - Fin % 4))
            Fin = X.shape[1]
        p = self.conv_x_i.convolutions[0].dropout if self.training else 0.0
        seed = (lambda: next_seed()) if p > 0 else (lambda: 0)
        dac = _fused.cap_of(Fin, True)
        wa = self._cached(("fa", 0), epoch, lambda: _fused.shared_pack(_fused.pack_fused_fn(self._convs("x", 0), dac)))
        wb = self._cached(("fb", 0), epoch, lambda: _fused.shared_pack(_fused.pack_fused_fn(self._convs("h", 0), _fused.FC)))
        flags = (bool(norm_h), bool(norm_c), bool(norm_o), bool(want_head), float(eps))

        def cfg(DA, GA, DB, GB, shared, mode):
            return (DA, GA, DB, GB, shared, mode, False, Cw) + flags + (float(p), seed())

        if S == 1:
            return _fused.FusedGroupFn.apply(X, wa, H, wb, C, params, concat, csr, cfg(Fin, 4, Cw, 4, True, 1))
        cur = _fused.FusedGroupFn.apply(X, wa, H, wb, None, None, None, csr, cfg(Fin, 4, Cw, 4, True, 0))
        for l in range(1, S):
            w = self._cached(("fb", l), epoch,
                             lambda: _fused.shared_pack(_fused.pack_fused_fn(self._convs("x", l) + self._convs("h", l), _fused.FC)))
            if l < S - 1:
                cur = _fused.FusedGroupFn.apply(None, None, cur, w, None, None, None, csr, cfg(0, 0, Cw, 8, False, 0))
            else:
                return _fused.FusedGroupFn.apply(None, None, cur, w, C, params, concat, csr, cfg(0, 0, Cw, 8, False, 1))

    def fused(self, X, edge_index, edge_weight=None, H=None, C=None, norm_h=None, norm_c=None, norm_o=None,
              concat=None, want_head=False, epoch=None):
        """Cell step plus the LayerNorms / head input the encoder and decoder apply to its outputs.
        Returns (O, H_out, C_out, head_in)."""
        epoch = new_epoch() if epoch is None else epoch
        N = X.shape[0]
        X = X.float()
        csr = get_csr(edge_index, edge_weight, N)
        if H is None:
            H = torch.zeros(N, self.out_channels, device=X.device)
        params = self._gate_params(epoch, norm_h, norm_c, norm_o)
        eps = norm_h.eps if norm_h is not None else 1e-5
        if self._fusable():
            return self._fused_cell(X, H, C, csr, params, norm_h is not None, norm_c is not None, norm_o is not None,
                                    concat, want_head, eps, epoch)
        if self._cheb_cell():
            mode, K, packs = self._cheb_packs(epoch)
            flags = (norm_h is not None, norm_c is not None, norm_o is not None, bool(want_head), float(eps))
            return ChebLstmCellFn.apply(X.contiguous(), H, C, params, concat, csr, mode, K, self.n_conv_layers, self.out_channels,
                                        flags, *packs)
        P = self._pre_activations(X, H, csr, epoch)
        return LstmGatesFn.apply(P, C, params, concat, norm_h is not None, norm_c is not None, norm_o is not None,
                                 want_head, eps)

    def forward(self, X, edge_index, edge_weight=None, H=None, C=None):
        """Returns (O, H', C') like the reference (model/model.py:430-463)."""
        O, Hn, Cn, _ = self.fused(X, edge_index, edge_weight, H, C)
        return O, Hn, Cn


class GConvGRU(nn.Module):
    r"""Graph-convolutional GRU cell, the reference's alternative to ``GConvLSTM`` (model/model.py:100-259): same constructor
    (no ``name`` argument, :133-139), same six ``conv_{x,h}_{z,r,h}`` stacks and state-dict keys, ``forward`` returns
    ``(H', H', None)`` and ignores ``C`` (:236-259).

    The six conv stacks run on the qmp kernels (``GraphConv`` -> ``convs.py``: CSR message passing + node GEMMs, forward and
    backward); the gate arithmetic is three element-wise expressions on their outputs.  ``conv_h_h`` reads ``H * R``, i.e. it
    depends on the reset gate, so unlike the LSTM cell the six stacks cannot all be batched into one group per layer:
    ``conv_x_*`` and ``conv_h_{z,r}`` first, ``conv_h_h`` after the reset gate."""

    def __init__(self, in_channels, out_channels, n_conv_layers=1, convolution_type='GCNConv'):
        super(GConvGRU, self).__init__()
        assert convolution_type in CONVOLUTIONS
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.n_conv_layers = n_conv_layers
        self.convolution_type = convolution_type
        for g in ("z", "r", "h"):      # creation order follows model/model.py:149-208 (same RNG stream for a seed)
            setattr(self, f"conv_x_{g}", GraphConv(convolution_type, in_channels, out_channels, n_conv_layers))
            setattr(self, f"conv_h_{g}", GraphConv(convolution_type, out_channels, out_channels, n_conv_layers))

    def forward(self, X, edge_index, edge_weight=None, H=None, C=None):
        """Returns (H', H', None) like the reference (model/model.py:236-259)."""
        X = X.float()
        if H is None:
            H = torch.zeros(X.shape[0], self.out_channels, device=X.device)
        from .ops import GruGates1Fn, GruGates2Fn
        H = H.float()
        # gate arithmetic as two element-wise launches around conv_h_h (csrc/lstm.cu: qmp_gru_gates1/2)
        Z, _, HR = GruGates1Fn.apply(self.conv_x_z(X, edge_index, edge_weight), self.conv_h_z(H, edge_index, edge_weight),
                                     self.conv_x_r(X, edge_index, edge_weight), self.conv_h_r(H, edge_index, edge_weight), H)
        Hn = GruGates2Fn.apply(self.conv_x_h(X, edge_index, edge_weight), self.conv_h_h(HR, edge_index, edge_weight), Z, H)
        return Hn, Hn, None


class MPNNLSTM(nn.Module):
    """Legacy model kept for API compatibility (model/model.py:613-684): three GCNConv blocks per frame,
    ``nn.LSTM`` over time, skip connection, two linears.  The graph convolutions run on the qmp kernels;
    the dense LSTM / linears are stock PyTorch modules as in the reference."""

    def __init__(self, hidden_size, dropout, input_timesteps=3, input_features=4, output_features=1):
        super(MPNNLSTM, self).__init__()
        self.dropout = dropout
        self.input_timesteps = input_timesteps
        self.convolution1 = GCNConv(input_features, hidden_size)
        self.convolution2 = GCNConv(hidden_size, hidden_size)
        self.convolution3 = GCNConv(hidden_size, hidden_size)
        self.bn1 = nn.LayerNorm(hidden_size)
        self.bn2 = nn.LayerNorm(hidden_size)
        self.bn3 = nn.LayerNorm(hidden_size)
        self.recurrents = nn.LSTM(hidden_size, hidden_size, 4)
        self.lin1 = nn.Linear(hidden_size + input_timesteps, hidden_size)
        self.lin2 = nn.Linear(hidden_size, output_features)

    def forward(self, X, edge_index, edge_weight=None):
        frames = []
        for i in range(X.shape[0]):
            H = X[i]
            for conv, norm in ((self.convolution1, self.bn1), (self.convolution2, self.bn2),
                               (self.convolution3, self.bn3)):
                H = F.relu(conv(H, edge_index, edge_weight))
                H = norm(H)
                H = F.dropout(H, p=self.dropout, training=self.training)
            frames.append(H)
        _, (H, _) = self.recurrents(torch.stack(frames))
        H = F.relu(H[-1])
        H = torch.cat([H, X[:, :, 0].T], dim=-1)
        H = self.lin2(F.relu(self.lin1(H)))
        H = F.dropout(H, p=self.dropout, training=self.training)
        return torch.sigmoid(H)


class MPNNLSTMI(nn.Module):
    """Legacy wrapper around a stack of GConvLSTM cells (model/model.py:727-802); API only."""

    def __init__(self, hidden_size, dropout, input_timesteps=3, input_features=4, n_layers=2, output_features=1):
        super(MPNNLSTMI, self).__init__()
        self.recurrents = nn.ModuleList([GConvLSTM(input_features, hidden_size)] +
                                        [GConvLSTM(hidden_size, hidden_size) for _ in range(n_layers - 1)])
        self.bn1 = nn.BatchNorm1d(hidden_size, track_running_stats=False)
        self.lin1 = nn.Linear(hidden_size, hidden_size)
        self.lin2 = nn.Linear(hidden_size, output_features)
        self.dropout = dropout
        self.input_timesteps = input_timesteps
        self.n_layers = n_layers

    def forward(self, X, edge_index, edge_weight=None):
        hs = [None] * self.n_layers
        cs = [None] * self.n_layers
        for x in X:
            _, h, c = self.recurrents[0](x, edge_index, edge_weight, H=hs[0], C=hs[1])   # sic: model.py:760
            hs[0], cs[0] = h, c
            for i in range(1, self.n_layers):
                _, h, c = self.recurrents[i](hs[i - 1], edge_index, edge_weight, H=hs[i], C=cs[i])
                hs[i], cs[i] = h, c
        x = F.relu(hs[-1])
        x = self.bn1(x)
        x = self.lin2(F.relu(self.lin1(x)))
        x = F.dropout(x, p=self.dropout, training=self.training)
        return torch.sigmoid(x)
