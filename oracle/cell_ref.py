"""Oracle restatement of the recurrent cell (reference model/model.py).  TEST INFRASTRUCTURE.

``GraphConv`` = model/model.py:59-97 (a chain of convs with no nonlinearity between; the third
positional argument is the edge weight for GCN/Cheb and the edge attribute for Transformer).
``GConvLSTM`` = model/model.py:263-463.  ``GConvGRU`` = model/model.py:100-259.  ``MPNNLSTM`` = model/model.py:613-684 (legacy, API only).
Module / parameter names follow the reference so state dicts are interchangeable; creation
order follows model/model.py:294-373 so a shared seed gives identical initial weights.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .convs_ref import CONVOLUTIONS, GCNConv

CONVOLUTION_KWARGS = {                                  # model/model.py:49-57
    "GCNConv": dict(add_self_loops=False),
    "TransformerConv": dict(heads=1, edge_dim=2, dropout=0.1, concat=False),
    "ChebConv": dict(K=3, normalization="sym", bias=True),
    "MHTransformerConv": dict(heads=3, edge_dim=2, dropout=0.1),
    "GATConv": dict(heads=1, edge_dim=2),
    "GATv2Conv": dict(heads=1, edge_dim=2),
}

GATES = ("i", "f", "c", "o")


class GraphConv(nn.Module):
    def __init__(self, convolution_type, in_channels, out_channels, n_layers):
        super().__init__()
        self.convolution_type, self.n_layers = convolution_type, n_layers
        make, kw = CONVOLUTIONS[convolution_type], CONVOLUTION_KWARGS[convolution_type]
        dims = [in_channels] + [out_channels] * n_layers
        self.convolutions = nn.ModuleList([make(dims[k], dims[k + 1], **kw) for k in range(n_layers)])

    def forward(self, x, edge_index, edge_attr=None):
        for conv in self.convolutions:
            x = conv(x, edge_index, edge_attr)
        return x


class GConvLSTM(nn.Module):
    """Peephole graph-conv LSTM (model/model.py:394-463):

        I  = sigmoid(conv_x_i(X) + conv_h_i(H) + w_c_i * C  + b_i)
        F  = sigmoid(conv_x_f(X) + conv_h_f(H) + w_c_f * C  + b_f)
        T  = tanh   (conv_x_c(X) + conv_h_c(H)              + b_c)
        C' = F * C + I * T
        O  = sigmoid(conv_x_o(X) + conv_h_o(H) + w_c_o * C' + b_o)
        H' = O * tanh(C')                                   returns (O, H', C')
    """

    def __init__(self, in_channels, out_channels, n_conv_layers=1, convolution_type="GCNConv", name="GConvLSTM"):
        super().__init__()
        assert convolution_type in CONVOLUTIONS
        self.in_channels, self.out_channels = in_channels, out_channels
        self.n_conv_layers, self.convolution_type, self.name = n_conv_layers, convolution_type, name
        for g in GATES:                                  # creation order: model/model.py:294-373
            setattr(self, f"conv_x_{g}", GraphConv(convolution_type, in_channels, out_channels, n_conv_layers))
            setattr(self, f"conv_h_{g}", GraphConv(convolution_type, out_channels, out_channels, n_conv_layers))
            if g != "c":
                setattr(self, f"w_c_{g}", nn.Parameter(torch.zeros(1, out_channels)))   # :375-382 zero init
            setattr(self, f"b_{g}", nn.Parameter(torch.zeros(1, out_channels)))

    def _pre(self, g, X, ei, ew, H):
        return getattr(self, f"conv_x_{g}")(X, ei, ew) + getattr(self, f"conv_h_{g}")(H, ei, ew)

    def forward(self, X, edge_index, edge_weight=None, H=None, C=None):
        n = X.shape[0]
        H = torch.zeros(n, self.out_channels) if H is None else H
        C = torch.zeros(n, self.out_channels) if C is None else C
        I = torch.sigmoid(self._pre("i", X, edge_index, edge_weight, H) + self.w_c_i * C + self.b_i)
        Fg = torch.sigmoid(self._pre("f", X, edge_index, edge_weight, H) + self.w_c_f * C + self.b_f)
        T = torch.tanh(self._pre("c", X, edge_index, edge_weight, H) + self.b_c)
        C = Fg * C + I * T
        O = torch.sigmoid(self._pre("o", X, edge_index, edge_weight, H) + self.w_c_o * C + self.b_o)
        return O, O * torch.tanh(C), C


class GConvGRU(nn.Module):
    """Graph-convolutional GRU cell (model/model.py:100-259):

        Z  = sigmoid(conv_x_z(X) + conv_h_z(H))                 update gate        (:215-219)
        R  = sigmoid(conv_x_r(X) + conv_h_r(H))                 reset gate         (:221-225)
        H~ = tanh   (conv_x_h(X) + conv_h_h(H * R))             candidate state    (:227-231)
        H' = Z * H + (1 - Z) * H~                               returns (H', H', None)   (:233-259)

    No biases or peepholes of its own; ``C`` is accepted and ignored (LSTM compatibility, :243).  The constructor takes no
    ``name`` argument (:133-139), which is why ``Seq2Seq(rnn_type='GRU')`` fails in the reference (seq2seq.py:43 passes one).
    Creation order z, r, h with x before h (:149-208)."""

    def __init__(self, in_channels, out_channels, n_conv_layers=1, convolution_type="GCNConv"):
        super().__init__()
        assert convolution_type in CONVOLUTIONS
        self.in_channels, self.out_channels = in_channels, out_channels
        self.n_conv_layers, self.convolution_type = n_conv_layers, convolution_type
        for g in ("z", "r", "h"):
            setattr(self, f"conv_x_{g}", GraphConv(convolution_type, in_channels, out_channels, n_conv_layers))
            setattr(self, f"conv_h_{g}", GraphConv(convolution_type, out_channels, out_channels, n_conv_layers))

    def forward(self, X, edge_index, edge_weight=None, H=None, C=None):
        H = torch.zeros(X.shape[0], self.out_channels) if H is None else H
        Z = torch.sigmoid(self.conv_x_z(X, edge_index, edge_weight) + self.conv_h_z(H, edge_index, edge_weight))
        R = torch.sigmoid(self.conv_x_r(X, edge_index, edge_weight) + self.conv_h_r(H, edge_index, edge_weight))
        Ht = torch.tanh(self.conv_x_h(X, edge_index, edge_weight) + self.conv_h_h(H * R, edge_index, edge_weight))
        H = Z * H + (1 - Z) * Ht
        return H, H, None


class MPNNLSTM(nn.Module):
    """Legacy model (model/model.py:613-684): 3 x (GCNConv -> relu -> LayerNorm -> dropout) per
    frame, nn.LSTM(C, C, 4) over time, skip X[:, :, 0].T, two linears, dropout, sigmoid."""

    def __init__(self, hidden_size, dropout, input_timesteps=3, input_features=4, output_features=1):
        super().__init__()
        self.dropout, self.input_timesteps = dropout, input_timesteps
        self.convolution1 = GCNConv(input_features, hidden_size)
        self.convolution2 = GCNConv(hidden_size, hidden_size)
        self.convolution3 = GCNConv(hidden_size, hidden_size)
        self.bn1, self.bn2, self.bn3 = nn.LayerNorm(hidden_size), nn.LayerNorm(hidden_size), nn.LayerNorm(hidden_size)
        self.recurrents = nn.LSTM(hidden_size, hidden_size, 4)
        self.lin1 = nn.Linear(hidden_size + input_timesteps, hidden_size)
        self.lin2 = nn.Linear(hidden_size, output_features)

    def forward(self, X, edge_index, edge_weight=None):
        frames = []
        for t in range(X.shape[0]):
            h = X[t]
            for conv, norm in ((self.convolution1, self.bn1), (self.convolution2, self.bn2),
                               (self.convolution3, self.bn3)):
                h = F.dropout(norm(F.relu(conv(h, edge_index, edge_weight))), p=self.dropout,
                              training=self.training)
            frames.append(h)
        _, (hn, _) = self.recurrents(torch.stack(frames))
        h = torch.cat([F.relu(hn[-1]), X[:, :, 0].T], dim=-1)
        h = self.lin2(F.relu(self.lin1(h)))
        return torch.sigmoid(F.dropout(h, p=self.dropout, training=self.training))
