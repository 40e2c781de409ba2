"""Oracle restatement of the seq2seq driver (reference model/seq2seq.py).  TEST INFRASTRUCTURE.

Keeps the reference's observable behaviour, including its quirks:
* encoder layers >= 1 are called without state every timestep, and layer 0 is seeded from the
  TOP layer's previous state (seq2seq.py:59-71, 315-316);
* the decoder always uses one conv per stack (seq2seq.py:106);
* the decoder head needs ``concat_layers`` (fc_out1 has hidden+1 inputs, seq2seq.py:115-120);
* after a remesh ``graph.concat_layers`` is overwritten by the new graph's channel 0 and then
  replaced again at the next step by the pooled ``concat_layers[t]`` (seq2seq.py:363-368, 471, 484).
Pinned against the unmodified reference driver in ``tests/test_oracle_pinned.py``.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import graph_ref as G
from .cell_ref import CONVOLUTION_KWARGS, GConvLSTM
from .convs_ref import CONVOLUTIONS


class _State:
    """Counterpart of graph_functions.Graph (graph_functions.py:23-33)."""

    def __init__(self, edge_index, edge_attr):
        self.edge_index, self.edge_attr = edge_index, edge_attr
        self.x = None
        self.mapping = self.n_pixels_per_node = self.hidden = self.cell = None
        self.image_shape = None
        self.concat = None


class Encoder(nn.Module):
    def __init__(self, input_features, hidden_size, dropout, n_layers=1, convolution_type="GCNConv",
                 rnn_type="LSTM", n_conv_layers=3, dummy=False):
        super().__init__()
        assert rnn_type == "LSTM" and not dummy, "oracle covers the LSTM cell only (SURVEY.md section 2)"
        self.n_layers = n_layers
        dims = [input_features] + [hidden_size] * n_layers
        self.rnns = nn.ModuleList([GConvLSTM(dims[k], hidden_size, n_conv_layers, convolution_type, name="encoder")
                                   for k in range(n_layers)])
        self.dropout = nn.Dropout(dropout)            # constructed, never applied (seq2seq.py:47)
        self.norm_h = nn.LayerNorm(hidden_size)
        self.norm_c = nn.LayerNorm(hidden_size)

    def forward(self, X, edge_index, edge_weight, H=None, C=None):
        X = X.squeeze(0)
        hs, cs = [], []
        inp = X
        for k, rnn in enumerate(self.rnns):
            _, h, c = rnn(inp, edge_index, edge_weight, H=H if k == 0 else None, C=C if k == 0 else None)
            h, c = self.norm_h(h), self.norm_c(c)
            hs.append(h)
            cs.append(c)
            inp = h
        return torch.stack(hs), torch.stack(cs)


class Decoder(nn.Module):
    def __init__(self, input_features, hidden_size, dropout, n_layers=1, concat_layers_dim=3,
                 convolution_type="GCNConv", rnn_type="LSTM", n_conv_layers=3, binary=False, dummy=False):
        super().__init__()
        assert rnn_type == "LSTM" and not dummy
        self.n_layers, self.binary = n_layers, binary
        dims = [input_features] + [hidden_size] * n_layers
        self.rnns = nn.ModuleList([GConvLSTM(dims[k], hidden_size, 1, convolution_type, name="decoder")
                                   for k in range(n_layers)])   # n_conv_layers forced to 1 (seq2seq.py:106)
        make, kw = CONVOLUTIONS[convolution_type], CONVOLUTION_KWARGS[convolution_type]
        self.fc_out1 = make(in_channels=hidden_size + concat_layers_dim, out_channels=hidden_size, **kw)
        self.fc_out2 = make(in_channels=hidden_size, out_channels=1, **kw)
        self.norm_o = nn.LayerNorm(hidden_size)
        self.norm_h = nn.LayerNorm(hidden_size)
        self.norm_c = nn.LayerNorm(hidden_size)
        self.dropout = nn.Dropout(dropout)

    def forward(self, X, edge_index, edge_weight, concat_layers, H, C):
        hs, cs = [], []
        inp = X
        for k, rnn in enumerate(self.rnns):
            out, h, c = rnn(inp, edge_index, edge_weight, H=H[k], C=C[k])
            h, c = self.norm_h(h), self.norm_c(c)
            hs.append(h)
            cs.append(c)
            inp = h
        out = F.relu(self.norm_o(out))                                     # seq2seq.py:160-161
        if concat_layers is not None:
            out = torch.cat([out, concat_layers], dim=-1)
        out = self.fc_out1(out, edge_index, edge_weight)                   # gnn_out, seq2seq.py:182-187
        out = self.fc_out2(F.relu(out), edge_index, edge_weight)
        out = self.dropout(out)
        out = torch.tanh(out) + X[:, [0]]                                  # :171-174
        if self.binary:
            out = torch.sigmoid(out)
        return out, torch.stack(hs), torch.stack(cs)


class Seq2Seq(nn.Module):
    def __init__(self, hidden_size, dropout, thresh, input_timesteps=3, input_features=4, output_timesteps=5,
                 n_layers=4, n_conv_layers=2, transform_func=None, condition="max_larger_than",
                 remesh_input=False, convolution_type="ChebConv", rnn_type="LSTM", binary=False,
                 dummy=False, device=None, debug=False):
        super().__init__()
        assert not remesh_input, "remesh_input=True raises IndexError in the reference (SURVEY.md 0.9)"
        self.encoder = Encoder(input_features, hidden_size, dropout, n_layers, convolution_type, rnn_type,
                               n_conv_layers, dummy)
        self.decoder = Decoder(1 + 3, hidden_size, dropout, n_layers, 1, convolution_type, rnn_type,
                               n_conv_layers, binary, dummy)
        self.input_timesteps, self.output_timesteps = input_timesteps, output_timesteps
        self.n_layers, self.condition, self.thresh = n_layers, condition, thresh
        self.transform_func, self.convolution_type = transform_func, convolution_type
        self.use_edge_attrs = convolution_type in ("MHTransformerConv", "TransformerConv", "GATConv")  # :244
        self.graph = None

    # -- graph construction helper
    def _build(self, img, mask, hir):
        return G.image_to_graph(img, thresh=self.thresh, mask=mask, high_interest_region=hir,
                                transform_func=self.transform_func, condition=self.condition,
                                use_edge_attrs=self.use_edge_attrs)

    def process_inputs(self, x, mask=None, high_interest_region=None, graph_structure=None):
        """seq2seq.py:254-336."""
        image_shape = tuple(x.shape[1:3])
        self.mask = mask
        x = G.add_positional_encoding(x)
        if graph_structure is None:
            graph_structure = self._build(x, mask, high_interest_region)
        else:
            data = G.pool(x, graph_structure["mapping"], graph_structure["n_pixels_per_node"], mask)
            sizes = (graph_structure["n_pixels_per_node"] / ((4 / 2) ** 2)).repeat(x.shape[0], 1)   # :291
            graph_structure["data"] = torch.cat([data, sizes.unsqueeze(-1)], -1)
        g = _State(graph_structure["edge_index"], graph_structure["edge_attrs"])
        g.x = graph_structure["data"]
        g.mapping, g.n_pixels_per_node = graph_structure["mapping"], graph_structure["n_pixels_per_node"]
        g.image_shape = image_shape
        self.graph = g
        for t in range(self.input_timesteps):
            g.hidden, g.cell = self.encoder(g.x[[t]], g.edge_index, g.edge_attr,
                                            H=g.hidden[-1] if g.hidden is not None else None,
                                            C=g.cell[-1] if g.cell is not None else None)
        g.x = g.x[-1][:, [0, -3, -2, -1]]                                   # :336

    def unroll_output(self, unroll_steps, y, concat_layers=None, teacher_forcing_ratio=0.5, mask=None,
                      high_interest_region=None, remesh_every=1):
        """seq2seq.py:339-398."""
        g = self.graph
        outputs, mappings = [], []
        for t in unroll_steps:
            if concat_layers is not None:
                g.concat = G.pool(concat_layers[t].unsqueeze(0), g.mapping, g.n_pixels_per_node, self.mask).squeeze(0)
            out, hidden, cell = self.decoder(g.x, g.edge_index, g.edge_attr, g.concat, g.hidden, g.cell)
            outputs.append(out)
            mappings.append(g.mapping)
            teacher_force = random.random() < teacher_forcing_ratio
            teacher_input = y[[t]] if teacher_force else None
            if self.thresh != -np.inf and (t + 1) % remesh_every == 0:
                self._remesh(out, hidden, cell, mask, high_interest_region, teacher_force, teacher_input)
            else:
                if teacher_force:                                         # :421-424
                    ti = G.add_positional_encoding(teacher_input)
                    px = G.pool(ti, g.mapping, g.n_pixels_per_node, self.mask).squeeze(0)
                    g.x = torch.cat([px, g.n_pixels_per_node.unsqueeze(-1)], dim=-1)
                else:
                    g.x = torch.cat([out, g.x[..., 1:]], dim=-1)         # :427-428
                g.hidden, g.cell = hidden, cell
        return outputs, mappings

    def _remesh(self, data, hidden, cell, mask, hir, teacher_force, teacher_input):
        """seq2seq.py:434-491: nodes -> pixels with the old mesh, rebuild the mesh from the new
        frame, pixels -> nodes for the recurrent state."""
        g = self.graph
        shape = g.image_shape
        data_img = G.unpool(data, g.mapping, shape)
        hidden_img = G.unpool(hidden, g.mapping, shape)
        cell_img = G.unpool(cell, g.mapping, shape)
        if teacher_force:
            gs = self._build(G.add_positional_encoding(teacher_input), mask, hir)
        else:
            gs = self._build(G.add_positional_encoding(data_img.unsqueeze(0)), mask, hir)
        # flatten(swapaxes(img, 0, -1)) then swap back (:474-477) == pooling each [H, W] plane
        g.hidden = G.pool(hidden_img, gs["mapping"], gs["n_pixels_per_node"])
        g.cell = G.pool(cell_img, gs["mapping"], gs["n_pixels_per_node"])
        g.edge_index, g.edge_attr = gs["edge_index"], gs["edge_attrs"]
        g.x = gs["data"].squeeze(0)
        g.concat = gs["data"][:, :, [0]]
        g.mapping, g.n_pixels_per_node = gs["mapping"], gs["n_pixels_per_node"]

    def forward(self, x, y=None, concat_layers=None, teacher_forcing_ratio=0.5, mask=None,
                high_interest_region=None, graph_structure=None, remesh_every=1):
        self.process_inputs(x, mask=mask, high_interest_region=high_interest_region, graph_structure=graph_structure)
        return self.unroll_output(range(self.output_timesteps), y, concat_layers=concat_layers,
                                  teacher_forcing_ratio=teacher_forcing_ratio, mask=mask,
                                  high_interest_region=high_interest_region, remesh_every=remesh_every)
