"""B200 mirror of the reference's ``model/seq2seq.py``: ``Encoder``, ``Decoder``, ``Seq2Seq`` with the
same constructor arguments, ``forward(x, y=None, concat_layers=None, teacher_forcing_ratio=0.5, mask=None,
high_interest_region=None, graph_structure=None, remesh_every=1)`` returning
``(outputs: list of [N_t, 1], output_mappings: list)``, the same ``process_inputs`` / ``unroll_output``
entry points the trainer calls (model/mpnnlstm.py:292-303) and the same state-dict keys.

Behaviour kept from the reference: encoder layers >= 1 run without state and layer 0
is seeded from the top layer's state; the decoder always uses one conv per stack; the decoder head needs
``concat_layers``; ``remesh_input=True`` raises like the reference does (IndexError).
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import fused as _fused
from .graph_csr import get_csr
from .graph_functions import Graph, Mesh, flatten, image_to_graph, regrid, unflatten
from .model import CONVOLUTION_KWARGS, CONVOLUTIONS, GConvLSTM, new_epoch
from .ops import HeadFinishFn, NodeLinearFn, SpmmFn, TConvFn, next_seed
from .convs import cheb_basis, pack_tconv
from .cheb_cell import ChebStackFn, pack_linear_group
from . import model as _model

# ChebConv / GCNConv decoder head as one autograd node (cheb_cell.ChebStackFn).  OFF by default -- measured: no host time gained
# on the configs[0]-like sample (32.5 vs 32.3 ms) and its two ~2 MB workspaces per forecast step fragment the caching allocator's
# 20 MB segments next to the cell's ~9 MB workspaces (51 cudaMalloc per sample, 32 -> 41 ... 290 ms without expandable segments).
CHEB_STACK_FN = os.environ.get("QMP_CHEB_STACK_FN", "0") == "1"
from .utils import add_positional_encoding


def _rnn_class(rnn_type):
    assert rnn_type in ['GRU', 'LSTM', 'SplitLSTM']
    if rnn_type == 'GRU':
        # the reference cannot build this either: Encoder / Decoder pass name=... (seq2seq.py:43-44, 108-109) to
        # GConvGRU.__init__, which has no such argument (model/model.py:133-139) -> TypeError.  The cell itself is
        # available stand-alone as model.GConvGRU.
        raise TypeError("GConvGRU.__init__() got an unexpected keyword argument 'name' (rnn_type='GRU' fails the same way "
                        "in the reference driver; use model.GConvGRU directly)")
    if rnn_type != 'LSTM':
        raise NotImplementedError(f"rnn_type={rnn_type!r}: only the LSTM cell is on the hot path (every reference "
                                  "config uses rnn_type='LSTM'; SURVEY.md section 2)")
    return GConvLSTM


def _stack_layers(states):
    """[L, N, C] from the per-layer states.  One layer (every ice configuration): a view -- torch.stack would copy the state
    every timestep, and its backward plus the select below would add a zero fill and two more copies per timestep."""
    return states[0].unsqueeze(0) if len(states) == 1 else torch.stack(states)


def _layer_states(S, n_layers):
    """Per-layer [N, C] states of a stacked [L, N, C] tensor (views; no select_backward zero fills for one layer)."""
    if n_layers == 1 and S.dim() == 3 and S.shape[0] == 1:
        return (S.view(S.shape[1], S.shape[2]),) if S.is_contiguous() else (S[0],)
    return [S[i] for i in range(n_layers)]


class Encoder(torch.nn.Module):
    def __init__(self, input_features, hidden_size, dropout, n_layers=1, convolution_type='GCNConv', rnn_type='LSTM',
                 n_conv_layers=3, dummy=False):
        super().__init__()
        rnn = _rnn_class(rnn_type)
        if dummy:
            raise NotImplementedError("dummy=True is a debugging switch of the reference; not implemented")
        self.rnn_type, self.hidden_size, self.n_layers, self.dummy = rnn_type, hidden_size, n_layers, dummy
        self.rnns = nn.ModuleList(
            [rnn(input_features, hidden_size, convolution_type=convolution_type, n_conv_layers=n_conv_layers, name='encoder')] +
            [rnn(hidden_size, hidden_size, convolution_type=convolution_type, n_conv_layers=n_conv_layers, name='encoder')
             for _ in range(n_layers - 1)])
        self.dropout = nn.Dropout(dropout)       # constructed but never applied, as in the reference (seq2seq.py:47)
        self.norm_h = nn.LayerNorm(hidden_size)
        self.norm_c = nn.LayerNorm(hidden_size)

    def forward(self, X, edge_index, edge_weight, H=None, C=None, _epoch=None):
        """One encoder timestep (seq2seq.py:52-82): layer 0 takes (H, C); layers >= 1 start from zeros."""
        epoch = new_epoch() if _epoch is None else _epoch
        if X.dim() == 3:
            X = X.squeeze(0)
        hidden, cell = [], []
        inp = X
        for i in range(self.n_layers):
            _, h, c, _ = self.rnns[i].fused(inp, edge_index, edge_weight, H=H if i == 0 else None,
                                            C=C if i == 0 else None, norm_h=self.norm_h, norm_c=self.norm_c,
                                            epoch=epoch)
            hidden.append(h)
            cell.append(c)
            inp = h
        return _stack_layers(hidden), _stack_layers(cell)


class Decoder(torch.nn.Module):
    def __init__(self, input_features, hidden_size, dropout, n_layers=1, concat_layers_dim=3, convolution_type='GCNConv',
                 rnn_type='LSTM', n_conv_layers=3, binary=False, dummy=False):
        super().__init__()
        rnn = _rnn_class(rnn_type)
        if dummy:
            raise NotImplementedError("dummy=True is a debugging switch of the reference; not implemented")
        self.rnn_type, self.input_features, self.hidden_size = rnn_type, input_features, hidden_size
        self.n_layers, self.binary, self.dummy = n_layers, binary, dummy
        self.convolution_type = convolution_type
        n_conv_layers = 1  # hard-coded single convolutional layer in the decoder (seq2seq.py:106)
        self.rnns = nn.ModuleList(
            [rnn(input_features, hidden_size, convolution_type=convolution_type, n_conv_layers=n_conv_layers, name='decoder')] +
            [rnn(hidden_size, hidden_size, convolution_type=convolution_type, n_conv_layers=n_conv_layers, name='decoder')
             for _ in range(n_layers - 1)])
        in_channels = hidden_size + concat_layers_dim
        conv_func = CONVOLUTIONS[convolution_type]
        conv_func_kwargs = CONVOLUTION_KWARGS[convolution_type]
        self.fc_out1 = conv_func(in_channels=in_channels, out_channels=hidden_size, **conv_func_kwargs)
        self.fc_out2 = conv_func(in_channels=hidden_size, out_channels=1, **conv_func_kwargs)
        self.norm_o = nn.LayerNorm(hidden_size)
        self.norm_h = nn.LayerNorm(hidden_size)
        self.norm_c = nn.LayerNorm(hidden_size)
        self.dropout = nn.Dropout(dropout)
        self._cache = {}

    def _cached(self, key, epoch, build):
        hit = self._cache.get(key)
        if hit is not None and hit[0] == epoch:
            return hit[1]
        val = build()
        self._cache[key] = (epoch, val)
        return val

    def forward(self, X, edge_index, edge_weight, concat_layers, H, C, _epoch=None, _want_next=False):
        """One decoder timestep (seq2seq.py:129-180).  Returns (output [N, 1], hidden [L, N, C], cell [L, N, C])."""
        epoch = new_epoch() if _epoch is None else _epoch
        if concat_layers is None:
            raise RuntimeError("Decoder needs concat_layers: fc_out1 is built with hidden_size + concat_layers_dim "
                               "inputs (the reference fails with a shape error here, seq2seq.py:115-120)")
        X = X.float()
        N = X.shape[0]
        csr = get_csr(edge_index, edge_weight, N)
        hidden, cell = [], []
        inp, head = X, None
        Hs, Cs = _layer_states(H, self.n_layers), _layer_states(C, self.n_layers)
        for i in range(self.n_layers):
            top = i == self.n_layers - 1
            _, h, c, head = self.rnns[i].fused(inp, csr, None, H=Hs[i], C=Cs[i], norm_h=self.norm_h, norm_c=self.norm_c,
                                               norm_o=self.norm_o if top else None,
                                               concat=concat_layers if top else None, want_head=top, epoch=epoch)
            hidden.append(h)
            cell.append(c)
            inp = h
        p = self.dropout.p if self.training else 0.0
        tail = self._head_tail(head, X, csr, epoch, p)      # fc_out1 -> relu -> fc_out2 -> tanh + residual, fewest launches
        if tail is not None:
            out, x_next = tail
        else:
            y = self._gnn_out(head, csr, epoch)             # fc_out1 -> relu -> fc_out2 (seq2seq.py:182-187)
            out, x_next = HeadFinishFn.apply(y, X, self.binary, p, next_seed() if p > 0 else 0)
        hidden, cell = _stack_layers(hidden), _stack_layers(cell)
        if _want_next:
            return out, hidden, cell, x_next
        return out, hidden, cell

    def _head_tail(self, head, X, csr, epoch, p_out):
        """(out, x_next) through FusedGroupFn (fc_out1, relu) + fused.HeadTailFn (fc_out2 and the tanh / residual tail in one
        autograd node), or None when that path does not apply (then _gnn_out + HeadFinishFn run)."""
        if not (self.convolution_type == 'TransformerConv' and _fused.ENABLED and _fused.SCALAR_HEAD and _fused.HEAD_TAIL
                and self.hidden_size == _fused.FC and head.shape[1] == _fused.HEADW and self.fc_out2.out_channels == 1):
            return None
        p = self.fc_out1.dropout if self.training else 0.0
        sd = (lambda: next_seed()) if p > 0 else (lambda: 0)
        w1 = self._cached("ffc1", epoch, lambda: _fused.shared_pack(_fused.pack_fused_fn([self.fc_out1], _fused.HEADW)))
        tail = (False, False, False, False, 1e-5, float(p))
        h1 = _fused.FusedGroupFn.apply(None, None, head, w1, None, None, None, csr,
                                       (0, 0, _fused.HEADW, 1, True, 0, 2, _fused.FC) + tail + (sd(),))     # relu_out = 2: see HeadTailFn
        P2 = self._cached("ftc1", epoch, lambda: _fused.shared_pack(_fused.pack_tconv1(self.fc_out2)))
        return _fused.HeadTailFn.apply(h1, P2, X, csr, float(p), sd(), self.binary, float(p_out), next_seed() if p_out > 0 else 0, True)

    def _gnn_out(self, head, csr, epoch):
        kind = self.convolution_type
        if kind == 'TransformerConv':
            p = self.fc_out1.dropout if self.training else 0.0
            sd = (lambda: next_seed()) if p > 0 else (lambda: 0)
            if _fused.ENABLED and self.hidden_size == _fused.FC and head.shape[1] == _fused.HEADW:
                # two fused launches (csrc/fused_fwd.inl): fc_out1 on the 36-wide padded head rows, relu; fc_out2 -> 1
                w1 = self._cached("ffc1", epoch, lambda: _fused.shared_pack(_fused.pack_fused_fn([self.fc_out1], _fused.HEADW)))
                tail = (False, False, False, False, 1e-5, float(p))
                h1 = _fused.FusedGroupFn.apply(None, None, head, w1, None, None, None, csr,
                                               (0, 0, _fused.HEADW, 1, True, 0, True, _fused.FC) + tail + (sd(),))
                if _fused.SCALAR_HEAD and self.fc_out2.out_channels == 1:
                    # one output channel: scalar query / key / value records instead of 32-wide rows (csrc/tconv1.cu)
                    P2 = self._cached("ftc1", epoch, lambda: _fused.shared_pack(_fused.pack_tconv1(self.fc_out2)))
                    return _fused.ScalarTConvFn.apply(h1, P2, csr, float(p), sd())
                w2 = self._cached("ffc2", epoch, lambda: _fused.shared_pack(_fused.pack_fused_fn([self.fc_out2], _fused.FC)))
                return _fused.FusedGroupFn.apply(None, None, h1, w2, None, None, None, csr,
                                                 (0, 0, _fused.FC, 1, True, 0, False, 1) + tail + (sd(),))
            pk1 = self._cached("fc1", epoch, lambda: pack_tconv([self.fc_out1]))
            pk2 = self._cached("fc2", epoch, lambda: pack_tconv([self.fc_out2]))
            h1 = TConvFn.apply(head, *pk1, csr, True, p, sd(), True, None)
            return TConvFn.apply(h1, *pk2, csr, True, p, sd(), False, None)
        if kind in ('GCNConv', 'ChebConv') and _model.CHEB_CELL_FN and CHEB_STACK_FN:
            # fc_out2(relu(fc_out1(.))) as one autograd node and one library call each way (cheb_cell.ChebStackFn)
            mode = "gcn" if kind == 'GCNConv' else "cheb"
            K = 1 if kind == 'GCNConv' else self.fc_out1.K
            pk = lambda key, conv: self._cached(key, epoch, lambda: _fused.shared_pack(pack_linear_group([conv], kind)))
            return ChebStackFn.apply(head, csr, mode, K, (True, False), pk("cs1", self.fc_out1), pk("cs2", self.fc_out2))
        if kind in ('GCNConv', 'ChebConv'):
            def lin(conv, z, relu):
                if kind == 'GCNConv':
                    t = SpmmFn.apply(z, None, csr, "gcn", 1.0, 0.0)
                    W = conv.lin.weight.unsqueeze(0)
                else:
                    t = torch.cat(cheb_basis(z, csr, conv.K), dim=1)
                    W = torch.cat([l.weight for l in conv.lins], dim=1).unsqueeze(0)
                o = NodeLinearFn.apply(t, W, conv.bias.unsqueeze(0), True)
                return torch.relu(o) if relu else o
            return lin(self.fc_out2, lin(self.fc_out1, head, True), False)
        # any other conv type: two module calls (seq2seq.py:182-187)
        ei, ea = csr._keepalive
        return self.fc_out2(torch.relu(self.fc_out1(head, ei, ea)), ei, ea)

    def gnn_out(self, x, edge_index, edge_weight):
        """Reference-shaped helper (seq2seq.py:182-187)."""
        csr = get_csr(edge_index, edge_weight, x.shape[0])
        y = self._gnn_out(x.float(), csr, new_epoch())
        return self.dropout(y)


class Seq2Seq(torch.nn.Module):
    def __init__(self,
                 hidden_size,
                 dropout,
                 thresh,
                 input_timesteps=3,
                 input_features=4,
                 output_timesteps=5,
                 n_layers=4,
                 n_conv_layers=2,
                 transform_func=None,
                 condition='max_larger_than',
                 remesh_input=False,
                 convolution_type='ChebConv',
                 rnn_type='LSTM',
                 binary=False,
                 dummy=False,
                 device=None,
                 debug=False):
        super().__init__()
        self.encoder = Encoder(input_features, hidden_size, dropout, n_layers=n_layers,
                               convolution_type=convolution_type, rnn_type=rnn_type, n_conv_layers=n_conv_layers,
                               dummy=dummy)
        self.decoder = Decoder(1 + 3, hidden_size, dropout, n_layers=n_layers, concat_layers_dim=1,
                               convolution_type=convolution_type, rnn_type=rnn_type, n_conv_layers=n_conv_layers,
                               binary=binary, dummy=dummy)
        self.input_timesteps = input_timesteps
        self.output_timesteps = output_timesteps
        self.n_layers = n_layers
        self.condition = condition
        self.remesh_input = remesh_input
        self.debug = debug
        self.convolution_type = convolution_type
        # These convolutions can accept edge attributes, the others cannot (seq2seq.py:244).
        self.use_edge_attrs = convolution_type in ['MHTransformerConv', 'TransformerConv', 'GATConv']
        self.thresh = thresh
        self.transform_func = transform_func
        self.graph = None
        self.device = device
        self._epoch = 0

    # -- helpers -----------------------------------------------------------------------------------
    def _image_to_graph(self, img, mask, high_interest_region):
        gs = image_to_graph(img, thresh=self.thresh, mask=mask, high_interest_region=high_interest_region,
                            transform_func=self.transform_func, condition=self.condition,
                            use_edge_attrs=self.use_edge_attrs)
        # register the CSR now (no validation read-back: the kernels produced these indices)
        get_csr(gs['edge_index'], gs['edge_attrs'], gs['data'].shape[1], validate=False)
        return gs

    def process_inputs(self, x, mask=None, high_interest_region=None, graph_structure=None):
        """Build (or adopt) the mesh, pool the input frames onto it and run the encoder (seq2seq.py:254-336)."""
        if not _lib.on_device(x):
            raise _lib.QmpError("Seq2Seq: CUDA tensors required (no CPU fallback)")
        num_samples, w, h, c = x.shape
        image_shape = (w, h)
        self.mask = mask
        self._epoch = new_epoch()
        if self.remesh_input:
            # reference: do_remesh_input(x[[t+1]]) runs off the end of x on the last encoder step (seq2seq.py:321-324)
            raise IndexError("remesh_input=True indexes x[[input_timesteps]] in the reference and fails; unsupported")
        x = add_positional_encoding(x)
        if graph_structure is None:
            graph_structure = self._image_to_graph(x, mask, high_interest_region)
        else:
            data = flatten(x, graph_structure['mapping'], graph_structure['n_pixels_per_node'], mask)
            node_sizes = graph_structure['n_pixels_per_node'].to(data.device).float() / ((4 / 2) ** 2)  # seq2seq.py:291
            node_sizes = node_sizes.reshape(1, -1, 1).expand(data.shape[0], -1, 1)
            graph_structure['data'] = torch.cat([data, node_sizes], -1)

        self.graph = Graph(graph_structure['edge_index'], graph_structure['edge_attrs'])
        self.graph.pyg.x = graph_structure['data']
        self.graph.mapping = graph_structure['mapping']
        self.graph.n_pixels_per_node = graph_structure['n_pixels_per_node']
        self.graph.image_shape = image_shape

        self.graph.hidden, self.graph.cell = None, None
        for t in range(self.input_timesteps):
            hidden, cell = self.encoder(
                X=self.graph.pyg.x[t],
                edge_index=self.graph.pyg.edge_index,
                edge_weight=self.graph.pyg.edge_attr,
                H=_layer_states(self.graph.hidden, self.encoder.n_layers)[-1] if self.graph.hidden is not None else None,
                C=_layer_states(self.graph.cell, self.encoder.n_layers)[-1] if self.graph.cell is not None else None,
                _epoch=self._epoch)
            self.graph.hidden = hidden
            self.graph.cell = cell

        # first decoder input = last encoder input, [value, ii, jj, node size] (seq2seq.py:336)
        last = self.graph.pyg.x[-1]
        self.graph.pyg.x = torch.cat([last[:, :1], last[:, -3:]], dim=1)

    def unroll_output(self, unroll_steps, y, concat_layers=None, teacher_forcing_ratio=0.5, mask=None,
                      high_interest_region=None, remesh_every=1):
        """Decoder rollout (seq2seq.py:339-398)."""
        self._unpooled = {}
        outputs = []
        output_mappings = []
        g = self.graph
        pooled_concat = None
        if concat_layers is not None and self.thresh == -np.inf:
            # static mesh: pool every forecast step's concat layer in one launch instead of one per step
            pooled_concat = flatten(concat_layers.float(), g.mapping, g.n_pixels_per_node, self.mask)
        for t in unroll_steps:
            if concat_layers is not None:
                if pooled_concat is not None:
                    g.concat_layers = pooled_concat[t]
                else:
                    g.concat_layers = flatten(concat_layers[t].unsqueeze(0).float(), g.mapping, g.n_pixels_per_node,
                                              self.mask).squeeze(0)
            output, hidden, cell, x_next = self.decoder(
                X=g.pyg.x, edge_index=g.pyg.edge_index, edge_weight=g.pyg.edge_attr,
                concat_layers=getattr(g, 'concat_layers', None), H=g.hidden, C=g.cell,
                _epoch=self._epoch, _want_next=True)
            outputs.append(output)
            output_mappings.append(g.mapping)

            teacher_force = random.random() < teacher_forcing_ratio
            teacher_input = y[[t]] if teacher_force else None

            if (self.thresh != -np.inf) and ((t + 1) % remesh_every == 0):
                self.do_remesh(output, hidden, cell, mask, high_interest_region, teacher_force=teacher_force,
                               teacher_input=teacher_input)
            else:
                self.update_without_remesh(output, hidden, cell, teacher_force=teacher_force,
                                           teacher_input=teacher_input, _x_next=x_next)
        return outputs, output_mappings

    def forward(self, x, y=None, concat_layers=None, teacher_forcing_ratio=0.5, mask=None, high_interest_region=None,
                graph_structure=None, remesh_every=1):
        self.process_inputs(x, mask=mask, high_interest_region=high_interest_region, graph_structure=graph_structure)
        return self.unroll_output(range(self.output_timesteps), y, concat_layers=concat_layers,
                                  teacher_forcing_ratio=teacher_forcing_ratio, mask=mask,
                                  high_interest_region=high_interest_region, remesh_every=remesh_every)

    def update_without_remesh(self, data, hidden, cell, teacher_force=False, teacher_input=None, _x_next=None):
        """seq2seq.py:420-431."""
        g = self.graph
        if teacher_force:
            teacher_input = add_positional_encoding(teacher_input.float())
            x = flatten(teacher_input, g.mapping, g.n_pixels_per_node, self.mask).squeeze(0)
            g.pyg.x = torch.cat([x, g.n_pixels_per_node.unsqueeze(-1)], dim=-1)   # raw pixel counts, as in the reference
        elif _x_next is not None:
            g.pyg.x = _x_next                                                     # [output, previous positional columns]
        else:
            g.pyg.x = torch.cat([data, g.pyg.x[..., 1:]], dim=-1)
        g.hidden = hidden
        g.cell = cell

    def do_remesh(self, data, hidden, cell, mask=None, high_interest_region=None, teacher_force=False, teacher_input=None):
        """Regrid onto a mesh rebuilt from the new frame (seq2seq.py:434-491): nodes -> pixels with the old
        mesh, quadtree on the new frame, pixels -> nodes for the recurrent state."""
        g = self.graph
        image_shape = g.image_shape
        data_img = unflatten(data, g.mapping, image_shape)
        # (the loss of a dynamic-mesh training step unpools this very output on this very mesh: train.TrainStep reuses the image)
        if getattr(self, "_unpooled", None) is not None:
            self._unpooled[id(data)] = (data, g.mapping, data_img)
        if teacher_force:
            gs = self._image_to_graph(add_positional_encoding(teacher_input.float()), mask, high_interest_region)
        else:
            gs = self._image_to_graph(add_positional_encoding(data_img.unsqueeze(0)), mask, high_interest_region)
        # flatten(swapaxes(img, 0, -1)) and swap back (seq2seq.py:474-477) == pooling every [H, W] plane
        # -- hidden and cell state go old nodes -> pixels -> new nodes in one launch, no [H, W, C] images (graph_functions.regrid)
        g.hidden, g.cell = regrid(g.mapping, gs['mapping'], hidden, cell, image_shape, gs['n_pixels_per_node'])
        g.pyg.edge_index = gs['edge_index']
        g.pyg.edge_attr = gs['edge_attrs']
        g.pyg.x = gs['data'].squeeze(0)
        g.concat_layers = gs['data'][:, :, :1]             # (a view: indexing with [0] costs an index tensor, a host-to-device copy and a gather)
        g.mapping = gs['mapping']
        g.n_pixels_per_node = gs['n_pixels_per_node']
        g.image_shape = image_shape
