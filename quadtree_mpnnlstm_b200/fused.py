"""Host side of the fused cell kernels (csrc/fused_fwd.cu, csrc/fused_bwd.cu): weight packing and the
autograd Function.  Used by GConvLSTM / Decoder when the configuration fits the fused path
(TransformerConv, hidden size 32, inputs <= 8 wide for the X stacks and <= 36 wide otherwise); anything
else runs the modular kernels (ops.py), which compute the same thing.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from . import _lib
from .convs import pack_tconv

FC = 32
HEADW = 36          # decoder-head input rows: 32 normalised outputs | concat layer | 3 zero pad columns (16-byte rows)
ENABLED = True      # tests flip this to cross-check the fused kernels against the modular ones
TC_WGRAD = True     # weight gradients: one tcgen05 launch per group (False: ten FFMA reductions, the cross-check)
TC_BWD = True       # backward (target / source side) on tcgen05 (csrc/fused_bwd_tc.inl); False: fp32-FFMA kernels
TC_FWD = True       # forward on tcgen05 (csrc/fused_fwd_tc.inl); False: the fp32-FFMA kernel (csrc/fused_fwd.inl)
SCALAR_HEAD = True  # decoder head's fc_out2 (hidden -> 1 channel): scalar query / key / value kernels (csrc/tconv1.cu)
ONEPASS_BWD = True  # tcgen05 backward of the other groups: source side of every edge by vector reductions inside the target kernel (one launch)
CELL_BWD = True     # ... and its backward: target + source side of every edge in one persistent launch (csrc/fused_cell_bwd.cu)
CELL_BWD_GATES = os.environ.get("QMP_CELL_BWD_GATES", "1") != "0"     # gate backward inside the decoder-cell backward kernel (False: qmp_lstm_gates_bwd launch first)
PANEL_WGRAD = os.environ.get("QMP_PANEL_WGRAD", "1") != "0"   # weight gradients of the per-conv groups: the streaming TMA -> tcgen05 kernel (csrc/panel_wgrad.cu)
CELL_FWD = True     # decoder cell (4 X convs + 4 H convs, gate mode): the persistent gates-batched kernel (csrc/fused_cell_fwd.cu)
_f32 = torch.float32


def cap_of(D, small):
    if small:
        return 4 if D <= 4 else 8
    return 32 if D <= 32 else 36


def conv_total(DC):
    return (DC + 2) * DC + (DC + 4) + FC * (DC + 4) + FC * DC + FC


def pack_fused(convs, DC):
    """[G, TOTAL] padded pack (layout in csrc/fused.cuh) from PyG-named TransformerConv parameters."""
    W1, b1, W2, W3, b3 = pack_tconv(convs)                      # unpadded, see convs.pack_tconv
    G, C, D = W3.shape
    assert C <= FC and D <= DC
    W1p = torch.cat([F.pad(W1[:, :D], (0, DC - D, 0, DC - D)), F.pad(W1[:, D:], (0, DC - D))], dim=1)   # [G, DC+2, DC]
    b1p = torch.cat([F.pad(b1[:, :D], (0, DC - D)), b1[:, D:], b1.new_zeros(G, 2)], dim=1)            # [G, DC+4]
    W2p = torch.cat([F.pad(W2[:, :, :D], (0, DC - D)), W2[:, :, D:], W2.new_zeros(G, C, 1)], dim=2)    # [G, C, DC+4]
    W2p = F.pad(W2p, (0, 0, 0, FC - C))
    W3p = F.pad(W3, (0, DC - D, 0, FC - C))
    b3p = F.pad(b3, (0, FC - C))
    return torch.cat([W1p.flatten(1), b1p, W2p.flatten(1), W3p.flatten(1), b3p], dim=1).contiguous()


_PARAMS_PER_CONV = 9      # lin_query.{weight,bias}, lin_key.{weight,bias}, lin_value.{weight,bias}, lin_edge.weight, lin_skip.{weight,bias}


def _conv_params(c):
    return (c.lin_query.weight, c.lin_query.bias, c.lin_key.weight, c.lin_key.bias, c.lin_value.weight, c.lin_value.bias,
            c.lin_edge.weight, c.lin_skip.weight, c.lin_skip.bias)


class _ParamView:
    """The attribute shape ``pack_tconv`` reads (``conv.lin_query.weight`` ...) over plain tensors."""

    class _Lin:
        __slots__ = ("weight", "bias")

        def __init__(self, weight, bias):
            self.weight, self.bias = weight, bias

    def __init__(self, ps, out_channels):
        self.out_channels = out_channels
        self.lin_query, self.lin_key = self._Lin(ps[0], ps[1]), self._Lin(ps[2], ps[3])
        self.lin_value, self.lin_edge, self.lin_skip = self._Lin(ps[4], ps[5]), self._Lin(ps[6], None), self._Lin(ps[7], ps[8])


_ptr_tables = {}     # parameter data_ptrs of a conv group -> (device int64 table [G, 9], the parameters): built once per group (the
                     # optimizer updates parameters in place, so the addresses are stable and a captured step can reuse the table)
PACK_KERNEL = os.environ.get("QMP_PACK_KERNEL", "1") != "0"    # weight packs by qmp_pack_tconv_fwd / _bwd (False: tensor ops, the cross-check)


def _ptr_table(params):
    key = tuple(p.data_ptr() for p in params)
    hit = _ptr_tables.get(key)
    if hit is None:
        if len(_ptr_tables) > 64:
            _ptr_tables.clear()
        hit = (torch.tensor(key, dtype=torch.int64, device=params[0].device), tuple(params))
        _ptr_tables[key] = hit
    return hit[0]


class PackFusedFn(torch.autograd.Function):
    """``pack_fused`` of G TransformerConvs as ONE autograd node: forward and backward are one launch each
    (csrc/pack_params.cu: the parameters are read through a device pointer table, the pack is a bilinear form per conv).

    Built from ~40 small tensor ops per group, the pack used to leave ~100 autograd nodes per group behind; their backward
    (slice / pad / cat / bmm gradients, then one ``AccumulateGrad`` COPY per parameter because the gradients arrived as
    strided views) was ~700 of the captured step's 2 181 launches in round 1, and even as one autograd node with batched
    tensor ops the eight packs of a step cost ~400 launches (1.9 ms).  The backward hands every parameter a contiguous slice
    of one flat buffer, which ``AccumulateGrad`` adopts without copying."""

    @staticmethod
    def forward(ctx, DC, C_out, *params):
        G = len(params) // _PARAMS_PER_CONV
        D = params[0].shape[1]
        ctx.DC, ctx.C_out, ctx.G, ctx.D = DC, C_out, G, D
        ctx.shapes = [tuple(p.shape) for p in params[:_PARAMS_PER_CONV]]
        kernel = PACK_KERNEL and params[0].is_cuda and all(p.is_contiguous() and p.dtype == _f32 for p in params)
        ctx.kernel = kernel
        if kernel:
            ctx.tab = _ptr_table(params)
            pack = torch.empty(G, conv_total(DC), dtype=_f32, device=params[0].device)
            _lib.call("qmp_pack_tconv_fwd", ctx.tab, G, D, DC, C_out, pack)
            return pack
        convs = [_ParamView(params[_PARAMS_PER_CONV * g:_PARAMS_PER_CONV * (g + 1)], C_out) for g in range(G)]
        with torch.no_grad():
            pack = pack_fused(convs, DC)
        ctx.save_for_backward(*params)
        return pack

    @staticmethod
    def backward(ctx, g):
        import math
        G, DC, C, D = ctx.G, ctx.DC, ctx.C_out, ctx.D
        g = g.contiguous()
        if ctx.kernel:
            flat = torch.empty(G, C * (4 * D + 6), dtype=_f32, device=g.device)
            _lib.call("qmp_pack_tconv_bwd", ctx.tab, G, D, DC, C, g, flat)
        else:
            ps = ctx.saved_tensors
            st = lambda k: torch.stack([ps[_PARAMS_PER_CONV * i + k] for i in range(G)])
            Wq, bq, Wk, We = st(0), st(1), st(2), st(6)
            o1 = (DC + 2) * DC
            o2 = o1 + DC + 4
            o3 = o2 + FC * (DC + 4)
            o4 = o3 + FC * DC
            gW1p = g[:, :o1].view(G, DC + 2, DC)
            gb1p = g[:, o1:o2]
            gW2p = g[:, o2:o3].view(G, FC, DC + 4)[:, :C]
            gW3 = g[:, o3:o4].view(G, FC, DC)[:, :C, :D]
            gb3 = g[:, o4:o4 + C]
            # un-pad: rows / columns 0..D-1 = u part, DC, DC+1 = edge-attribute part
            gW1b = torch.cat([torch.cat([gW1p[:, :D, :D], gW1p[:, DC:DC + 2, :D]], dim=1),
                              torch.cat([gb1p[:, :D], gb1p[:, DC:DC + 2]], dim=1).unsqueeze(-1)], dim=2)        # [G, D+2, D+1]
            s = 1.0 / math.sqrt(C)
            KE = torch.cat([Wk, We], dim=2)                                   # [G, C, D+2]
            QB = torch.cat([Wq, bq.unsqueeze(-1)], dim=2)                     # [G, C, D+1]
            gKE = torch.bmm(QB, gW1b.transpose(1, 2)) * s                     # [G, C, D+2]
            gQB = torch.bmm(KE, gW1b) * s                                     # [G, C, D+1]
            pieces = [gQB[..., :D], gQB[..., D], gKE[..., :D], torch.zeros_like(bq), gW2p[..., :D], gW2p[..., DC + 2],
                      gKE[..., D:] + gW2p[..., DC:DC + 2], gW3, gb3]
            flat = torch.cat([t.reshape(G, -1) for t in pieces], dim=1)       # [G, per-conv parameter count]: one bucket per group
        views, off = [], 0
        for shape in ctx.shapes:
            n = 1
            for d in shape:
                n *= d
            views.append((off, n, shape))
            off += n
        grads = []
        for i in range(G):
            for k, (o, n, shape) in enumerate(views):
                grads.append(flat[i, o:o + n].view(shape) if ctx.needs_input_grad[2 + _PARAMS_PER_CONV * i + k] else None)
        return (None, None) + tuple(grads)


def pack_fused_fn(convs, DC):
    """``pack_fused(convs, DC)`` through PackFusedFn (one autograd node, copy-free parameter gradients)."""
    ps = []
    for c in convs:
        ps.extend(_conv_params(c))
    return PackFusedFn.apply(DC, convs[0].out_channels, *ps)


class RowsPackFn(torch.autograd.Function):
    """``torch.cat`` of flattened parameters into one vector / ``torch.stack`` of equal-length rows, with a backward that hands
    every parameter a contiguous slice of the incoming gradient (adopted by ``AccumulateGrad`` without a copy)."""

    @staticmethod
    def forward(ctx, rows, *params):
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.rows = rows
        flat = torch.cat([p.reshape(-1) for p in params])
        return flat.view(rows, -1) if rows else flat

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().view(-1).clone()      # own storage: the slices outlive this node as the parameters' .grad
        out, off = [], 0
        for k, shape in enumerate(ctx.shapes):
            n = 1
            for d in shape:
                n *= d
            out.append(g[off:off + n].view(shape) if ctx.needs_input_grad[1 + k] else None)
            off += n
        return (None,) + tuple(out)


def _pad8(v):
    return (v + 7) // 8 * 8


def _pad16(v):
    return (v + 15) // 16 * 16


def tc_image_bytes(DC, kind=0):
    """Bytes of one conv's weight image (csrc/fused_tc.cuh: TcFwdLayout / TcBwdTLayout / TcBwdSLayout)."""
    K1, K2 = _pad8(DC), _pad8(DC + 4)
    if kind == 0:
        N1 = 16 if DC + 2 <= 16 else 48
        return 8 * (N1 * K1 + FC * K1 + FC * K2) + 4 * (48 + FC)
    if kind == 1:
        N2, N1P = _pad16(K2), _pad16(K1)
        return 8 * (N2 * FC + N1P * FC + N1P * K2) + 4 * (DC * K1 + K1)
    assert kind == 2
    return 8 * _pad16(K1) * (FC + K1 + 8)


_img_cache = {}


def tc_image(w, DC, kind=0):
    """[G, image bytes] uint8: the tensor-core image of a padded pack (built by qmp_fused_pack_tc, cached per pack)."""
    key = (w.data_ptr(), tuple(w.shape), DC, kind, w._version)
    hit = _img_cache.get(key)
    if hit is not None and hit[0] is w:
        return hit[1]
    G = w.shape[0]
    img = torch.empty(G, tc_image_bytes(DC, kind), dtype=torch.uint8, device=w.device)
    _lib.call("qmp_fused_pack_tc", w.detach().contiguous(), G, DC, kind, img)
    if len(_img_cache) > 64:
        _img_cache.clear()
    _img_cache[key] = (w, img)
    return img


class _Holder:
    """Gradient accumulator of a weight pack shared by the FusedGroupFn calls of one forward pass."""
    __slots__ = ("acc", "handed")

    def __init__(self, like):
        self.acc = torch.zeros_like(like)
        self.handed = False


class _GradAccum(torch.autograd.Function):
    """Identity on a pack that many FusedGroupFn calls of one forward pass consume (one per timestep).  The consumers
    accumulate their gradients of the pack IN PLACE in ``holder.acc`` (the weight-gradient and gate-backward kernels add
    with reductions); only the first consumer to run in the backward pass hands ``acc`` to autograd, the others return
    None.  Autograd runs this node once, after all of them, with the complete sum: no per-timestep zero fill and no
    per-timestep gradient-accumulation add (11 small launches per decoder frame before)."""

    @staticmethod
    def forward(ctx, pack, holder):
        ctx.holder = holder
        return pack.clone()

    @staticmethod
    def backward(ctx, grad):
        h = ctx.holder
        if h.handed and grad.data_ptr() != h.acc.data_ptr():
            # the consumers kept adding into h.acc after the first one handed it over: that is only the complete sum while
            # autograd passes the handed tensor through by reference (no copy, no second non-None gradient for the pack)
            raise RuntimeError("fused._GradAccum: autograd handed over a copy of the shared accumulator; the in-place "
                               "accumulation of the pack gradients is not valid with this PyTorch build")
        out = grad.clone()
        h.acc.zero_()                 # ready for another backward pass over the same graph
        h.handed = False
        return out, None


ACC_HITS = 0        # consumers that accumulated into a shared holder (tests check the path is taken)


def shared_pack(pack):
    """``pack`` routed through _GradAccum when it needs a gradient (training), unchanged otherwise."""
    if not (torch.is_grad_enabled() and pack.requires_grad):
        return pack
    holder = _Holder(pack)
    out = _GradAccum.apply(pack, holder)
    out._qmp_acc = holder
    return out


def hand_over(holder, grad):
    """What a consumer returns to autograd for a pack: the accumulator once, None afterwards."""
    if holder is None:
        return grad
    if holder.handed:
        return None
    holder.handed = True
    return holder.acc


_cell_cache = {}


def cell_image(wa, wb):
    """uint8 image of the decoder cell's eight convs (qmp_fused_pack_cell; layout csrc/fused_cell.cuh), cached per pack pair."""
    key = (wa.data_ptr(), wb.data_ptr(), wa._version, wb._version)
    hit = _cell_cache.get(key)
    if hit is not None and hit[0] is wa and hit[1] is wb:
        return hit[2]
    img = torch.empty(int(_lib.lib().qmp_fused_cell_image_bytes()), dtype=torch.uint8, device=wb.device)
    _lib.call("qmp_fused_pack_cell", wa.detach().contiguous(), wb.detach().contiguous(), img)
    if len(_cell_cache) > 64:
        _cell_cache.clear()
    _cell_cache[key] = (wa, wb, img)
    return img


_cellb_cache = {}


def cell_bwd_image(wa, wb):
    """uint8 image for qmp_fused_cell_bwd (qmp_fused_pack_cell_bwd), cached per pack pair."""
    key = (wa.data_ptr(), wb.data_ptr(), wa._version, wb._version)
    hit = _cellb_cache.get(key)
    if hit is not None and hit[0] is wa and hit[1] is wb:
        return hit[2]
    img = torch.empty(int(_lib.lib().qmp_fused_cell_bwd_image_bytes()), dtype=torch.uint8, device=wb.device)
    _lib.call("qmp_fused_pack_cell_bwd", wa.detach().contiguous(), wb.detach().contiguous(), img)
    if len(_cellb_cache) > 64:
        _cellb_cache.clear()
    _cellb_cache[key] = (wa, wb, img)
    return img


_headb_cache = {}
HEAD_BWD = os.environ.get("QMP_HEAD_BWD", "1") != "0"     # head conv fc_out1 backward: the persistent octet kernel (csrc/head_bwd.cu)


def head_bwd_image(wb):
    """uint8 image for qmp_head_bwd (qmp_pack_head_bwd), cached per pack."""
    key = (wb.data_ptr(), wb._version)
    hit = _headb_cache.get(key)
    if hit is not None and hit[0] is wb:
        return hit[1]
    img = torch.empty(int(_lib.lib().qmp_head_bwd_image_bytes()), dtype=torch.uint8, device=wb.device)
    _lib.call("qmp_pack_head_bwd", wb.detach().contiguous(), img)
    if len(_headb_cache) > 64:
        _headb_cache.clear()
    _headb_cache[key] = (wb, img)
    return img


def is_decoder_cell(DA, GA, DB, GB, sharedB, mode, C, xa, xb):
    return (mode == 1 and GA == 4 and GB == 4 and sharedB and DA == 4 and DB == 32 and C == FC and xa is not None
            and xa.shape[1] % 4 == 0 and xb.shape[1] % 4 == 0)


class FusedGroupFn(torch.autograd.Function):
    """One fused launch: segment A (xa, wa: GA convs on a narrow shared input) + segment B (xb, wb: GB convs),
    mode 1 -> LSTM gates (+ norms, head input), mode 0 -> plain conv outputs [N, (GA+GB)*C]."""

    @staticmethod
    def forward(ctx, xa, wa, xb, wb, Cprev, params, concat, csr, cfg):
        (DA, GA, DB, GB, sharedB, mode, relu_out, C, norm_h, norm_c, norm_o, want_head, eps, drop_p, seed) = cfg
        N = xb.shape[0]
        dev = xb.device
        xb = xb.contiguous()
        xa = xa.contiguous() if xa is not None else None
        wa = wa.contiguous() if wa is not None else None
        wb = wb.contiguous()
        NC = GA + GB
        E = csr.n_edges
        logit = torch.empty(max(E, 1), NC, dtype=_f32, device=dev)
        mstat = torch.empty(N, NC, dtype=_f32, device=dev)
        linv = torch.empty(N, NC, dtype=_f32, device=dev)
        out = gates = Craw = O = H = Cn = head = None
        if mode == 1:
            gates = torch.empty(N, 4 * FC, dtype=_f32, device=dev)
            Craw, O, H, Cn = (torch.empty(N, FC, dtype=_f32, device=dev) for _ in range(4))
            head = torch.empty(N, HEADW, dtype=_f32, device=dev) if want_head else None
            if want_head and concat is None:
                head.zero_()
        else:
            out = torch.empty(N, NC * C, dtype=_f32, device=dev)
        Cp = Cprev.contiguous() if Cprev is not None else None
        cc = concat.contiguous().reshape(-1) if (want_head and concat is not None) else None
        prm = params.contiguous() if params is not None else None
        entry, wa_k, wb_k = "qmp_fused_fwd", wa, wb
        usave = None
        if TC_FWD and CELL_FWD and is_decoder_cell(DA, GA, DB, GB, sharedB, mode, C, xa, xb):
            if CELL_BWD and TC_BWD and any(ctx.needs_input_grad):      # the backward kernel reads the logit projections
                usave = torch.empty(N, 4 * FC, dtype=_f32, device=dev)
            _lib.call("qmp_fused_cell_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, xa.shape[1], xb, xb.shape[1],
                      cell_image(wa, wb), Cp, prm, int(norm_h), int(norm_c), int(norm_o), float(eps), gates, Craw, O, H, Cn, head,
                      HEADW, cc, logit, mstat, linv, usave, float(drop_p), int(seed))
            entry = None
        elif TC_FWD:
            entry = "qmp_fused_fwd_tc"
            wa_k = tc_image(wa, cap_of(DA, True)) if GA else None
            wb_k = tc_image(wb, cap_of(DB, False))
        if entry is not None:
            _lib.call(entry, N, csr.in_ptr, csr.in_src, csr.edge_attr_in,
                      xa, xa.shape[1] if xa is not None else 0, DA, GA, wa_k,
                      xb, xb.shape[1], DB, GB, int(sharedB), wb_k,
                      mode, int(bool(relu_out)), C, out, NC * C, Cp, prm, int(norm_h), int(norm_c), int(norm_o), float(eps),
                      gates, Craw, O, H, Cn, head, HEADW, cc, logit, mstat, linv, float(drop_p), int(seed))
        ctx.save_for_backward(xa, wa, xb, wb, Cp, prm, logit, mstat, linv, gates, Craw, out if relu_out == 1 else None, usave)
        ctx.set_materialize_grads(False)          # unused outputs arrive as None, not as zero-filled tensors
        ctx.holders = tuple(getattr(t, "_qmp_acc", None) if t is not None else None for t in (wa, wb, prm))
        ctx.csr, ctx.cfg = csr, cfg
        ctx.concat_shape = tuple(concat.shape) if concat is not None else None
        if mode == 1:
            return O, H, Cn, head
        return out

    @staticmethod
    def backward(ctx, *grads):
        from .fused_bwd import fused_group_backward
        return fused_group_backward(ctx, *grads)


# ---- one-output-channel TransformerConv (the decoder's fc_out2) ---------------------------------------------------------
def pack_tconv1(conv):
    """P [136] = Wq | Wk | Wv | Ws (32 each) | bq bk bv bs | we0 we1 | 0 0 from a PyG-named TransformerConv(32 -> 1)
    (layout: csrc/tconv1.cu).  No 1/sqrt(C) factor: C = 1."""
    assert conv.out_channels == 1 and conv.in_channels == FC
    return RowsPackFn.apply(0, conv.lin_query.weight, conv.lin_key.weight, conv.lin_value.weight, conv.lin_skip.weight,
                            conv.lin_query.bias, conv.lin_key.bias, conv.lin_value.bias, conv.lin_skip.bias,
                            conv.lin_edge.weight, conv.lin_edge.weight.new_zeros(2))


HEAD_TAIL = os.environ.get("QMP_HEAD_TAIL", "1") != "0"    # fc_out2 + tanh / residual tail of the decoder head as ONE autograd node (two launches each way)


class HeadTailFn(torch.autograd.Function):
    """The tail of the decoder head (model/seq2seq.py:167-187, 427-428) in two launches each way:
    ``y = TransformerConv(32 -> 1)(h)``, ``out = tanh(dropout(y)) + x[:, :1]`` (-> sigmoid if binary), ``x_next = [out, x[:, 1:]]``
    (qmp_head_tail_fwd / qmp_head_tail_bwd = ScalarTConvFn + ops.HeadFinishFn without the launches in between).  With
    ``relu_mask`` the gradient returned for ``h`` is already masked by ``h > 0``: the producer of ``h = relu(.)``
    (FusedGroupFn with ``relu_out = 2``) then skips its own mask launch."""

    @staticmethod
    def forward(ctx, h, P, x, csr, drop_attn, seed_attn, binary, drop_out, seed_out, relu_mask):
        N = h.shape[0]
        h, P, x = h.contiguous(), P.contiguous(), x.contiguous()
        dev = h.device
        s4 = torch.empty(N, 4, dtype=_f32, device=dev)
        y = torch.empty(N, 1, dtype=_f32, device=dev)
        out = torch.empty(N, 1, dtype=_f32, device=dev)
        x_next = torch.empty_like(x)
        cfg = (int(bool(binary)), float(drop_attn), int(seed_attn), float(drop_out), int(seed_out))
        _lib.call("qmp_head_tail_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, h, h.shape[1], P, x, x.shape[1], cfg[0], cfg[1], cfg[2],
                  cfg[3], cfg[4], s4, y, out, x_next)
        ctx.save_for_backward(h, P, s4, y, out, x)
        ctx.set_materialize_grads(False)
        ctx.csr, ctx.cfg, ctx.relu_mask = csr, cfg, int(bool(relu_mask))
        ctx.holder = getattr(P, "_qmp_acc", None)
        return out, x_next

    @staticmethod
    def backward(ctx, d_out, d_xnext):
        global ACC_HITS
        h, P, s4, y, out, x = ctx.saved_tensors
        N, csr, hd = h.shape[0], ctx.csr, ctx.holder
        binary, drop_attn, seed_attn, drop_out, seed_out = ctx.cfg
        ACC_HITS += hd is not None
        dev = h.device
        d_out = d_out.contiguous() if d_out is not None else None
        d_xnext = d_xnext.contiguous() if d_xnext is not None else None
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        gP = (hd.acc if hd is not None else torch.zeros_like(P)) if ctx.needs_input_grad[1] else None
        dx = torch.empty_like(x)
        ds4 = torch.empty(N, 4, dtype=_f32, device=dev)
        _lib.call("qmp_head_tail_bwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, h, h.shape[1], P, s4, y, out, x, x.shape[1], binary,
                  drop_attn, seed_attn, drop_out, seed_out, d_out, d_xnext, ds4, dh, h.shape[1], ctx.relu_mask, dx, gP)
        return (dh, (hand_over(hd, gP) if gP is not None else None), dx if ctx.needs_input_grad[2] else None, None, None, None, None, None,
                None, None)


class ScalarTConvFn(torch.autograd.Function):
    """out [N, 1] = TransformerConv(32 -> 1)(x) on the scalar-record kernels (qmp_tconv1_fwd / qmp_tconv1_bwd)."""

    @staticmethod
    def forward(ctx, x, P, csr, drop_p, seed):
        N = x.shape[0]
        x = x.contiguous()
        P = P.contiguous()
        s4 = torch.empty(N, 4, dtype=_f32, device=x.device)
        out = torch.empty(N, 1, dtype=_f32, device=x.device)
        _lib.call("qmp_tconv1_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, x, x.shape[1], P, s4, out, float(drop_p), int(seed))
        ctx.save_for_backward(x, P, s4)
        ctx.csr, ctx.drop_p, ctx.seed = csr, float(drop_p), int(seed)
        ctx.holder = getattr(P, "_qmp_acc", None)
        return out

    @staticmethod
    def backward(ctx, g):
        global ACC_HITS
        x, P, s4 = ctx.saved_tensors
        N, csr, h = x.shape[0], ctx.csr, ctx.holder
        ACC_HITS += h is not None
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gP = (h.acc if h is not None else torch.zeros_like(P)) if ctx.needs_input_grad[1] else None
        ds4 = torch.empty(N, 4, dtype=_f32, device=x.device)
        _lib.call("qmp_tconv1_bwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, x, x.shape[1], P, s4, g.contiguous(), ds4, dx,
                  x.shape[1], gP, ctx.drop_p, ctx.seed)
        return dx, (hand_over(h, gP) if gP is not None else None), None, None, None
