"""The eight ChebConv / GCNConv stacks of one GConvLSTM step as ONE autograd node and ONE C call each way.

Reference: model/model.py:394-463 (``GConvLSTM`` gate pre-activations ``conv_x_g(X) + conv_h_g(H)``), model/model.py:60-97
(``GraphConv``: ``n_conv_layers`` convs applied one after the other, no activation in between), PyG 2.2.0 ``ChebConv``
(K = 3, sym, lambda_max = 2: ``T0 = x, T1 = L^ x, Tk = 2 L^ T(k-1) - T(k-2)``, ``out = sum_k Tk W_k^T + b``) and ``GCNConv``.

``Seq2Seq``'s default conv is ChebConv (model/seq2seq.py:203) and the reference's CPU-runnable configuration (BASELINE configs[0],
MNIST 64 x 64, hidden 16, dynamic quadtree) runs on meshes of a few hundred to a few thousand nodes, where a step is bound by
the NUMBER of launches, autograd nodes and Python calls, not by bytes.  The modular path (SpmmFn x K -> cat / stack ->
NodeLinearFn -> add) left ~15 autograd nodes and ~30 launches per cell step behind, and twice that in the backward pass.  Here:

* one ``torch.autograd.Function`` per cell step; its forward and backward are ONE call into the library each
  (``qmp_cheb_cell_fwd`` / ``qmp_cheb_cell_bwd``, csrc/cheb_cell.cu), which issues the propagation, contraction and
  weight-gradient launches back to back from C++ into one workspace (no Python, no allocator, no autograd between launches);
* the propagated copies are written straight into the column blocks of the basis buffer (leading dimensions instead of ``cat``);
* layers >= 1 contract ``T_k`` block by block with K accumulating grouped GEMMs (no ``stack`` to make the K blocks adjacent);
* the x-half and the h-half meet in the GEMM epilogue (``accumulate``) when there is one conv layer;
* weight and bias of a group travel as ONE pack ``[G, C, K w + 1]`` (bias = last column) whose gradient is accumulated in place
  (``qmp_gemm_tn_acc`` with a ones column) in the shared accumulator of ``fused.shared_pack`` -- no per-timestep gradient adds;
* the backward pass runs the transposed recurrence ``dT(k-1) += 2 L^T dT(k), dT(k-2) -= dT(k)`` in place on the gradient blocks.

``_fwd_py`` / ``_bwd_py`` below issue the SAME launches in the same order from Python on the same workspace layout: they are the
cross-check of the C++ sequencing (``QMP_CHEB_CELL_FN=py``, tests/test_parity_convs.py) and what the test-only CPU emulation of
the two entry points runs.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from . import fused as _fused
from .ops import gemm, gemm_tn_acc

_f32 = torch.float32
MAX_LAYERS = 3                       # conv layers per stack the C entry points take (the reference uses 1 ... 3)
USE_C = os.environ.get("QMP_CHEB_CELL_FN", "1") != "py"


def pack_linear_group(convs, kind):
    """``[G, C, K*w + 1]``: per conv the K weight blocks side by side (``lins[k].weight`` [C, w]) and the bias as last column."""
    if kind == "GCNConv":
        rows = [torch.cat([c.lin.weight, c.bias.unsqueeze(1)], dim=1) for c in convs]
    else:
        rows = [torch.cat([lin.weight for lin in c.lins] + [c.bias.unsqueeze(1)], dim=1) for c in convs]
    return torch.stack(rows).contiguous()


def bias_of(pack):
    """Contiguous copy of the bias column (the GEMM epilogue reads a dense vector), cached on the pack."""
    b = getattr(pack, "_qmp_bias", None)
    if b is None:
        b = pack.detach()[:, :, -1].contiguous()
        pack._qmp_bias = b
    return b


def layout(N, F, C, K, S, cheb):
    """Float offsets into the forward workspace (saved for the backward pass) and its size:
    Tx [N, K F] | Th [N, K C] | per layer l >= 1: in_l [N, 8C], then its propagated copies [N, 8C] each (K - 1 for Cheb: T0 is
    in_l itself; 1 for GCN) | out of the last layer [N, 8C] when S > 1."""
    w8 = 8 * C
    off = {"Tx": 0, "Th": N * K * F}
    pos = N * K * (F + C)
    nb = (K - 1) if cheb else 1
    for l in range(1, S):
        off["in", l] = pos
        pos += N * w8
        for k in range(nb):
            off["T", l, k + (1 if cheb else 0)] = pos
            pos += N * w8
    if S > 1:
        off["out"] = pos
        pos += N * w8
    return off, pos


def scratch_size(N, F, C, K, S):
    """Backward scratch: two gradient buffers [N, 8C], K gradient blocks [N, 8C], dTx [N, K F], dTh [N, K C]."""
    return (2 + K) * N * 8 * C + N * K * (F + C)


def _spmm(g, cheb, transposed, N, width, x, ldx, alpha, beta, z, ldz, y, ldy):
    if transposed:
        _lib.call("qmp_spmm", N, width, g[3], g[4], g[5], g[2], x, ldx, float(alpha), float(beta), z, ldz, y, ldy)
    else:
        _lib.call("qmp_spmm", N, width, g[0], g[1], None, g[2], x, ldx, float(alpha), float(beta), z, ldz, y, ldy)


def _fwd_py(N, F, C, K, S, cheb, g, X, H, packs, biases, ws, P):
    """The launch sequence of qmp_cheb_cell_fwd.  g = (in_ptr, in_src, val, out_ptr, out_dst, out_kin)."""
    off, _ = layout(N, F, C, K, S, cheb)
    w8 = 8 * C
    seg = lambda o, n: ws[o:o + n]
    last = S == 1
    cur = P if last else seg(off["in", 1], N * w8)
    ldc = 4 * C if last else w8
    for which, (inp, w, key) in enumerate(((X, F, "Tx"), (H, C, "Th"))):
        T = seg(off[key], N * K * w)
        ld = K * w
        if cheb:
            T.view(N, ld)[:, :w].copy_(inp.view(N, w))
            if K > 1:
                _spmm(g, cheb, False, N, w, T, ld, 1.0, 0.0, None, w, T[w:], ld)
            for k in range(2, K):
                _spmm(g, cheb, False, N, w, T[(k - 1) * w:], ld, 2.0, -1.0, T[(k - 2) * w:], ld, T[k * w:], ld)
        else:
            _spmm(g, cheb, False, N, w, inp, w, 1.0, 0.0, None, w, T, ld)
        out = cur if last else cur[which * 4 * C:]
        gemm(T, packs[which], biases[which], out, N, C, ld, ld, ld + 1, ldc, sA=0, sB=C * (ld + 1), sC=C, sBias=C, batch=4,
             accumulate=1 if (last and which == 1) else 0)
    for l in range(1, S):
        inp = seg(off["in", l], N * w8)
        if cheb:
            Ts = [inp] + [seg(off["T", l, k], N * w8) for k in range(1, K)]
            if K > 1:
                _spmm(g, cheb, False, N, w8, inp, w8, 1.0, 0.0, None, w8, Ts[1], w8)
            for k in range(2, K):
                _spmm(g, cheb, False, N, w8, Ts[k - 1], w8, 2.0, -1.0, Ts[k - 2], w8, Ts[k], w8)
        else:
            Ts = [seg(off["T", l, 0], N * w8)]
            _spmm(g, cheb, False, N, w8, inp, w8, 1.0, 0.0, None, w8, Ts[0], w8)
        nxt = seg(off["in", l + 1], N * w8) if l + 1 < S else seg(off["out"], N * w8)
        flatW = packs[l + 1].view(-1)
        for k in range(K):
            gemm(Ts[k], flatW[k * C:], biases[l + 1] if k == 0 else None, nxt, N, C, C, w8, K * C + 1, w8, sA=C, sB=C * (K * C + 1),
                 sC=C, sBias=C, batch=8, accumulate=1 if k else 0)
    if not last:
        o = seg(off["out"], N * w8).view(N, w8)
        torch.add(o[:, :4 * C], o[:, 4 * C:], out=P.view(N, 4 * C))


def _bwd_py(N, F, C, K, S, cheb, g, dP, packs, accs, ws, ws2, need_dx, need_dh, dX, dH):
    """The launch sequence of qmp_cheb_cell_bwd."""
    off, _ = layout(N, F, C, K, S, cheb)
    w8 = 8 * C
    seg = lambda o, n: ws[o:o + n]
    buf = [ws2[i * N * w8:(i + 1) * N * w8] for i in range(2 + K)]
    dTx = ws2[(2 + K) * N * w8:(2 + K) * N * w8 + N * K * F]
    dTh = ws2[(2 + K) * N * w8 + N * K * F:]

    def basis_bwd(w, dT, ld, out):
        """dT[k]: gradient blocks (leading dimension ld, changed in place) -> gradient of the basis input in ``out`` [N, w]."""
        if not cheb:
            _spmm(g, cheb, True, N, w, dT[0], ld, 1.0, 0.0, None, w, out, w)
            return
        for k in range(K - 1, 1, -1):
            _spmm(g, cheb, True, N, w, dT[k], ld, 2.0, 1.0, dT[k - 1], ld, dT[k - 1], ld)        # dT(k-1) += 2 L^T dT(k)
            a = torch.as_strided(dT[k - 2], (N, w), (ld, 1))
            a.sub_(torch.as_strided(dT[k], (N, w), (ld, 1)))                                      # dT(k-2) -= dT(k)
        if K > 1:
            _spmm(g, cheb, True, N, w, dT[1], ld, 1.0, 1.0, dT[0], ld, out, w)                    # d inp = dT0 + L^T dT1
        else:
            out[:N * w].view(N, w).copy_(torch.as_strided(dT[0], (N, w), (ld, 1)))

    if S == 1:
        dOut, ldo = dP, 4 * C
    else:
        dOut, ldo = buf[0], w8
        torch.cat([dP.view(N, 4 * C), dP.view(N, 4 * C)], dim=1, out=dOut.view(N, w8))
    nxt = 1
    for l in range(S - 1, 0, -1):
        inp = seg(off["in", l], N * w8)
        Ts = ([inp] + [seg(off["T", l, k], N * w8) for k in range(1, K)]) if cheb else [seg(off["T", l, 0], N * w8)]
        flatW, acc = packs[l + 1].view(-1), accs[l + 1].view(-1)
        dT = buf[2:2 + K]
        for k in range(K):
            gemm(dOut, flatW[k * C:], None, dT[k], N, C, C, w8, K * C + 1, w8, sA=C, sB=C * (K * C + 1), sC=C, batch=8, b_is_kxm=1)
            ones = 1 if k == K - 1 else 0
            gemm_tn_acc(dOut, Ts[k], acc[k * C:], N, C, C + ones, w8, w8, K * C + 1, sA=C, sB=C, sC=C * (K * C + 1), batch=8, b_ones=ones)
        basis_bwd(w8, dT, w8, buf[nxt])
        dOut, nxt = buf[nxt], 1 - nxt
    for which, (w, key, need, dT, res) in enumerate(((F, "Tx", need_dx, dTx, dX), (C, "Th", need_dh, dTh, dH))):
        T = seg(off[key], N * K * w)
        ld = K * w
        dout = dOut if S == 1 else dOut[which * 4 * C:]
        gemm_tn_acc(dout, T, accs[which], N, C, ld + 1, ldo, ld, ld + 1, sA=C, sB=0, sC=C * (ld + 1), batch=4, b_ones=1)
        if not need:
            continue
        gemm(dout, packs[which], None, dT, N, ld, 4 * C, ldo, ld + 1, ld, b_is_kxm=1)
        basis_bwd(w, [dT[k * w:] for k in range(K)], ld, res)


def _cheb_forward(X, H, csr, mode, K, S, C, packs):
    """P [N, 4C] and the state the backward pass needs."""
    N, F = X.shape
    cheb = mode == "cheb"
    assert 1 <= S <= MAX_LAYERS and len(packs) == S + 1
    biases = [bias_of(p) for p in packs]
    _, n_ws = layout(N, F, C, K, S, cheb)
    ws = torch.empty(n_ws, dtype=_f32, device=X.device)
    P = torch.empty(N, 4 * C, dtype=_f32, device=X.device)
    g = (csr.in_ptr, csr.in_src, csr.norm(mode), csr.out_ptr, csr.out_dst, csr.out_kin)
    if USE_C:
        pad = [None] * (MAX_LAYERS + 1 - len(packs))
        _lib.call("qmp_cheb_cell_fwd", N, F, C, K, S, int(cheb), g[0], g[1], g[2], X, H, *packs, *pad, *biases, *pad, ws, P)
    else:
        _fwd_py(N, F, C, K, S, cheb, g, X, H, packs, biases, ws, P.view(-1))
    holders = [getattr(p, "_qmp_acc", None) for p in packs]
    return P, (g, (cheb, K, S, C, N, F), ws, packs, holders, csr)        # (csr: keeps the CSR arrays alive)


def _cheb_backward(state, dP, need_dx, need_dh):
    """(dX, dH, gradients of the packs as autograd expects them)."""
    g, (cheb, K, S, C, N, F), ws, packs, holders, _ = state
    dev = dP.device
    grads = [None] * len(packs)
    accs = []
    for i, h in enumerate(holders):
        if h is not None:
            _fused.ACC_HITS += 1
            accs.append(h.acc)
        else:
            grads[i] = torch.zeros_like(packs[i])
            accs.append(grads[i])
    dX = torch.empty(N, F, dtype=_f32, device=dev) if need_dx else None
    dH = torch.empty(N, C, dtype=_f32, device=dev) if need_dh else None
    ws2 = torch.empty(scratch_size(N, F, C, K, S), dtype=_f32, device=dev)
    if USE_C:
        pad = [None] * (MAX_LAYERS + 1 - len(packs))
        _lib.call("qmp_cheb_cell_bwd", N, F, C, K, S, int(cheb), g[3], g[4], g[5], g[2], dP, *packs, *pad, *accs, *pad, ws, ws2,
                  int(need_dx), int(need_dh), dX, dH)
    else:
        _bwd_py(N, F, C, K, S, cheb, g, dP.view(-1), packs, accs, ws, ws2, need_dx, need_dh,
                dX.view(-1) if need_dx else None, dH.view(-1) if need_dh else None)
    out = [_fused.hand_over(h, gr) if h is not None else gr for h, gr in zip(holders, grads)]
    return dX, dH, tuple(out)


class ChebCellFn(torch.autograd.Function):
    """``P [N, 4C] = sum over the x and the h stack of conv_{x,h}_g(...)`` for the gates g = i, f, c, o.

    ``packs``: layer 0 -> (Wb_x [4, C, K F + 1], Wb_h [4, C, K C + 1]); layer l >= 1 -> Wb_l [8, C, K C + 1] (x stacks then h
    stacks); all through ``fused.shared_pack`` when they need gradients."""

    @staticmethod
    def forward(ctx, X, H, csr, mode, K, S, C, *packs):
        P, ctx.state = _cheb_forward(X.contiguous(), H.contiguous(), csr, mode, K, S, C, packs)
        return P

    @staticmethod
    def backward(ctx, dP):
        dX, dH, gp = _cheb_backward(ctx.state, dP.contiguous(), bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1]))
        return (dX, dH, None, None, None, None, None) + gp


class ChebLstmCellFn(torch.autograd.Function):
    """The whole cell step -- the eight stacks AND the gate epilogue (peepholes, LayerNorms, head input; csrc/lstm.cu) -- as one
    autograd node: ``(O, H', C', head_in)`` from ``X, H, C`` (model/model.py:430-463).  ``flags = (norm_h, norm_c, norm_o,
    want_head, eps)``; ``params`` [13, C] as ``GConvLSTM._gate_params`` packs them (through ``fused.shared_pack``: the gate kernel's
    parameter gradients then accumulate in place over the timesteps)."""

    @staticmethod
    def forward(ctx, X, H, Cprev, params, concat, csr, mode, K, S, C, flags, *packs):
        norm_h, norm_c, norm_o, want_head, eps = flags
        P, ctx.state = _cheb_forward(X.contiguous(), H.contiguous(), csr, mode, K, S, C, packs)
        N = P.shape[0]
        dev = P.device
        params = params.contiguous()
        Cp = Cprev.contiguous() if Cprev is not None else None
        gates = torch.empty(N, 4 * C, dtype=_f32, device=dev)
        Craw, O, Hn, Cn = (torch.empty(N, C, dtype=_f32, device=dev) for _ in range(4))
        head = torch.empty(N, C + 1, dtype=_f32, device=dev) if want_head else None
        cc = concat.contiguous().reshape(-1) if (want_head and concat is not None) else None
        if want_head and cc is None:
            head.zero_()
        _lib.call("qmp_lstm_gates_fwd", N, C, P, 4 * C, Cp, params, int(norm_h), int(norm_c), int(norm_o), float(eps),
                  gates, Craw, O, Hn, Cn, head, C + 1, cc)
        ctx.gate = (gates, Craw, Cp, params, getattr(params, "_qmp_acc", None), (int(norm_h), int(norm_c), int(norm_o), float(eps)))
        return O, Hn, Cn, head

    @staticmethod
    def backward(ctx, dO, dH, dC, dHead):
        gates, Craw, Cp, params, holder, (norm_h, norm_c, norm_o, eps) = ctx.gate
        N, C = Craw.shape
        dev = gates.device
        c_ = lambda t: t.contiguous() if t is not None else None
        dO, dH, dC, dHead = c_(dO), c_(dH), c_(dC), c_(dHead)
        dP = torch.empty(N, 4 * C, dtype=_f32, device=dev)
        dCprev = torch.empty(N, C, dtype=_f32, device=dev) if (Cp is not None and ctx.needs_input_grad[2]) else None
        if holder is not None:
            _fused.ACC_HITS += 1
            dparams = holder.acc
        else:
            dparams = torch.zeros(13, C, dtype=_f32, device=dev)
        _lib.call("qmp_lstm_gates_bwd", N, C, gates, Craw, Cp, params, norm_h, norm_c, norm_o, eps, dH, dC, dO, dHead,
                  C + 1, dP, 4 * C, dCprev, dparams)
        dconcat = dHead[:, C:].clone() if (dHead is not None and ctx.needs_input_grad[4]) else None
        dX, dHin, gp = _cheb_backward(ctx.state, dP, bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1]))
        gparams = _fused.hand_over(holder, dparams) if holder is not None else dparams
        return (dX, dHin, dCprev, gparams, dconcat, None, None, None, None, None, None) + gp


# ------------------------------------------------------------------------------------------------------------------------------
# A chain of ChebConv / GCNConv layers on one input (decoder head fc_out2(relu(fc_out1(.))), model/seq2seq.py:182-187)
def stack_layout(N, K, w0, Ms):
    """Offsets of T_l [N, K w_l] and out_l [N, M_l] (every layer but the last) in the forward workspace, widths, total."""
    pos, w, T, out, ws_ = 0, w0, [], [], []
    for l, M in enumerate(Ms):
        ws_.append(w)
        T.append(pos)
        pos += N * K * w
        out.append(pos)
        if l + 1 < len(Ms):
            pos += N * M
        w = M
    return T, out, ws_, pos


def _stack_fwd_py(N, K, cheb, w0, Ms, relus, g, X, packs, biases, ws, out):
    """The launch sequence of qmp_cheb_stack_fwd."""
    T_off, o_off, widths, _ = stack_layout(N, K, w0, Ms)
    inp = X
    for l, M in enumerate(Ms):
        w = widths[l]
        ld = K * w
        T = ws[T_off[l]:T_off[l] + N * ld]
        if cheb:
            T.view(N, ld)[:, :w].copy_(inp.view(-1)[:N * w].view(N, w))
            if K > 1:
                _spmm(g, cheb, False, N, w, T, ld, 1.0, 0.0, None, w, T[w:], ld)
            for k in range(2, K):
                _spmm(g, cheb, False, N, w, T[(k - 1) * w:], ld, 2.0, -1.0, T[(k - 2) * w:], ld, T[k * w:], ld)
        else:
            _spmm(g, cheb, False, N, w, inp, w, 1.0, 0.0, None, w, T, ld)
        o = ws[o_off[l]:o_off[l] + N * M] if l + 1 < len(Ms) else out
        gemm(T, packs[l], biases[l], o, N, M, ld, ld, ld + 1, M, relu=int(relus[l]))
        inp = o


def _stack_bwd_py(N, K, cheb, w0, Ms, relus, g, dOut, out_last, packs, accs, ws, ws2, need_dx, dX):
    """The launch sequence of qmp_cheb_stack_bwd."""
    T_off, o_off, widths, _ = stack_layout(N, K, w0, Ms)
    L = len(Ms)
    mx = max(max(Ms), max(widths))
    gbuf = [ws2[:N * mx], ws2[N * mx:2 * N * mx]]
    dT0 = ws2[2 * N * mx:]
    gcur, flip = dOut, 0
    for l in range(L - 1, -1, -1):
        w, M = widths[l], Ms[l]
        ld = K * w
        T = ws[T_off[l]:T_off[l] + N * ld]
        if relus[l]:
            y = ws[o_off[l]:o_off[l] + N * M] if l + 1 < L else out_last
            _lib.call("qmp_relu_mask_to", y, gcur, gbuf[flip], N * M)
            gcur = gbuf[flip]
            flip ^= 1
        gemm_tn_acc(gcur, T, accs[l], N, M, ld + 1, M, ld, ld + 1, b_ones=1)
        if l == 0 and not need_dx:
            break
        gemm(gcur, packs[l], None, dT0, N, ld, M, M, ld + 1, ld, b_is_kxm=1)
        res = dX if l == 0 else gbuf[flip]
        blocks = [dT0[k * w:] for k in range(K)]
        if not cheb:
            _spmm(g, cheb, True, N, w, blocks[0], ld, 1.0, 0.0, None, w, res, w)
        else:
            for k in range(K - 1, 1, -1):
                _spmm(g, cheb, True, N, w, blocks[k], ld, 2.0, 1.0, blocks[k - 1], ld, blocks[k - 1], ld)
                torch.as_strided(blocks[k - 2], (N, w), (ld, 1)).sub_(torch.as_strided(blocks[k], (N, w), (ld, 1)))
            if K > 1:
                _spmm(g, cheb, True, N, w, blocks[1], ld, 1.0, 1.0, blocks[0], ld, res, w)
            else:
                res[:N * w].view(N, w).copy_(torch.as_strided(blocks[0], (N, w), (ld, 1)))
        gcur = res
        flip ^= 1


class ChebStackFn(torch.autograd.Function):
    """``x -> conv_L(...relu?(conv_1(x)))`` for up to three ChebConv / GCNConv layers as one autograd node and one library call
    each way (``qmp_cheb_stack_fwd`` / ``_bwd``).  ``packs[l]``: ``[1, M_l, K w_l + 1]`` (``pack_linear_group([conv], kind)``)."""

    @staticmethod
    def forward(ctx, x, csr, mode, K, relus, *packs):
        x = x.contiguous()
        N, w0 = x.shape
        cheb = mode == "cheb"
        L = len(packs)
        assert 1 <= L <= MAX_LAYERS and len(relus) == L
        Ms = [int(p.shape[1]) for p in packs]
        biases = [bias_of(p) for p in packs]
        *_, n_ws = stack_layout(N, K, w0, Ms)
        ws = torch.empty(max(n_ws, 1), dtype=_f32, device=x.device)
        out = torch.empty(N, Ms[-1], dtype=_f32, device=x.device)
        g = (csr.in_ptr, csr.in_src, csr.norm(mode), csr.out_ptr, csr.out_dst, csr.out_kin)
        pad = [None] * (MAX_LAYERS - L)
        if USE_C:
            _lib.call("qmp_cheb_stack_fwd", N, K, int(cheb), L, w0, *(Ms + [0] * (MAX_LAYERS - L)),
                      *([int(r) for r in relus] + [0] * (MAX_LAYERS - L)), g[0], g[1], g[2], x, *packs, *pad, *biases, *pad, ws, out)
        else:
            _stack_fwd_py(N, K, cheb, w0, Ms, relus, g, x.view(-1), packs, biases, ws, out.view(-1))
        ctx.state = (g, (N, K, cheb, w0, Ms, tuple(relus)), ws, out, packs, [getattr(p, "_qmp_acc", None) for p in packs], csr)
        return out

    @staticmethod
    def backward(ctx, dOut):
        g, (N, K, cheb, w0, Ms, relus), ws, out, packs, holders, _ = ctx.state
        dev = dOut.device
        dOut = dOut.contiguous()
        L = len(packs)
        grads = [None] * L
        accs = []
        for i, h in enumerate(holders):
            if h is not None:
                _fused.ACC_HITS += 1
                accs.append(h.acc)
            else:
                grads[i] = torch.zeros_like(packs[i])
                accs.append(grads[i])
        need_dx = bool(ctx.needs_input_grad[0])
        dX = torch.empty(N, w0, dtype=_f32, device=dev) if need_dx else None
        widths = stack_layout(N, K, w0, Ms)[2]
        mx = max(max(Ms), max(widths))
        ws2 = torch.empty(2 * N * mx + N * K * max(widths), dtype=_f32, device=dev)
        pad = [None] * (MAX_LAYERS - L)
        if USE_C:
            _lib.call("qmp_cheb_stack_bwd", N, K, int(cheb), L, w0, *(Ms + [0] * (MAX_LAYERS - L)),
                      *([int(r) for r in relus] + [0] * (MAX_LAYERS - L)), g[3], g[4], g[5], g[2], dOut, out, *packs, *pad, *accs, *pad,
                      ws, ws2, int(need_dx), dX)
        else:
            _stack_bwd_py(N, K, cheb, w0, Ms, relus, g, dOut.view(-1), out.view(-1), packs, accs, ws, ws2, need_dx,
                          dX.view(-1) if need_dx else None)
        gp = [_fused.hand_over(h, gr) if h is not None else gr for h, gr in zip(holders, grads)]
        return (dX, None, None, None, None) + tuple(gp)
