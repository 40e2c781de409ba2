"""Backward of FusedGroupFn: gate backward (qmp_lstm_gates_bwd) -> decoder-cell kernel, or the one-pass per-conv kernel
(target + source side of every edge in one launch; fused.ONEPASS_BWD = False: separate target / source launches) ->
weight-gradient reductions (qmp_gemm_tn_acc) written straight into the padded pack layout."""
from __future__ import annotations

import torch

from . import _lib
from .ops import gemm_tn_acc

FC = 32
_f32 = torch.float32
_bwd_cache = {}
_NO_GATES = (None, None, None, None, 0, 0, 0, 0.0, None, None, None, None, 0, None, None)     # qmp_fused_cell_bwd with dP as an input


def bwd_pack(w, DC):
    """[G, TOTAL_bwd] = W1 | b1 | W1T | W2T | W3T from the forward pack (layout: csrc/fused_bwd.inl).  Cached per
    forward pack: the weights are constant over the timesteps of one forward pass."""
    key = (w.data_ptr(), tuple(w.shape), DC, w._version)
    hit = _bwd_cache.get(key)
    if hit is not None and hit[0] is w:
        return hit[1]
    G = w.shape[0]
    orig, w = w, w.detach()
    o1 = (DC + 2) * DC
    o2 = o1 + DC + 4
    o3 = o2 + FC * (DC + 4)
    o4 = o3 + FC * DC
    W1 = w[:, :o1].view(G, DC + 2, DC)
    b1 = w[:, o1:o2]
    W2 = w[:, o2:o3].view(G, FC, DC + 4)
    W3 = w[:, o3:o4].view(G, FC, DC)
    W1T = torch.zeros(G, DC, DC + 4, dtype=_f32, device=w.device)
    W1T[:, :, :DC + 2] = W1.transpose(1, 2)
    pack = torch.cat([W1.flatten(1), b1, W1T.flatten(1), W2.transpose(1, 2).flatten(1), W3.transpose(1, 2).flatten(1)],
                     dim=1).contiguous()
    if len(_bwd_cache) > 64:
        _bwd_cache.clear()
    _bwd_cache[key] = (orig, pack)
    return pack


def _weight_grads(gw, DC, D, G, x, ldx, shared, Zs, dUs, ldz, dP, lddp, dp_off, dp_stride, ma, N):
    """Reductions over the N nodes, written into the flat pack gradient gw [G, TOTAL] (zero-initialised)."""
    total = gw.shape[1]
    o1 = (DC + 2) * DC
    o2 = o1 + DC + 4
    o3 = o2 + FC * (DC + 4)
    o4 = o3 + FC * DC
    W = DC + 4
    sx = 0 if shared else D
    dPv = dP[:, dp_off:]
    # logit weights: [du, dw] (x) [x | 1]
    gemm_tn_acc(dUs, x, gw[:, 0:], N, DC + 2, D, ldz, ldx, DC, sA=W, sB=sx, sC=total, batch=G)
    gemm_tn_acc(dUs, x, gw[:, o1:], N, DC + 2, 1, ldz, ldx, 1, sA=W, sB=sx, sC=total, batch=G, b_ones=1)
    # value / edge / value-bias: dP (x) [z | ze | zs]
    gemm_tn_acc(dPv, Zs, gw[:, o2:], N, ma, DC + 3, lddp, ldz, DC + 4, sA=dp_stride, sB=W, sC=total, batch=G)
    # skip: dP (x) [x | 1]
    gemm_tn_acc(dPv, x, gw[:, o3:], N, ma, D, lddp, ldx, DC, sA=dp_stride, sB=sx, sC=total, batch=G)
    gemm_tn_acc(dPv, x, gw[:, o4:], N, ma, 1, lddp, ldx, 1, sA=dp_stride, sB=sx, sC=total, batch=G, b_ones=1)


def fused_group_backward(ctx, *grads):
    from .fused import HEADW, cap_of, hand_over
    from . import fused as _fz
    if all(g is None for g in grads):
        return (None,) * 9
    xa, wa, xb, wb, Cp, prm, logit, mstat, linv, gates, Craw, out_relu, usave = ctx.saved_tensors
    ha, hb, hp = ctx.holders
    _fz.ACC_HITS += sum(h is not None for h in (ha, hb, hp))
    csr = ctx.csr
    (DA, GA, DB, GB, sharedB, mode, relu_out, C, norm_h, norm_c, norm_o, want_head, eps, drop_p, seed) = ctx.cfg
    N = xb.shape[0]
    dev = xb.device
    NC = GA + GB
    E = csr.n_edges
    DAC = cap_of(DA, True) if GA else 0
    DBC = cap_of(DB, False)
    c_ = lambda t: t.contiguous() if t is not None else None
    dCprev = dparams = dconcat = None
    need_dxa = GA > 0 and ctx.needs_input_grad[0]
    need_dxb = ctx.needs_input_grad[2]
    # (the kernel always writes both input gradients: a frame whose x or H needs none -- the first forecast step, whose
    # x is an input frame -- hands it a scratch row buffer instead of falling back to the per-conv kernels)
    cell = (_fz.CELL_BWD and _fz.TC_BWD and usave is not None and (need_dxa or need_dxb)
            and _fz.is_decoder_cell(DA, GA, DB, GB, sharedB, mode, C, xa, xb))
    gate_args = None
    if mode == 1:
        dO, dH, dC, dHead = (c_(g) for g in grads)
        dP = torch.empty(N, 4 * FC, dtype=_f32, device=dev)
        dCprev = torch.empty(N, FC, dtype=_f32, device=dev) if (Cp is not None and ctx.needs_input_grad[4]) else None
        dparams = hp.acc if hp is not None else torch.zeros(13, FC, dtype=_f32, device=dev)
        if cell and _fz.CELL_BWD_GATES:
            # the gate epilogue's backward runs in the prologue of the decoder-cell backward kernel (no launch of its own)
            gate_args = (gates, Craw, Cp, prm, int(norm_h), int(norm_c), int(norm_o), float(eps), dH, dC, dO, dHead, HEADW, dCprev,
                         dparams)
        else:
            _lib.call("qmp_lstm_gates_bwd", N, FC, gates, Craw, Cp, prm, int(norm_h), int(norm_c), int(norm_o), float(eps), dH, dC,
                      dO, dHead, HEADW, dP, 4 * FC, dCprev, dparams)
        if dHead is not None and ctx.concat_shape is not None and ctx.needs_input_grad[6]:
            dconcat = dHead[:, FC].reshape(ctx.concat_shape).clone()
        lddp = 4 * FC
    else:
        dP = c_(grads[0])
        if relu_out == 1:         # relu_out == 2: the consumer (fused.HeadTailFn) returns its gradient already masked by out > 0
            g0, dP = dP, torch.empty_like(dP)
            _lib.call("qmp_relu_mask_to", out_relu, g0, dP, dP.numel())
        lddp = NC * C

    ds = None if cell else torch.empty(max(E, 1), NC, dtype=_f32, device=dev)
    ZsA = dUsA = ZsB = dUsB = None
    if cell:        # panel layout of the streaming weight-gradient kernel (csrc/cell_wgrad.cu)
        zB = torch.empty(N, 4 * FC, dtype=_f32, device=dev)
        duB = torch.empty(N, 4 * FC, dtype=_f32, device=dev)
        sd = torch.empty(N, 64, dtype=_f32, device=dev)
        sg = torch.empty(N, 32, dtype=_f32, device=dev)
    else:
        if GA:
            ZsA = torch.empty(N, GA, DAC + 4, dtype=_f32, device=dev)
            dUsA = torch.empty(N, GA, DAC + 4, dtype=_f32, device=dev)
        ZsB = torch.empty(N, GB, DBC + 4, dtype=_f32, device=dev)
        dUsB = torch.empty(N, GB, DBC + 4, dtype=_f32, device=dev)
    from . import fused as _f
    tcb = _f.TC_BWD
    onepass = tcb and _f.ONEPASS_BWD and not cell and (need_dxa or need_dxb)
    dxa = torch.empty_like(xa) if (need_dxa or (cell and GA)) else None
    dxb = torch.empty_like(xb) if (need_dxb or cell) else None
    lda = xa.shape[1] if xa is not None else 0
    ldb = xb.shape[1]
    if cell:        # target and source side of every edge in one persistent launch (csrc/fused_cell_bwd.cu)
        _lib.call("qmp_fused_cell_bwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, lda, xb, ldb, _f.cell_bwd_image(wa, wb),
                  usave, dP, lddp, *(gate_args or _NO_GATES), logit, mstat, linv, zB, duB, sd, sg, dxa, dxb, float(drop_p), int(seed))
    elif tcb:
        pa = _f.tc_image(wa, DAC, 1) if GA else None
        pb = _f.tc_image(wb, DBC, 1)
    else:
        pa = bwd_pack(wa, DAC) if GA else None
        pb = bwd_pack(wb, DBC)
    head = (not cell and onepass and _f.HEAD_BWD and GA == 0 and GB == 1 and DB == 36 and C == FC and mode == 0 and need_dxb
            and ldb % 4 == 0 and lddp % 4 == 0)
    if head:        # the head conv fc_out1: persistent octet kernel (csrc/head_bwd.cu), same outputs as the one-pass kernel
        _lib.call("qmp_head_bwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xb, ldb, _f.head_bwd_image(wb), dP, lddp, logit, mstat,
                  linv, ZsB, dUsB, dxb, float(drop_p), int(seed))
    elif not cell:
        _lib.call(("qmp_fused_bwd_onepass_tc" if onepass else "qmp_fused_bwd_target_tc") if tcb else "qmp_fused_bwd_target", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, lda,
                  DA, GA, pa, xb, ldb, DB, GB, int(sharedB), pb, mode, C, dP, lddp, logit, mstat, linv, ds, ZsA, dUsA, ZsB, dUsB, dxa,
                  dxb, float(drop_p), int(seed))
    if (need_dxa or need_dxb) and not cell and not onepass:
        if tcb:
            pa = _f.tc_image(wa, DAC, 2) if GA else None
            pb = _f.tc_image(wb, DBC, 2)
        _lib.call("qmp_fused_bwd_source_tc" if tcb else "qmp_fused_bwd_source", N, csr.out_ptr, csr.out_dst, csr.out_kin, xa, lda, DA, GA, pa, xb, ldb, DB, GB,
                  int(sharedB), pb, mode, C, dP, lddp, logit, mstat, linv, ds, dxa, dxb, float(drop_p), int(seed))

    gwa = (ha.acc if ha is not None else torch.zeros_like(wa)) if GA else None
    gwb = hb.acc if hb is not None else torch.zeros_like(wb)
    if cell:        # all eight convs in one streaming launch: TMA panels -> two wide tcgen05 products per 8 nodes
        _lib.call("qmp_cell_wgrad", N, xb, ldb, dP, lddp, zB, duB, sd, sg, gwa, gwb)
    elif _f.TC_WGRAD:
        # streaming TMA -> tcgen05 launch (csrc/panel_wgrad.cu) wherever it applies; the per-problem kernel is the cross-check
        tma = _f.PANEL_WGRAD and (mode == 1 or C == FC or NC == 1) and lda % 4 == 0 and ldb % 4 == 0 and lddp % 4 == 0
        _lib.call("qmp_fused_wgrad_tma" if tma else "qmp_fused_wgrad", N, xa, lda, DA, GA, xb, ldb, DB, GB, int(sharedB), mode, C, dP, lddp, ZsA, dUsA, ZsB,
                  dUsB, gwa, gwb)
    elif mode == 1:
        if GA:
            _weight_grads(gwa, DAC, DA, GA, xa, lda, True, ZsA, dUsA, GA * (DAC + 4), dP, lddp, 0, FC, FC, N)
        if GB == 4:
            _weight_grads(gwb, DBC, DB, 4, xb, ldb, sharedB, ZsB, dUsB, 4 * (DBC + 4), dP, lddp, 0, FC, FC, N)
        else:   # 8 convs, one input block each: convs g and 4+g feed gate g
            W = DBC + 4
            for half in (0, 1):
                _weight_grads(gwb[4 * half:], DBC, DB, 4, xb[:, 4 * half * DB:], ldb, False, ZsB[:, 4 * half:], dUsB[:, 4 * half:],
                              8 * W, dP, lddp, 0, FC, FC, N)
    else:
        if GA:
            _weight_grads(gwa, DAC, DA, GA, xa, lda, True, ZsA, dUsA, GA * (DAC + 4), dP, lddp, 0, C, C, N)
        _weight_grads(gwb, DBC, DB, GB, xb, ldb, sharedB, ZsB, dUsB, GB * (DBC + 4), dP, lddp, GA * C, C, C, N)
    return (dxa if need_dxa else None, hand_over(ha, gwa) if GA else None, dxb if need_dxb else None, hand_over(hb, gwb), dCprev,
            hand_over(hp, dparams) if dparams is not None else None, dconcat, None, None)
