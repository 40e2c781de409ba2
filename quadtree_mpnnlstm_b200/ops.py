"""Autograd wrappers around the C-ABI kernels (host-side plumbing only: allocation, strides,
saved tensors).  No arithmetic on the hot path happens in PyTorch here; parameter packing (a handful
of tiny tensor ops per forward pass, outside the per-timestep loop) is the one exception and is
noted where it happens (convs.py).
"""
from __future__ import annotations

import torch

from . import _lib

_f32 = torch.float32


def gemm(A, B, bias, C, n, m, k, lda, ldb, ldc, sA=0, sB=0, sC=0, sBias=0, batch=1, b_is_kxm=0, accumulate=0, relu=0):
    _lib.call("qmp_gemm", A, B, bias, C, n, m, k, lda, ldb, ldc, sA, sB, sC, sBias, batch, b_is_kxm, accumulate, relu)


def gemm_tn_acc(A, B, C, n, ma, mb, lda, ldb, ldc, sA=0, sB=0, sC=0, batch=1, b_ones=0):
    _lib.call("qmp_gemm_tn_acc", A, B, C, n, ma, mb, lda, ldb, ldc, sA, sB, sC, batch, b_ones)


_seed_counter = [0x1234567]


def next_seed():
    """Per-call dropout seed: derived from torch's CPU generator so torch.manual_seed controls it."""
    _seed_counter[0] += 1
    return (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


class dropout_salt:
    """``with dropout_salt(t):`` -- every seeded kernel launched inside mixes the DEVICE value ``t[0]`` (int64 CUDA tensor) into
    its by-value seed when it runs.  A captured CUDA graph freezes kernel arguments, so without this every replay would
    drop the same attention edges / head outputs; the step bumps ``t`` (a device op inside the graph) and the masks are
    resampled per replay like the reference resamples per call (train.TrainStep, infer.Rollout)."""

    def __init__(self, t):
        self.t = t

    def __enter__(self):
        if self.t is not None:
            _lib.set_dropout_salt(self.t)
        return self.t

    def __exit__(self, *exc):
        if self.t is not None:
            _lib.set_dropout_salt(None)
        return False


# =============================================================================== TransformerConv group
class TConvFn(torch.autograd.Function):
    """G TransformerConvs over one graph in one pass.

    x:  [N, D] shared by the G convs (``shared=True``) or [N, G*D] (conv g reads columns g*D..g*D+D).
    W1 [G, D+2, D], b1 [G, D+2]: folded logit weights (u_i = W1[:D] x_i + b1[:D], w_i = W1[D:] x_i + b1[D:]).
    W2 [G, C, D+3]: (lin_value.weight | lin_edge.weight | lin_value.bias).
    W3 [G, C, D], b3 [G, C]: lin_skip.
    Returns out [N, G*C].  ``relu_out`` applies max(., 0) to the output (decoder head, seq2seq.py:184).
    ``base`` [N, G*C] (optional) is accumulated into IN PLACE and returned (conv_x(X) + conv_h(H) of the
    cell without a separate add).
    """

    @staticmethod
    def forward(ctx, x, W1, b1, W2, W3, b3, csr, shared, drop_p, seed, relu_out, base):
        G, C, D = W3.shape
        N = x.shape[0]
        x = x.contiguous()
        W1, b1, W2, W3, b3 = W1.contiguous(), b1.contiguous(), W2.contiguous(), W3.contiguous(), b3.contiguous()
        ldx = x.shape[1]
        assert ldx == (D if shared else G * D), (tuple(x.shape), G, D, shared)
        assert csr.edge_dim in (0, 2), "TransformerConv needs [E, 2] edge attributes"
        xoff = 0 if shared else D
        dev = x.device
        E = csr.n_edges
        U = torch.empty(N, G * (D + 2), dtype=_f32, device=dev)
        gemm(x, W1, b1, U, N, D + 2, D, ldx, D, G * (D + 2), sA=xoff, sB=(D + 2) * D, sC=D + 2, sBias=D + 2, batch=G)
        Z = torch.empty(N, G * (D + 3), dtype=_f32, device=dev)
        logit = torch.empty(max(E, 1), G, dtype=_f32, device=dev)
        mstat = torch.empty(N, G, dtype=_f32, device=dev)
        linv = torch.empty(N, G, dtype=_f32, device=dev)
        _lib.call("qmp_attn_fwd", N, G, D, csr.in_ptr, csr.in_src, csr.edge_attr_in, x, ldx, xoff, U, Z, logit, mstat,
                  linv, float(drop_p), int(seed))
        if base is not None:
            assert base.is_contiguous() and tuple(base.shape) == (N, G * C) and not relu_out
            out = base
            ctx.mark_dirty(base)
        else:
            out = torch.empty(N, G * C, dtype=_f32, device=dev)
        gemm(Z, W2, None, out, N, C, D + 3, G * (D + 3), D + 3, G * C, sA=D + 3, sB=C * (D + 3), sC=C, batch=G,
             accumulate=1 if base is not None else 0)
        gemm(x, W3, b3, out, N, C, D, ldx, D, G * C, sA=xoff, sB=C * D, sC=C, sBias=C, batch=G, accumulate=1,
             relu=1 if relu_out else 0)
        ctx.save_for_backward(x, W1, W2, W3, U, Z, logit, mstat, linv, out if relu_out else None)
        ctx.csr, ctx.shared, ctx.drop_p, ctx.seed, ctx.relu_out = csr, shared, float(drop_p), int(seed), relu_out
        ctx.has_base = base is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W1, W2, W3, U, Z, logit, mstat, linv, out = ctx.saved_tensors
        csr, shared = ctx.csr, ctx.shared
        G, C, D = W3.shape
        N, ldx = x.shape
        xoff = 0 if shared else D
        dev = x.device
        E = csr.n_edges
        dout = dout.contiguous()
        if ctx.relu_out:
            dout = dout.clone()
            _lib.call("qmp_relu_mask", out, dout, dout.numel())
        need_dx = ctx.needs_input_grad[0]
        # dZ = dOut W2
        dZ = torch.empty(N, G * (D + 3), dtype=_f32, device=dev)
        gemm(dout, W2, None, dZ, N, D + 3, C, G * C, D + 3, G * (D + 3), sA=C, sB=C * (D + 3), sC=D + 3, batch=G, b_is_kxm=1)
        ds = torch.empty(max(E, 1), G, dtype=_f32, device=dev)
        dU = torch.empty(N, G * (D + 2), dtype=_f32, device=dev)
        _lib.call("qmp_attn_bwd_target", N, G, D, csr.in_ptr, csr.in_src, csr.edge_attr_in, x, ldx, xoff, logit, mstat,
                  linv, dZ, ds, dU, ctx.drop_p, ctx.seed)
        dx = None
        if need_dx:
            dx = torch.empty_like(x)
            _lib.call("qmp_attn_bwd_source", N, G, D, csr.out_ptr, csr.out_dst, csr.out_kin, logit, mstat, linv, ds, dZ,
                      U, dx, ldx, xoff, 1 if shared else 0, 0, ctx.drop_p, ctx.seed)
            if shared:  # one contraction over all G blocks: dx += dU [N, G(D+2)] W1 [G(D+2), D] + dOut [N, GC] W3 [GC, D]
                gemm(dU, W1, None, dx, N, D, G * (D + 2), G * (D + 2), D, ldx, b_is_kxm=1, accumulate=1)
                gemm(dout, W3, None, dx, N, D, G * C, G * C, D, ldx, b_is_kxm=1, accumulate=1)
            else:
                gemm(dU, W1, None, dx, N, D, D + 2, G * (D + 2), D, ldx, sA=D + 2, sB=(D + 2) * D, sC=D, batch=G,
                     b_is_kxm=1, accumulate=1)
                gemm(dout, W3, None, dx, N, D, C, G * C, D, ldx, sA=C, sB=C * D, sC=D, batch=G, b_is_kxm=1, accumulate=1)
        # weight gradients (reductions over the N nodes)
        dW1b = torch.zeros(G, D + 2, D + 1, dtype=_f32, device=dev)
        gemm_tn_acc(dU, x, dW1b, N, D + 2, D + 1, G * (D + 2), ldx, D + 1, sA=D + 2, sB=xoff, sC=(D + 2) * (D + 1),
                    batch=G, b_ones=1)
        dW2 = torch.zeros(G, C, D + 3, dtype=_f32, device=dev)
        gemm_tn_acc(dout, Z, dW2, N, C, D + 3, G * C, G * (D + 3), D + 3, sA=C, sB=D + 3, sC=C * (D + 3), batch=G)
        dW3b = torch.zeros(G, C, D + 1, dtype=_f32, device=dev)
        gemm_tn_acc(dout, x, dW3b, N, C, D + 1, G * C, ldx, D + 1, sA=C, sB=xoff, sC=C * (D + 1), batch=G, b_ones=1)
        return (dx, dW1b[..., :D], dW1b[..., D], dW2, dW3b[..., :D], dW3b[..., D], None, None, None, None, None,
                dout if ctx.has_base else None)


# =============================================================================== GCN / Cheb pieces
class SpmmFn(torch.autograd.Function):
    """y = alpha * (S x) + beta * z, S given per in-CSR slot (csr.norm(mode)); x, z: [N, width]."""

    @staticmethod
    def forward(ctx, x, z, csr, mode, alpha, beta):
        x = x.contiguous()
        N, width = x.shape
        val = csr.norm(mode)
        y = torch.empty_like(x)
        zz = z.contiguous() if z is not None else None
        _lib.call("qmp_spmm", N, width, csr.in_ptr, csr.in_src, None, val, x, width, float(alpha), float(beta), zz,
                  width, y, width)
        ctx.csr, ctx.mode, ctx.alpha, ctx.beta, ctx.has_z = csr, mode, float(alpha), float(beta), z is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        csr = ctx.csr
        dy = dy.contiguous()
        N, width = dy.shape
        dx = dz = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(dy)
            _lib.call("qmp_spmm", N, width, csr.out_ptr, csr.out_dst, csr.out_kin, csr.norm(ctx.mode), dy, width,
                      ctx.alpha, 0.0, None, width, dx, width)
        if ctx.has_z and ctx.needs_input_grad[1]:
            dz = dy * ctx.beta
        return dx, dz, None, None, None, None


class NodeLinearFn(torch.autograd.Function):
    """out[:, g*M:(g+1)*M] = x_g W[g]^T + b[g] with x_g = x (shared) or x[:, g*K:(g+1)*K].
    W [G, M, K], b [G, M] or None."""

    @staticmethod
    def forward(ctx, x, W, b, shared):
        x, W = x.contiguous(), W.contiguous()
        G, M, K = W.shape
        N, ldx = x.shape
        assert ldx == (K if shared else G * K)
        xoff = 0 if shared else K
        out = torch.empty(N, G * M, dtype=_f32, device=x.device)
        bb = b.contiguous() if b is not None else None
        gemm(x, W, bb, out, N, M, K, ldx, K, G * M, sA=xoff, sB=M * K, sC=M, sBias=M, batch=G)
        ctx.save_for_backward(x, W)
        ctx.shared, ctx.has_b = shared, b is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W = ctx.saved_tensors
        G, M, K = W.shape
        N, ldx = x.shape
        shared = ctx.shared
        xoff = 0 if shared else K
        dout = dout.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if shared:
                gemm(dout, W, None, dx, N, K, G * M, G * M, K, ldx, b_is_kxm=1)
            else:
                gemm(dout, W, None, dx, N, K, M, G * M, K, ldx, sA=M, sB=M * K, sC=K, batch=G, b_is_kxm=1)
        dWb = torch.zeros(G, M, K + 1, dtype=_f32, device=x.device)
        gemm_tn_acc(dout, x, dWb, N, M, K + 1, G * M, ldx, K + 1, sA=M, sB=xoff, sC=M * (K + 1), batch=G, b_ones=1)
        return dx, dWb[..., :K], (dWb[..., K] if ctx.has_b else None), None


# =============================================================================== GAT edge phase
class GatFn(torch.autograd.Function):
    """out [N, C] = sum_e alpha_e XL[j_e] with alpha = softmax over the in-edges of additive attention logits (csrc/gat.cu).
    mode 1 (GATConv): (a1, a2, a3) = (alpha_src [N], alpha_dst [N], we [2]); mode 2 (GATv2Conv): (x_r [N, C], lin_edge W [C, 2],
    att [C])."""

    @staticmethod
    def forward(ctx, XL, a1, a2, a3, csr, mode, slope):
        XL, a1, a2, a3 = XL.contiguous(), a1.contiguous(), a2.contiguous(), a3.contiguous()
        N, C = XL.shape
        out = torch.empty(N, C, dtype=_f32, device=XL.device)
        alpha = torch.empty(max(csr.n_edges, 1), dtype=_f32, device=XL.device)
        m1 = (a1, a2, a3) if mode == 1 else (None, None, None)
        m2 = (a1, C, a2, a3) if mode == 2 else (None, 0, None, None)
        _lib.call("qmp_gat_fwd", N, C, mode, csr.in_ptr, csr.in_src, csr.edge_attr_in, XL, C, *m1, *m2, float(slope), out, C, alpha)
        ctx.save_for_backward(XL, a1, a2, a3, alpha)
        ctx.csr, ctx.mode, ctx.slope = csr, mode, float(slope)
        return out

    @staticmethod
    def backward(ctx, dOut):
        XL, a1, a2, a3, alpha = ctx.saved_tensors
        csr, mode, slope = ctx.csr, ctx.mode, ctx.slope
        N, C = XL.shape
        dev = XL.device
        dOut = dOut.contiguous()
        dlog = torch.empty(max(csr.n_edges, 1), dtype=_f32, device=dev)
        dXL = torch.empty_like(XL)
        m1 = (a1, a2, a3) if mode == 1 else (None, None, None)
        m2 = (a1, C, a2, a3) if mode == 2 else (None, 0, None, None)
        if mode == 1:
            das, dad, dwe = torch.empty(N, dtype=_f32, device=dev), torch.empty(N, dtype=_f32, device=dev), torch.zeros(2, dtype=_f32, device=dev)
            g = (das, dad, dwe, None, None, None)
            ret = (das, dad, dwe)
        else:
            dXR, dWe, datt = torch.empty_like(XL), torch.zeros(C, 2, dtype=_f32, device=dev), torch.zeros(C, dtype=_f32, device=dev)
            g = (None, None, None, dXR, dWe, datt)
            ret = (dXR, dWe, datt)
        _lib.call("qmp_gat_bwd", N, C, mode, csr.in_ptr, csr.in_src, csr.edge_attr_in, XL, C, *m1, *m2, slope, alpha, dOut, C, dlog, dXL, *g)
        return (dXL,) + ret + (None, None, None)


# =============================================================================== GRU gates
class GruGates1Fn(torch.autograd.Function):
    """(Z, R, H * R) from the four conv outputs of the update / reset gates (model/model.py:240-250)."""

    @staticmethod
    def forward(ctx, az, bz, ar, br, H):
        az, bz, ar, br, H = (t.contiguous() for t in (az, bz, ar, br, H))
        Z, R, HR = torch.empty_like(H), torch.empty_like(H), torch.empty_like(H)
        _lib.call("qmp_gru_gates1_fwd", H.numel(), az, bz, ar, br, H, Z, R, HR)
        ctx.save_for_backward(Z, R, H)
        ctx.set_materialize_grads(False)
        return Z, R, HR

    @staticmethod
    def backward(ctx, dZ, dR, dHR):
        Z, R, H = ctx.saved_tensors
        c_ = lambda t: t.contiguous() if t is not None else None
        dpz, dpr, dH = torch.empty_like(H), torch.empty_like(H), torch.empty_like(H)
        _lib.call("qmp_gru_gates1_bwd", H.numel(), Z, R, H, c_(dZ), c_(dR), c_(dHR), dpz, dpr, dH)
        return dpz, dpz, dpr, dpr, dH


class GruGates2Fn(torch.autograd.Function):
    """H' = Z * H + (1 - Z) * tanh(ah + bh) (model/model.py:251-258)."""

    @staticmethod
    def forward(ctx, ah, bh, Z, H):
        ah, bh, Z, H = (t.contiguous() for t in (ah, bh, Z, H))
        Ht, Hn = torch.empty_like(H), torch.empty_like(H)
        _lib.call("qmp_gru_gates2_fwd", H.numel(), ah, bh, Z, H, Ht, Hn)
        ctx.save_for_backward(Z, H, Ht)
        return Hn

    @staticmethod
    def backward(ctx, dHn):
        Z, H, Ht = ctx.saved_tensors
        dph, dZ, dH = torch.empty_like(H), torch.empty_like(H), torch.empty_like(H)
        _lib.call("qmp_gru_gates2_bwd", H.numel(), Z, H, Ht, dHn.contiguous(), dph, dZ, dH)
        return dph, dph, dZ, dH


# =============================================================================== LSTM gates
class LstmGatesFn(torch.autograd.Function):
    """Gate epilogue + LayerNorms (+ decoder head input).  P [N, 4C]; Cprev [N, C] or None;
    params [13, C] (see csrc/lstm.cu).  Returns (O, H, Cn, head_in or None)."""

    @staticmethod
    def forward(ctx, P, Cprev, params, concat, norm_h, norm_c, norm_o, want_head, eps):
        P, params = P.contiguous(), params.contiguous()
        N, C4 = P.shape
        C = C4 // 4
        dev = P.device
        Cp = Cprev.contiguous() if Cprev is not None else None
        gates = torch.empty(N, 4 * C, dtype=_f32, device=dev)
        Craw, O, H, Cn = (torch.empty(N, C, dtype=_f32, device=dev) for _ in range(4))
        head = torch.empty(N, C + 1, dtype=_f32, device=dev) if want_head else None
        cc = concat.contiguous().reshape(-1) if (want_head and concat is not None) else None
        if want_head and cc is None:
            head.zero_()
        _lib.call("qmp_lstm_gates_fwd", N, C, P, 4 * C, Cp, params, int(norm_h), int(norm_c), int(norm_o), float(eps),
                  gates, Craw, O, H, Cn, head, C + 1, cc)
        ctx.save_for_backward(gates, Craw, Cp, params)
        ctx.cfg = (N, C, int(norm_h), int(norm_c), int(norm_o), float(eps), want_head)
        return O, H, Cn, head

    @staticmethod
    def backward(ctx, dO, dH, dC, dHead):
        gates, Craw, Cp, params = ctx.saved_tensors
        N, C, norm_h, norm_c, norm_o, eps, want_head = ctx.cfg
        dev = gates.device
        c_ = lambda t: t.contiguous() if t is not None else None
        dO, dH, dC, dHead = c_(dO), c_(dH), c_(dC), c_(dHead)
        dP = torch.empty(N, 4 * C, dtype=_f32, device=dev)
        dCprev = torch.empty(N, C, dtype=_f32, device=dev) if (Cp is not None and ctx.needs_input_grad[1]) else None
        dparams = torch.zeros(13, C, dtype=_f32, device=dev)
        _lib.call("qmp_lstm_gates_bwd", N, C, gates, Craw, Cp, params, norm_h, norm_c, norm_o, eps, dH, dC, dO, dHead,
                  C + 1, dP, 4 * C, dCprev, dparams)
        dconcat = dHead[:, C:].clone() if (dHead is not None and ctx.needs_input_grad[3]) else None
        return dP, dCprev, dparams, dconcat, None, None, None, None, None


class HeadFinishFn(torch.autograd.Function):
    """out = tanh(dropout(y)) + x[:, :1] (-> sigmoid if binary); x_next = [out, x[:, 1:]]
    (model/seq2seq.py:167-178, 427-428)."""

    @staticmethod
    def forward(ctx, y, x, binary, drop_p, seed):
        y, x = y.contiguous(), x.contiguous()
        N, F = x.shape
        out = torch.empty(N, 1, dtype=_f32, device=x.device)
        x_next = torch.empty_like(x)
        _lib.call("qmp_head_finish_fwd", y, x, N, F, int(binary), float(drop_p), int(seed), out, x_next)
        ctx.save_for_backward(y, out, x)
        ctx.cfg = (N, F, int(binary), float(drop_p), int(seed))
        return out, x_next

    @staticmethod
    def backward(ctx, d_out, d_xnext):
        y, out, x = ctx.saved_tensors
        N, F, binary, drop_p, seed = ctx.cfg
        d_out = d_out.contiguous() if d_out is not None else None
        d_xnext = d_xnext.contiguous() if d_xnext is not None else None
        dy = torch.empty(N, 1, dtype=_f32, device=x.device)
        dx = torch.empty_like(x)               # the kernel writes whole rows: column 0 and the pass-through columns
        _lib.call("qmp_head_finish_bwd", y, out, x, d_out, d_xnext, N, F, binary, drop_p, seed, dy, dx)
        return dy, dx, None, None, None
