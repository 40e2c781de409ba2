"""TEST-ONLY emulation of the C ABI on CPU tensors.

Purpose: exercise the HOST logic of ``quadtree_mpnnlstm_b200`` (argument order, strides, saved tensors,
autograd wiring, driver control flow) and the kernels' FORMULAS (folded attention weights, the LSTM /
LayerNorm backward, CSR gathers) in the no-GPU container, against the oracle.  Each function below
restates what the CUDA kernel of the same name computes, on flat memory with the same pointer / leading-
dimension conventions.  It is installed by monkey-patching in ``tests/test_host_logic.py`` only; the
product never imports it and has no CPU path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from oracle import graph_ref as G


def flat(t, count=None):
    """1-D window on the storage behind ``t`` starting at t's first element (pointer semantics)."""
    if t is None:
        return None
    base = t.detach()
    n = base.untyped_storage().nbytes() // base.element_size() - base.storage_offset()
    w = torch.as_strided(base, (n,), (1,), base.storage_offset())
    return w if count is None else w[:count]


def win(f, shape, strides, off=0):
    """Strided window whose offset is RELATIVE to the first element of ``f`` (as_strided's is absolute)."""
    return torch.as_strided(f, shape, strides, f.storage_offset() + off)


def rows(t, n, ld, width):
    """[n, width] strided window with leading dimension ld starting at t's pointer."""
    base = t.detach()
    return torch.as_strided(base, (n, width), (ld, 1), base.storage_offset())


# ------------------------------------------------------------------------------------------ graph side
def qmp_exclusive_scan_i32(inp, out, n, total, scratch):
    v = flat(inp, n).long()
    c = torch.cumsum(v, 0)
    flat(out, n).copy_((c - v).int())
    if total is not None:
        flat(total, 1)[0] = int(c[-1]) if n else 0


def qmp_add_positional_encoding(x, B, H, W, C, out):
    xi = flat(x, B * H * W * C).view(B, H, W, C)
    flat(out, B * H * W * (C + 2)).view(B, H, W, C + 2).copy_(G.add_positional_encoding(xi))


def qmp_frame_max_pad(x, T, H, W, C, n_pad, m_pad, crit):
    xi = flat(x, T * H * W * C).view(T, H, W, C)
    fr = xi[..., 0].max(0).values.numpy()
    fr = np.pad(fr, ((0, n_pad - H), (0, m_pad - W)), mode="edge")
    flat(crit, n_pad * m_pad).copy_(torch.from_numpy(fr).reshape(-1))


def qmp_quadtree_labels(crit, mask, hir, n, m, S, cond, thresh, labels, rect, npix, n_nodes, split, cnt, base_off,
                        top_f, top_b):
    n_pad, m_pad = -(n // -S) * S, -(m // -S) * S
    cr = flat(crit, n_pad * m_pad).view(n_pad, m_pad).numpy()
    mk = flat(mask, n * m).view(n, m).numpy().astype(bool) if mask is not None else None
    hr = flat(hir, n * m).view(n, m).numpy().astype(bool) if hir is not None else None
    lab = G.quadtree_labels_on_padded(cr, n, m, thresh, S, mk, hr, G.CONDITIONS[cond])
    flat(labels, n * m).copy_(torch.from_numpy(lab.reshape(-1)).int())
    N = int(lab.max()) + 1 if lab.size and lab.max() >= 0 else 0
    flat(n_nodes, 1)[0] = N
    rc, npx = flat(rect, 4 * max(N, 1)).view(-1, 4), flat(npix, max(N, 1))
    for v in range(N):
        ys, xs = np.nonzero(lab == v)
        rc[v] = torch.tensor([ys.min(), xs.min(), ys.max() - ys.min() + 1, xs.max() - xs.min() + 1], dtype=torch.int32)
        npx[v] = float(len(ys))


def qmp_mesh_pixels_from_rects(labels, n, m, rect, npix, n_nodes, cap, pix_ptr, pix_idx, tmp, bs):
    N = int(flat(n_nodes, 1)[0])
    lab = flat(labels, n * m).long()
    counts = torch.zeros(cap, dtype=torch.int64)
    counts[:N] = flat(npix, N).long()
    ptr = torch.zeros(cap + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    flat(pix_ptr, cap + 1).copy_(ptr.int())
    order = torch.argsort(torch.where(lab >= 0, lab, torch.full_like(lab, cap + 1)), stable=True)
    nv = int((lab >= 0).sum())
    flat(pix_idx, nv).copy_(order[:nv].int())


def qmp_mesh_pixelwise(mask, P, labels, pix_ptr, pix_idx, npix, n_nodes, keep, rank, bs):
    mk = flat(mask, P).bool() if mask is not None else torch.zeros(P, dtype=torch.bool)
    lab = torch.from_numpy(G.pixelwise_labels(mk.numpy().reshape(1, -1)).reshape(-1))
    N = int((~mk).sum())
    flat(labels, P).copy_(lab.int())
    flat(n_nodes, 1)[0] = N
    flat(pix_idx, N).copy_(torch.nonzero(~mk).squeeze(1).int())
    flat(pix_ptr, N + 1).copy_(torch.arange(N + 1).int())
    flat(npix, N).fill_(1.0)


def qmp_segment_sum(img, B, P, C, pix_ptr, pix_idx, npix, n_cap, n_nodes_dev, divide, single, out):
    from oracle.graph_ref import lane_tree_segment_sum
    N = int(flat(n_nodes_dev, 1)[0]) if n_nodes_dev is not None else n_cap
    im = flat(img, B * P * C).view(B, P, C)
    ptr = flat(pix_ptr, N + 1).long()
    idx = flat(pix_idx, int(ptr[N])).long()
    res = lane_tree_segment_sum(im[:, idx], ptr)          # the defined summation order of csrc/pool.cu
    if divide:
        res = res / flat(npix, N)[None, :, None]
    flat(out, B * n_cap * C).view(B, n_cap, C)[:, :N] = res


def qmp_gather_by_label(data, B, P, C, n_stride, labels, npix, divide, fill, img):
    lab = flat(labels, P).long()
    d = flat(data, B * n_stride * C).view(B, n_stride, C)
    res = d[:, lab.clamp(min=0)]
    if divide:
        res = res / flat(npix, n_stride)[lab.clamp(min=0)][None, :, None]
    res = torch.where((lab >= 0)[None, :, None], res, torch.full_like(res, fill))
    flat(img, B * P * C).view(B, P, C).copy_(res)


def qmp_regrid(src0, src1, B, P, C, n_src, lab_s, npix_s, src_divide, fill, pix_ptr, pix_idx, npix_d, n_dst, dst_divide, out0, out1):
    """csrc/pool.cu regrid_kernel: gather by the source labels, then the defined segment sum over the destination pixel lists."""
    for src, out in ((src0, out0), (src1, out1)):
        if src is None or out is None:
            continue
        if lab_s is None:
            img = src
        else:
            img = torch.empty(B * P * C, dtype=torch.float32)
            qmp_gather_by_label(src, B, P, C, n_src, lab_s, npix_s, src_divide, fill, img)
        qmp_segment_sum(img, B, P, C, pix_ptr, pix_idx, npix_d, n_dst, None, dst_divide, 0, out)


def _emit_edges(ei, src64, dst64, src32, dst32, n_edges):
    E = ei.shape[1]
    flat(src64, E).copy_(torch.from_numpy(ei[0]))
    flat(dst64, E).copy_(torch.from_numpy(ei[1]))
    flat(src32, E).copy_(torch.from_numpy(ei[0]).int())
    flat(dst32, E).copy_(torch.from_numpy(ei[1]).int())
    flat(n_edges, 1)[0] = E


def qmp_adjacency_quadtree(labels, rows_, cols, src64, dst64, src32, dst32, n_edges, keys, vals, table_cap, emit, count,
                           offset, bs):
    lab = flat(labels, rows_ * cols).view(rows_, cols).long().numpy()
    _emit_edges(G.adjacency(lab), src64, dst64, src32, dst32, n_edges)


def qmp_adjacency_pixelwise(labels, rows_, cols, src64, dst64, src32, dst32, n_edges, count, offset, bs):
    lab = flat(labels, rows_ * cols).view(rows_, cols).long().numpy()
    _emit_edges(G.adjacency_pixelwise(lab), src64, dst64, src32, dst32, n_edges)


def qmp_edge_attrs(src, dst, e_cap, n_edges_dev, pos_ii, pos_jj, pos_stride, img_w, img_h, res, two_cols, out):
    E = int(flat(n_edges_dev, 1)[0]) if n_edges_dev is not None else e_cap
    s, d = flat(src, E).long(), flat(dst, E).long()
    nmax = int(max(s.max(), d.max())) + 1 if E else 0
    ii = torch.as_strided(pos_ii.detach(), (nmax,), (pos_stride,), pos_ii.storage_offset())
    jj = torch.as_strided(pos_jj.detach(), (nmax,), (pos_stride,), pos_jj.storage_offset())
    xx, yy = ii * img_w * res, jj * img_h * res
    if two_cols:
        flat(out, 2 * E).view(E, 2).copy_(torch.stack((G.edge_angle(s, d, xx, yy), G.edge_dist(s, d, xx, yy))).T)
    else:
        flat(out, E).copy_(G.edge_dist(s, d, xx, yy))


def qmp_csr_from_edge_index(ei, E, N, src32, dst32, in_ptr, in_src, in_eid, out_ptr, out_dst, out_kin, bad, tmp, bs,
                            eid_out, kin_of_edge):
    e = flat(ei, 2 * E).view(2, E)
    s, d = e[0], e[1]
    flat(src32, E).copy_(s.int())
    flat(dst32, E).copy_(d.int())
    order_in = torch.argsort(d, stable=True)
    order_out = torch.argsort(s, stable=True)
    ptr = lambda key: torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(torch.bincount(key, minlength=N), 0)])
    flat(in_ptr, N + 1).copy_(ptr(d).int())
    flat(out_ptr, N + 1).copy_(ptr(s).int())
    flat(in_eid, E).copy_(order_in.int())
    flat(in_src, E).copy_(s[order_in].int())
    kin = torch.empty(E, dtype=torch.int64)
    kin[order_in] = torch.arange(E)
    flat(out_dst, E).copy_(d[order_out].int())
    flat(out_kin, E).copy_(kin[order_out].int())


_GB = {}      # arena pointer -> the compacted result of the last emulated qmp_quadtree_graph


def qmp_quadtree_graph(img, T, n, m, C, crit, mask, hir, S, cond, thresh, resolution, two_cols, counts_host, arena):
    """csrc/graph_build.cu: the phases of the one-launch build, composed from the per-kernel emulations above."""
    import ctypes
    P = n * m
    i32 = lambda k: torch.zeros(k, dtype=torch.int32)
    labels, rect, npix, nn = i32(P), i32(4 * P), torch.zeros(P), i32(1)
    qmp_quadtree_labels(crit, mask, hir, n, m, S, cond, thresh, labels, rect, npix, nn, None, None, None, None, None)
    N = int(nn[0])
    pix_ptr, pix_idx = i32(P + 1), i32(P)
    qmp_mesh_pixels_from_rects(labels, n, m, rect, npix, nn, P, pix_ptr, pix_idx, None, None)
    data_cap = torch.zeros(T, P, C)
    qmp_segment_sum(img, T, P, C, pix_ptr, pix_idx, npix, P, nn, 1, 0, data_cap)
    e_cap = 4 * P
    ei, s32, d32, ne = torch.zeros(2, e_cap, dtype=torch.int64), i32(e_cap), i32(e_cap), i32(1)
    qmp_adjacency_quadtree(labels, n, m, ei[0], ei[1], s32, d32, ne, None, None, 0, None, None, None, None)
    E = int(ne[0])
    attrs = torch.zeros(e_cap, 2) if two_cols else torch.zeros(e_cap)
    qmp_edge_attrs(s32, d32, e_cap, ne, data_cap[0, :, C - 2:], data_cap[0, :, C - 1:], C, m, n, resolution, two_cols, attrs)
    edge_index = ei[:, :E].contiguous()
    in_ptr, out_ptr = i32(N + 1), i32(N + 1)
    in_src, in_eid, out_dst, out_kin = i32(E), i32(E), i32(E), i32(E)
    qmp_csr_from_edge_index(edge_index, E, N, i32(E), i32(E), in_ptr, in_src, in_eid, out_ptr, out_dst, out_kin, None, None, None, None,
                            None)
    attrs = attrs[:E].contiguous()
    data = torch.cat([data_cap[:, :N], (npix[:N] / ((S / 2) ** 2)).reshape(1, N, 1).expand(T, N, 1)], -1).contiguous()
    im = flat(img, T * P * C)
    _GB[arena.data_ptr()] = dict(
        ints=[labels, pix_ptr[:N + 1], pix_idx, s32[:E], d32[:E], in_ptr, in_src, in_eid, out_ptr, out_dst, out_kin],
        floats=[npix[:N], data.reshape(-1), attrs.reshape(-1), attrs[in_eid.long()].reshape(-1)], edge_index=edge_index)
    host = (ctypes.c_int * 4).from_address(int(counts_host))
    host[0], host[1], host[2], host[3] = N, E, int(torch.isnan(im).sum()), 0


def qmp_quadtree_graph_export(arena, n, m, S, T, C, two_cols, N, E, ipack, fpack, edge_index):
    res = _GB[arena.data_ptr()]
    r4 = lambda v: (v + 3) & ~3
    for pack, segs in ((ipack, res["ints"]), (fpack, res["floats"])):
        off = 0
        for seg in segs:
            flat(pack)[off:off + seg.numel()].copy_(seg)
            off += r4(seg.numel())
    flat(edge_index, 2 * E).copy_(res["edge_index"].reshape(-1))


def qmp_gather_rows(inp, idx, n, width, out):
    flat(out, n * width).view(n, width).copy_(flat(inp, n * width).view(n, width)[flat(idx, n).long()])


# ------------------------------------------------------------------------------------------ dense
def qmp_gemm(A, B, bias, C, n, m, k, lda, ldb, ldc, sA, sB, sC, sBias, batch, b_is_kxm, accumulate, relu):
    fa, fb, fc = flat(A), flat(B), flat(C)
    for b in range(batch):
        a = win(fa, (n, k), (lda, 1), b * sA)
        bm = win(fb, (k, m), (ldb, 1), b * sB) if b_is_kxm else win(fb, (m, k), (ldb, 1), b * sB).T
        c = win(fc, (n, m), (ldc, 1), b * sC)
        v = a @ bm
        if bias is not None:
            v = v + flat(bias)[b * sBias: b * sBias + m]
        if accumulate:
            v = v + c
        if relu:
            v = torch.relu(v)
        c.copy_(v)


def qmp_gemm_tn_acc(A, B, C, n, ma, mb, lda, ldb, ldc, sA, sB, sC, batch, b_ones):
    fa, fb, fc = flat(A), flat(B), flat(C)
    real = mb - 1 if b_ones else mb
    for b in range(batch):
        a = win(fa, (n, ma), (lda, 1), b * sA)
        bm = win(fb, (n, real), (ldb, 1), b * sB)
        if b_ones:
            bm = torch.cat([bm, torch.ones(n, 1)], 1)
        c = win(fc, (ma, mb), (ldc, 1), b * sC)
        c.add_(a.T @ bm)


# ------------------------------------------------------------------------------------------ attention
def _attn_common(N, G_, D, ptr, nbr, ea, x, ldx, xoff):
    p = flat(ptr, N + 1).long()
    E = int(p[N])
    j = flat(nbr, E).long()
    i = torch.repeat_interleave(torch.arange(N), p[1:] - p[:-1])
    eattr = flat(ea, 2 * E).view(E, 2) if ea is not None else torch.zeros(E, 2)
    xs = [win(flat(x), (N, D), (ldx, 1), g * xoff) for g in range(G_)]
    return E, i, j, eattr, xs


def qmp_attn_fwd(N, G_, D, in_ptr, in_src, ea, x, ldx, xoff, U, Z, logit, mstat, linv, drop_p, seed):
    assert drop_p == 0.0, "emulation covers dropout = 0"
    E, i, j, eattr, xs = _attn_common(N, G_, D, in_ptr, in_src, ea, x, ldx, xoff)
    Uv = flat(U, N * G_ * (D + 2)).view(N, G_, D + 2)
    Zv = flat(Z, N * G_ * (D + 3)).view(N, G_, D + 3)
    lg, ms, li = flat(logit, max(E, 1) * G_).view(-1, G_), flat(mstat, N * G_).view(N, G_), flat(linv, N * G_).view(N, G_)
    for g in range(G_):
        s = (Uv[i, g, :D] * xs[g][j]).sum(-1) + (Uv[i, g, D:] * eattr).sum(-1)
        lg[:E, g] = s
        m = torch.full((N,), -math.inf).scatter_reduce(0, i, s, "amax", include_self=True)
        p = (s - m[i]).exp()
        l = torch.zeros(N).index_add(0, i, p)
        inv = torch.where(l > 0, 1 / l, torch.zeros_like(l))
        al = p * inv[i]
        Zv[:, g, :D] = torch.zeros(N, D).index_add(0, i, al[:, None] * xs[g][j])
        Zv[:, g, D:D + 2] = torch.zeros(N, 2).index_add(0, i, al[:, None] * eattr)
        Zv[:, g, D + 2] = torch.zeros(N).index_add(0, i, al)
        ms[:, g], li[:, g] = m, inv


def qmp_attn_bwd_target(N, G_, D, in_ptr, in_src, ea, x, ldx, xoff, logit, mstat, linv, dZ, ds, dU, drop_p, seed):
    E, i, j, eattr, xs = _attn_common(N, G_, D, in_ptr, in_src, ea, x, ldx, xoff)
    dZv = flat(dZ, N * G_ * (D + 3)).view(N, G_, D + 3)
    dUv = flat(dU, N * G_ * (D + 2)).view(N, G_, D + 2)
    lg, ms, li = flat(logit, max(E, 1) * G_).view(-1, G_), flat(mstat, N * G_).view(N, G_), flat(linv, N * G_).view(N, G_)
    dsv = flat(ds, max(E, 1) * G_).view(-1, G_)
    for g in range(G_):
        al = (lg[:E, g] - ms[i, g]).exp() * li[i, g]
        dal = (dZv[i, g, :D] * xs[g][j]).sum(-1) + (dZv[i, g, D:D + 2] * eattr).sum(-1) + dZv[i, g, D + 2]
        t = torch.zeros(N).index_add(0, i, al * dal)
        d = al * (dal - t[i])
        dsv[:E, g] = d
        dUv[:, g, :D] = torch.zeros(N, D).index_add(0, i, d[:, None] * xs[g][j])
        dUv[:, g, D:] = torch.zeros(N, 2).index_add(0, i, d[:, None] * eattr)


def qmp_attn_bwd_source(N, G_, D, out_ptr, out_dst, out_kin, logit, mstat, linv, ds, dZ, U, dx, lddx, dxoff, shared,
                        accumulate, drop_p, seed):
    p = flat(out_ptr, N + 1).long()
    E = int(p[N])
    i, kin = flat(out_dst, E).long(), flat(out_kin, E).long()
    j = torch.repeat_interleave(torch.arange(N), p[1:] - p[:-1])
    dZv = flat(dZ, N * G_ * (D + 3)).view(N, G_, D + 3)
    Uv = flat(U, N * G_ * (D + 2)).view(N, G_, D + 2)
    lg, ms, li = flat(logit, max(E, 1) * G_).view(-1, G_), flat(mstat, N * G_).view(N, G_), flat(linv, N * G_).view(N, G_)
    dsv = flat(ds, max(E, 1) * G_).view(-1, G_)
    total = torch.zeros(N, D)
    for g in range(G_):
        al = (lg[kin, g] - ms[i, g]).exp() * li[i, g]
        contrib = torch.zeros(N, D).index_add(0, j, al[:, None] * dZv[i, g, :D] + dsv[kin, g][:, None] * Uv[i, g, :D])
        if shared:
            total += contrib
        else:
            dst = win(flat(dx), (N, D), (lddx, 1), g * dxoff)
            dst.copy_(dst + contrib if accumulate else contrib)
    if shared:
        dst = win(flat(dx), (N, D), (lddx, 1), 0)
        dst.copy_(dst + total if accumulate else total)


# ------------------------------------------------------------------------------------------ GCN / Cheb
def qmp_edge_norm(mode, N, in_ptr, in_src, out_ptr, out_dst, out_kin, w, dis, val):
    p = flat(in_ptr, N + 1).long()
    E = int(p[N])
    j = flat(in_src, E).long()
    i = torch.repeat_interleave(torch.arange(N), p[1:] - p[:-1])
    wk = flat(w, E) if w is not None else torch.ones(E)
    if mode == 0:
        deg = torch.zeros(N).index_add(0, i, wk)
    else:
        deg = torch.zeros(N).index_add(0, j, torch.where(i == j, torch.zeros(E), wk))
    r = deg.pow(-0.5)
    r = r.masked_fill(r == math.inf, 0.0)
    v = r[j] * wk * r[i]
    if mode == 1:
        v = torch.where(i == j, torch.zeros(E), -v)
    flat(val, E).copy_(v)


def qmp_spmm(N, width, ptr, nbr, vidx, val, x, ldx, alpha, beta, z, ldz, y, ldy):
    p = flat(ptr, N + 1).long()
    E = int(p[N])
    nb = flat(nbr, E).long()
    row = torch.repeat_interleave(torch.arange(N), p[1:] - p[:-1])
    v = flat(val, E)[flat(vidx, E).long()] if vidx is not None else flat(val, E)
    xv = rows(x, N, ldx, width)
    out = alpha * torch.zeros(N, width).index_add(0, row, v[:, None] * xv[nb])
    if z is not None:
        out = out + beta * rows(z, N, ldz, width)
    rows(y, N, ldy, width).copy_(out)


def qmp_cheb_cell_fwd(N, F, C, K, S, cheb, in_ptr, in_src, val, X, H, p0, p1, p2, p3, b0, b1, b2, b3, ws, P):
    """csrc/cheb_cell.cu issues the launch sequence that cheb_cell._fwd_py restates (on the emulated kernels here)."""
    from quadtree_mpnnlstm_b200 import cheb_cell as CC
    packs = [p for p in (p0, p1, p2, p3) if p is not None]
    biases = [b for b in (b0, b1, b2, b3) if b is not None]
    CC._fwd_py(N, F, C, K, S, bool(cheb), (in_ptr, in_src, val, None, None, None), X.reshape(-1), H.reshape(-1), packs, biases, ws,
               P.view(-1))


def qmp_cheb_cell_bwd(N, F, C, K, S, cheb, out_ptr, out_dst, out_kin, val, dP, p0, p1, p2, p3, a0, a1, a2, a3, ws, ws2, need_dx,
                      need_dh, dX, dH):
    from quadtree_mpnnlstm_b200 import cheb_cell as CC
    packs = [p for p in (p0, p1, p2, p3) if p is not None]
    accs = [a for a in (a0, a1, a2, a3) if a is not None]
    CC._bwd_py(N, F, C, K, S, bool(cheb), (None, None, val, out_ptr, out_dst, out_kin), dP.reshape(-1), packs, accs, ws, ws2,
               bool(need_dx), bool(need_dh), dX.view(-1) if dX is not None else None, dH.view(-1) if dH is not None else None)


def qmp_cheb_stack_fwd(N, K, cheb, L, w0, M0, M1, M2, r0, r1, r2, in_ptr, in_src, val, X, p0, p1, p2, b0, b1, b2, ws, out):
    from quadtree_mpnnlstm_b200 import cheb_cell as CC
    CC._stack_fwd_py(N, K, bool(cheb), w0, [M0, M1, M2][:L], [r0, r1, r2][:L], (in_ptr, in_src, val, None, None, None), X.reshape(-1),
                     [p0, p1, p2][:L], [b0, b1, b2][:L], ws, out.view(-1))


def qmp_cheb_stack_bwd(N, K, cheb, L, w0, M0, M1, M2, r0, r1, r2, out_ptr, out_dst, out_kin, val, dOut, out_last, p0, p1, p2, a0, a1,
                       a2, ws, ws2, need_dx, dX):
    from quadtree_mpnnlstm_b200 import cheb_cell as CC
    CC._stack_bwd_py(N, K, bool(cheb), w0, [M0, M1, M2][:L], [r0, r1, r2][:L], (None, None, val, out_ptr, out_dst, out_kin),
                     dOut.reshape(-1), out_last.reshape(-1), [p0, p1, p2][:L], [a0, a1, a2][:L], ws, ws2, bool(need_dx),
                     dX.view(-1) if dX is not None else None)


# ------------------------------------------------------------------------------------------ LSTM gates
def _ln(x, gamma, beta, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    rstd = (var + eps).rsqrt()
    xh = (x - mu) * rstd
    return xh * gamma + beta, xh, rstd


def _ln_bwd(xh, dy, gamma, rstd):
    g = dy * gamma
    return rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))


def qmp_lstm_gates_fwd(N, C, P, ldp, Cprev, params, norm_h, norm_c, norm_o, eps, gates, Craw, Oout, Hout, Cout, head_in,
                       ldh, concat):
    Pv = rows(P, N, ldp, 4 * C)
    prm = flat(params, 13 * C).view(13, C)
    cp = flat(Cprev, N * C).view(N, C) if Cprev is not None else torch.zeros(N, C)
    I = torch.sigmoid(Pv[:, :C] + prm[0] * cp + prm[3])
    Fg = torch.sigmoid(Pv[:, C:2 * C] + prm[1] * cp + prm[4])
    T = torch.tanh(Pv[:, 2 * C:3 * C] + prm[5])
    Cn = Fg * cp + I * T
    O = torch.sigmoid(Pv[:, 3 * C:] + prm[2] * Cn + prm[6])
    H = O * torch.tanh(Cn)
    flat(gates, N * 4 * C).view(N, 4 * C).copy_(torch.cat([I, Fg, T, O], 1))
    flat(Craw, N * C).view(N, C).copy_(Cn)
    if Oout is not None:
        flat(Oout, N * C).view(N, C).copy_(O)
    flat(Hout, N * C).view(N, C).copy_(_ln(H, prm[7], prm[8], eps)[0] if norm_h else H)
    flat(Cout, N * C).view(N, C).copy_(_ln(Cn, prm[9], prm[10], eps)[0] if norm_c else Cn)
    if head_in is not None:
        hv = rows(head_in, N, ldh, C + 1)
        hv[:, :C] = torch.relu(_ln(O, prm[11], prm[12], eps)[0] if norm_o else O)
        if concat is not None:
            hv[:, C] = flat(concat, N)


def qmp_lstm_gates_bwd(N, C, gates, Craw, Cprev, params, norm_h, norm_c, norm_o, eps, dHout, dCout, dOdirect, dHead, lddh,
                       dP, lddp, dCprev, dparams):
    gt = flat(gates, N * 4 * C).view(N, 4 * C)
    I, Fg, T, O = gt[:, :C], gt[:, C:2 * C], gt[:, 2 * C:3 * C], gt[:, 3 * C:]
    Cn = flat(Craw, N * C).view(N, C)
    cp = flat(Cprev, N * C).view(N, C) if Cprev is not None else torch.zeros(N, C)
    prm = flat(params, 13 * C).view(13, C)
    z = lambda t: flat(t, N * C).view(N, C).clone() if t is not None else torch.zeros(N, C)
    dH, dC, dO = z(dHout), z(dCout), z(dOdirect)
    dprm = torch.zeros(13, C)
    tc = torch.tanh(Cn)
    if norm_h:
        _, xh, rstd = _ln(O * tc, prm[7], prm[8], eps)
        dprm[7], dprm[8] = (dH * xh).sum(0), dH.sum(0)
        dH = _ln_bwd(xh, dH, prm[7], rstd)
    if norm_c:
        _, xh, rstd = _ln(Cn, prm[9], prm[10], eps)
        dprm[9], dprm[10] = (dC * xh).sum(0), dC.sum(0)
        dC = _ln_bwd(xh, dC, prm[9], rstd)
    if dHead is not None:
        dh = rows(dHead, N, lddh, C)
        if norm_o:
            y, xh, rstd = _ln(O, prm[11], prm[12], eps)
            dy = torch.where(y > 0, dh, torch.zeros_like(dh))
            dprm[11], dprm[12] = (dy * xh).sum(0), dy.sum(0)
            dO = dO + _ln_bwd(xh, dy, prm[11], rstd)
        else:
            dO = dO + torch.where(O > 0, dh, torch.zeros_like(dh))
    dOt = dH * tc + dO
    dOp = dOt * O * (1 - O)
    dCn = dC + dH * O * (1 - tc * tc) + dOp * prm[2]
    dIp = dCn * T * I * (1 - I)
    dFp = dCn * cp * Fg * (1 - Fg)
    dTp = dCn * I * (1 - T * T)
    rows(dP, N, lddp, 4 * C).copy_(torch.cat([dIp, dFp, dTp, dOp], 1))
    if dCprev is not None:
        flat(dCprev, N * C).view(N, C).copy_(dCn * Fg + dIp * prm[0] + dFp * prm[1])
    dprm[0], dprm[1], dprm[2] = (dIp * cp).sum(0), (dFp * cp).sum(0), (dOp * Cn).sum(0)
    dprm[3], dprm[4], dprm[5], dprm[6] = dIp.sum(0), dFp.sum(0), dTp.sum(0), dOp.sum(0)
    if dparams is not None:
        flat(dparams, 13 * C).view(13, C).add_(dprm)


def qmp_gru_gates1_fwd(n, az, bz, ar, br, H, Z, R, HR):
    z, r = torch.sigmoid(flat(az, n) + flat(bz, n)), torch.sigmoid(flat(ar, n) + flat(br, n))
    flat(Z, n).copy_(z)
    flat(R, n).copy_(r)
    flat(HR, n).copy_(flat(H, n) * r)


def qmp_gru_gates1_bwd(n, Z, R, H, dZ, dR, dHR, dpz, dpr, dH):
    z, r, h = flat(Z, n), flat(R, n), flat(H, n)
    g = lambda t: flat(t, n) if t is not None else torch.zeros(n)
    flat(dpz, n).copy_(g(dZ) * z * (1 - z))
    flat(dpr, n).copy_((g(dR) + g(dHR) * h) * r * (1 - r))
    flat(dH, n).copy_(g(dHR) * r)


def qmp_gru_gates2_fwd(n, ah, bh, Z, H, Ht, Hn):
    t = torch.tanh(flat(ah, n) + flat(bh, n))
    flat(Ht, n).copy_(t)
    flat(Hn, n).copy_(flat(Z, n) * flat(H, n) + (1 - flat(Z, n)) * t)


def qmp_gru_gates2_bwd(n, Z, H, Ht, dHn, dph, dZ, dH):
    z, h, t, g = flat(Z, n), flat(H, n), flat(Ht, n), flat(dHn, n)
    flat(dZ, n).copy_(g * (h - t))
    flat(dH, n).copy_(g * z)
    flat(dph, n).copy_(g * (1 - z) * (1 - t * t))


def qmp_head_finish_fwd(y, x, N, F, binary, drop_p, seed, out, x_next):
    xv = flat(x, N * F).view(N, F)
    o = torch.tanh(flat(y, N)) + xv[:, 0]
    if binary:
        o = torch.sigmoid(o)
    flat(out, N).copy_(o)
    if x_next is not None:
        xn = flat(x_next, N * F).view(N, F)
        xn.copy_(xv)
        xn[:, 0] = o


def qmp_head_finish_bwd(y, out, x, d_out, d_xnext, N, F, binary, drop_p, seed, dy, dx):
    g = torch.zeros(N)
    if d_out is not None:
        g = g + flat(d_out, N)
    if d_xnext is not None:
        g = g + flat(d_xnext, N * F).view(N, F)[:, 0]
    if binary:
        o = flat(out, N)
        g = g * o * (1 - o)
    th = torch.tanh(flat(y, N))
    flat(dy, N).copy_(g * (1 - th * th))
    rows = flat(dx, N * F).view(N, F)
    rows.copy_(flat(d_xnext, N * F).view(N, F) if d_xnext is not None else torch.zeros(N, F))
    rows[:, 0] = g


def qmp_relu_mask_to(y, g, out, n):
    flat(out, n).copy_(torch.where(flat(y, n) > 0, flat(g, n), torch.zeros(n)))


def qmp_relu_mask(y, dy, n):
    d = flat(dy, n)
    d[~(flat(y, n) > 0)] = 0.0


# ------------------------------------------------------------------------------------------ fused kernels
_FC = 32


def _unpack_fwd(w, G, DC):
    w = flat(w, G * ((DC + 2) * DC + DC + 4 + _FC * (DC + 4) + _FC * DC + _FC)).view(G, -1)
    o1 = (DC + 2) * DC
    o2 = o1 + DC + 4
    o3 = o2 + _FC * (DC + 4)
    o4 = o3 + _FC * DC
    return (w[:, :o1].view(G, DC + 2, DC), w[:, o1:o2], w[:, o2:o3].view(G, _FC, DC + 4), w[:, o3:o4].view(G, _FC, DC), w[:, o4:])


def _unpack_bwd(w, G, DC):
    tot = (DC + 2) * DC + DC + 4 + DC * (DC + 4) + (DC + 4) * _FC + DC * _FC
    w = flat(w, G * tot).view(G, -1)
    o1 = (DC + 2) * DC
    o2 = o1 + DC + 4
    o3 = o2 + DC * (DC + 4)
    o4 = o3 + (DC + 4) * _FC
    W1, b1 = w[:, :o1].view(G, DC + 2, DC), w[:, o1:o2]
    W1T, W2T, W3T = w[:, o2:o3].view(G, DC, DC + 4), w[:, o3:o4].view(G, DC + 4, _FC), w[:, o4:].view(G, DC, _FC)
    assert torch.equal(W1T[:, :, :DC + 2], W1.transpose(1, 2)), "backward pack: W1T is not the transpose of W1"
    return W1, b1, W2T.transpose(1, 2), W3T.transpose(1, 2)            # W1, b1, W2 [G,32,DC+4], W3 [G,32,DC]


def _cap(D, small):
    return (4 if D <= 4 else 8) if small else (32 if D <= 32 else 36)


def _rows_pad(x, N, ld, off, D, DC):
    v = win(flat(x), (N, D), (ld, 1), off)
    return torch.cat([v, torch.zeros(N, DC - D)], 1)


def _fused_convs(xa, lda, DA, GA, xb, ldb, DB, GB, sharedB):
    """(conv index, segment, index in segment, x pointer, ld, offset, D, cap)."""
    out = []
    for g in range(GA):
        out.append((g, 0, g, xa, lda, 0, DA, _cap(DA, True)))
    for g in range(GB):
        out.append((GA + g, 1, g, xb, ldb, 0 if sharedB else g * DB, DB, _cap(DB, False)))
    return out


def _edge_lists(N, ptr, nbr):
    p = flat(ptr, N + 1).long()
    E = int(p[N])
    return E, torch.repeat_interleave(torch.arange(N), p[1:] - p[:-1]), flat(nbr, E).long()


def qmp_fused_fwd(N, in_ptr, in_src, ea, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, mode, relu_out, C, out, ldo, Cprev,
                  params, norm_h, norm_c, norm_o, eps, gates, Craw, Oout, Hout, Cout, head_in, ldh, concat, logit, mstat, linv,
                  drop_p, seed):
    assert drop_p == 0.0
    NC = GA + GB
    E, ti, sj = _edge_lists(N, in_ptr, in_src)
    eattr = flat(ea, 2 * E).view(E, 2) if ea is not None else torch.zeros(E, 2)
    packs = {0: _unpack_fwd(wa, GA, _cap(DA, True)) if GA else None, 1: _unpack_fwd(wb, GB, _cap(DB, False))}
    lg, ms, li = flat(logit, max(E, 1) * NC).view(-1, NC), flat(mstat, N * NC).view(N, NC), flat(linv, N * NC).view(N, NC)
    nslots = 4 if mode == 1 else NC
    P = torch.zeros(N, nslots, _FC)
    for (c, seg, g, xp, ld, off, D, DC) in _fused_convs(xa, lda, DA, GA, xb, ldb, DB, GB, sharedB):
        W1, b1, W2, W3, b3 = (t[g] for t in packs[seg])
        x = _rows_pad(xp, N, ld, off, D, DC)
        u = x @ W1[:DC].T + b1[:DC]
        w = x @ W1[DC:].T + b1[DC:DC + 2]
        s = (u[ti] * x[sj]).sum(-1) + (w[ti] * eattr).sum(-1)
        lg[:E, c] = s
        m = torch.full((N,), -math.inf).scatter_reduce(0, ti, s, "amax", include_self=True)
        p = (s - m[ti]).exp()
        l = torch.zeros(N).index_add(0, ti, p)
        inv = torch.where(l > 0, 1 / l, torch.zeros_like(l))
        al = p * inv[ti]
        z = torch.zeros(N, DC + 4)
        z[:, :DC] = torch.zeros(N, DC).index_add(0, ti, al[:, None] * x[sj])
        z[:, DC:DC + 2] = torch.zeros(N, 2).index_add(0, ti, al[:, None] * eattr)
        z[:, DC + 2] = torch.zeros(N).index_add(0, ti, al)
        ms[:, c], li[:, c] = m, inv
        slot = (c if c < GA else (c - GA) % 4) if mode == 1 else c
        P[:, slot] += z @ W2.T + x @ W3.T + b3
    if mode == 1:
        Pf = P.reshape(N, 4 * _FC).contiguous()
        qmp_lstm_gates_fwd(N, _FC, Pf, 4 * _FC, Cprev, params, norm_h, norm_c, norm_o, eps, gates, Craw, Oout, Hout, Cout,
                           head_in, ldh, concat)
        if head_in is not None and ldh > _FC + 1:
            rows(head_in, N, ldh, ldh)[:, _FC + 1:] = 0.0
    else:
        o = torch.relu(P) if relu_out else P
        win(flat(out), (N, NC, C), (ldo, C, 1)).copy_(o[:, :, :C])


def _dP_of(dP, lddp, N, mode, C, c, GA):
    if mode == 1:
        slot = c if c < GA else (c - GA) % 4
        return win(flat(dP), (N, _FC), (lddp, 1), slot * _FC)
    v = win(flat(dP), (N, C), (lddp, 1), c * C)
    return torch.cat([v, torch.zeros(N, _FC - C)], 1)


def qmp_fused_bwd_target(N, in_ptr, in_src, ea, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, mode, C, dP, lddp, logit, mstat,
                         linv, ds, ZsA, dUsA, ZsB, dUsB, dxa, dxb, drop_p, seed):
    NC = GA + GB
    E, ti, sj = _edge_lists(N, in_ptr, in_src)
    eattr = flat(ea, 2 * E).view(E, 2) if ea is not None else torch.zeros(E, 2)
    packs = {0: _unpack_bwd(wa, GA, _cap(DA, True)) if GA else None, 1: _unpack_bwd(wb, GB, _cap(DB, False))}
    lg, ms, li = flat(logit, max(E, 1) * NC).view(-1, NC), flat(mstat, N * NC).view(N, NC), flat(linv, N * NC).view(N, NC)
    dsv = flat(ds, max(E, 1) * NC).view(-1, NC)
    first = {0: True, 1: True}
    for (c, seg, g, xp, ld, off, D, DC) in _fused_convs(xa, lda, DA, GA, xb, ldb, DB, GB, sharedB):
        W1, b1, W2, W3 = (t[g] for t in packs[seg])
        G = GA if seg == 0 else GB
        x = _rows_pad(xp, N, ld, off, D, DC)
        g32 = _dP_of(dP, lddp, N, mode, C, c, GA)
        dz = g32 @ W2                                              # [N, DC+4]
        al = (lg[:E, c] - ms[ti, c]).exp() * li[ti, c]
        dal = (dz[ti, :DC] * x[sj]).sum(-1) + (dz[ti, DC:DC + 2] * eattr).sum(-1) + dz[ti, DC + 2]
        t = torch.zeros(N).index_add(0, ti, al * dal)
        d = al * (dal - t[ti])
        dsv[:E, c] = d
        du = torch.zeros(N, DC + 4)
        du[:, :DC] = torch.zeros(N, DC).index_add(0, ti, d[:, None] * x[sj])
        du[:, DC:DC + 2] = torch.zeros(N, 2).index_add(0, ti, d[:, None] * eattr)
        z = torch.zeros(N, DC + 4)
        z[:, :DC] = torch.zeros(N, DC).index_add(0, ti, al[:, None] * x[sj])
        z[:, DC:DC + 2] = torch.zeros(N, 2).index_add(0, ti, al[:, None] * eattr)
        z[:, DC + 2] = torch.zeros(N).index_add(0, ti, al)
        Zs, dUs = (ZsA, dUsA) if seg == 0 else (ZsB, dUsB)
        win(flat(Zs), (N, DC + 4), (G * (DC + 4), 1), g * (DC + 4)).copy_(z)
        win(flat(dUs), (N, DC + 4), (G * (DC + 4), 1), g * (DC + 4)).copy_(du)
        dxp = dxa if seg == 0 else dxb
        if dxp is not None:
            dxs = (g32 @ W3 + du[:, :DC + 2] @ W1)[:, :D]
            dst = win(flat(dxp), (N, D), (ld, 1), off)
            shared = seg == 0 or sharedB
            if shared and not first[seg]:
                dst.add_(dxs)
            else:
                dst.copy_(dxs)
            first[seg] = False


def qmp_fused_bwd_source(N, out_ptr, out_dst, out_kin, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, mode, C, dP, lddp, logit,
                         mstat, linv, ds, dxa, dxb, drop_p, seed):
    NC = GA + GB
    p = flat(out_ptr, N + 1).long()
    E = int(p[N])
    ti, kin = flat(out_dst, E).long(), flat(out_kin, E).long()
    sj = torch.repeat_interleave(torch.arange(N), p[1:] - p[:-1])
    packs = {0: _unpack_bwd(wa, GA, _cap(DA, True)) if GA else None, 1: _unpack_bwd(wb, GB, _cap(DB, False))}
    lg, ms, li = flat(logit, max(E, 1) * NC).view(-1, NC), flat(mstat, N * NC).view(N, NC), flat(linv, N * NC).view(N, NC)
    dsv = flat(ds, max(E, 1) * NC).view(-1, NC)
    for (c, seg, g, xp, ld, off, D, DC) in _fused_convs(xa, lda, DA, GA, xb, ldb, DB, GB, sharedB):
        dxp = dxa if seg == 0 else dxb
        if dxp is None:
            continue
        W1, b1, W2, W3 = (t[g] for t in packs[seg])
        x = _rows_pad(xp, N, ld, off, D, DC)
        g32 = _dP_of(dP, lddp, N, mode, C, c, GA)
        al = (lg[kin, c] - ms[ti, c]).exp() * li[ti, c]
        a = torch.zeros(N, _FC).index_add(0, sj, al[:, None] * g32[ti])
        b = torch.zeros(N, DC).index_add(0, sj, dsv[kin, c][:, None] * x[ti])
        sds = torch.zeros(N).index_add(0, sj, dsv[kin, c])
        contrib = a @ W2[:, :DC] + b @ W1[:DC].T + sds[:, None] * b1[:DC]
        win(flat(dxp), (N, D), (ld, 1), off).add_(contrib[:, :D])


def qmp_fused_pack_tc(pack, G, DC, which, out):
    """Emulated image = the raw pack bytes at the start of each image row (the emulated kernels unpack it again)."""
    total = (DC + 2) * DC + DC + 4 + _FC * (DC + 4) + _FC * DC + _FC
    src = flat(pack, G * total).view(G, total)
    rows = out.view(G, -1)
    rows[:, :total * 4] = src.contiguous().view(torch.uint8).view(G, total * 4)


def _pack_from_image(img, G, DC):
    if img is None:
        return None
    total = (DC + 2) * DC + DC + 4 + _FC * (DC + 4) + _FC * DC + _FC
    return img.view(G, -1)[:, :total * 4].contiguous().view(torch.float32).view(G, total)


def qmp_fused_fwd_tc(N, in_ptr, in_src, ea, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, *rest):
    qmp_fused_fwd(N, in_ptr, in_src, ea, xa, lda, DA, GA, _pack_from_image(wa, GA, _cap(DA, True)) if GA else None, xb, ldb, DB,
                  GB, sharedB, _pack_from_image(wb, GB, _cap(DB, False)), *rest)


def _total(DC):
    return (DC + 2) * DC + DC + 4 + _FC * (DC + 4) + _FC * DC + _FC


def qmp_fused_pack_cell(packA, packB, out):
    """Emulated cell image = the raw bytes of the two packs (the emulated kernel unpacks them again)."""
    na, nb = 4 * _total(4) * 4, 4 * _total(32) * 4
    out[:na] = flat(packA, 4 * _total(4)).contiguous().view(torch.uint8)
    out[na:na + nb] = flat(packB, 4 * _total(32)).contiguous().view(torch.uint8)


def qmp_fused_cell_fwd(N, in_ptr, in_src, ea, xa, lda, xb, ldb, image, Cprev, params, norm_h, norm_c, norm_o, eps, gates, Craw,
                       Oout, Hout, Cout, head_in, ldh, concat, logit, mstat, linv, usave, drop_p, seed):
    na, nb = 4 * _total(4) * 4, 4 * _total(32) * 4
    wa = image[:na].contiguous().view(torch.float32).view(4, _total(4))
    wb = image[na:na + nb].contiguous().view(torch.float32).view(4, _total(32))
    qmp_fused_fwd(N, in_ptr, in_src, ea, xa, lda, 4, 4, wa, xb, ldb, 32, 4, 1, wb, 1, 0, _FC, None, 8 * _FC, Cprev, params, norm_h,
                  norm_c, norm_o, eps, gates, Craw, Oout, Hout, Cout, head_in, ldh, concat, logit, mstat, linv, drop_p, seed)
    if usave is not None:
        W1, b1 = _unpack_fwd(wb, 4, 32)[:2]
        h = rows(xb, N, ldb, 32)
        flat(usave, N * 128).view(N, 4, 32).copy_(torch.stack([h @ W1[g, :32].T + b1[g, :32] for g in range(4)], 1))


def qmp_pack_head_bwd(pack, out):
    qmp_fused_pack_tc(pack, 1, 36, 1, out)


def qmp_head_bwd(N, in_ptr, in_src, ea, x, ldx, image, g, ldg, logit, mstat, linv, Zs, dUs, dx, drop_p, seed):
    E, _, _ = _edge_lists(N, in_ptr, in_src)
    qmp_fused_bwd_onepass_tc(N, in_ptr, in_src, ea, None, 0, 0, 0, None, x, ldx, 36, 1, 1, image, 0, _FC, g, ldg, logit, mstat, linv,
                             torch.zeros(max(E, 1), 1), None, None, Zs, dUs, None, dx, drop_p, seed)


def qmp_fused_pack_cell_bwd(packA, packB, out):
    qmp_fused_pack_cell(packA, packB, out)


def qmp_fused_cell_bwd(N, in_ptr, in_src, ea, xa, lda, xb, ldb, image, usave, dP, lddp, gates, Craw, Cprev, params, norm_h, norm_c,
                       norm_o, eps, dHout, dCout, dOdirect, dHead, lddh, dCprev, dparams, logit, mstat, linv, zB, duB, sd, sg,
                       dxa, dxb, drop_p, seed):
    """Target side by the emulated per-conv kernel; source side of every edge added from the same in-CSR edge list; the rows
    for the weight gradients repacked into the panel layout of csrc/cell_wgrad.cu."""
    if gates is not None:       # the fused gate backward: dP is an output
        qmp_lstm_gates_bwd(N, _FC, gates, Craw, Cprev, params, norm_h, norm_c, norm_o, eps, dHout, dCout, dOdirect, dHead, lddh, dP,
                           lddp, dCprev, dparams)
    ZsA, dUsA, ZsB, dUsB = torch.zeros(N, 4, 8), torch.zeros(N, 4, 8), torch.zeros(N, 4, 36), torch.zeros(N, 4, 36)
    _cell_bwd_old_layout(N, in_ptr, in_src, ea, xa, lda, xb, ldb, image, usave, dP, lddp, logit, mstat, linv, ZsA, dUsA, ZsB, dUsB,
                         dxa, dxb, drop_p, seed)
    flat(zB, N * 128).view(N, 4, 32).copy_(ZsB[:, :, :32])
    flat(duB, N * 128).view(N, 4, 32).copy_(dUsB[:, :, :32])
    sdv, sgv = flat(sd, N * 64).view(N, 64), flat(sg, N * 32).view(N, 32)
    sdv.zero_()
    sdv[:, :4] = rows(xa, N, lda, 4)
    sdv[:, 4] = 1.0
    sdv[:, 8:24].view(N, 4, 4)[:, :, :3] = ZsB[:, :, 32:35]
    sdv[:, 32:].view(N, 4, 8)[:, :, :7] = ZsA[:, :, :7]
    sgv[:, :8] = dUsB[:, :, 32:34].reshape(N, 8)
    sgv[:, 8:24] = dUsA[:, :, :4].reshape(N, 16)
    sgv[:, 24:] = dUsA[:, :, 4:6].reshape(N, 8)


def qmp_cell_wgrad(N, h, ldh, dP, lddp, zB, duB, sd, sg, gwa, gwb):
    """Reductions over the nodes from the panel layout (csrc/cell_wgrad.cu) into the two padded packs."""
    g = rows(dP, N, lddp, 128).view(N, 4, 32)
    H = rows(h, N, ldh, 32)
    z, du = flat(zB, N * 128).view(N, 4, 32), flat(duB, N * 128).view(N, 4, 32)
    sdv, sgv = flat(sd, N * 64).view(N, 64), flat(sg, N * 32).view(N, 32)
    x, one = sdv[:, :4], sdv[:, 4]
    ga, gb = flat(gwa, 4 * _total(4)).view(4, -1), flat(gwb, 4 * _total(32)).view(4, -1)
    for c in range(4):
        gc = g[:, c]
        # H conv c
        Zh = torch.cat([z[:, c], sdv[:, 8 + 4 * c:12 + 4 * c]], 1)                  # [N, 36]
        dUh = torch.cat([du[:, c], sgv[:, 2 * c:2 * c + 2]], 1)                     # [N, 34]
        o1, o2, o3, o4 = 34 * 32, 34 * 32 + 36, 34 * 32 + 36 + 32 * 36, 34 * 32 + 36 + 32 * 36 + 32 * 32
        gb[c, :o1] += (dUh.T @ H).reshape(-1)
        gb[c, o1:o1 + 34] += (dUh * one[:, None]).sum(0)
        gb[c, o2:o3] += (gc.T @ Zh).reshape(-1)
        gb[c, o3:o4] += (gc.T @ H).reshape(-1)
        gb[c, o4:] += (gc * one[:, None]).sum(0)
        # X conv c
        Zx = sdv[:, 32 + 8 * c:40 + 8 * c]
        dUx = torch.cat([sgv[:, 8 + 4 * c:12 + 4 * c], sgv[:, 24 + 2 * c:26 + 2 * c]], 1)     # [N, 6]
        o1, o2, o3, o4 = 24, 32, 32 + 256, 32 + 256 + 128
        ga[c, :o1] += (dUx.T @ x).reshape(-1)
        ga[c, o1:o1 + 6] += (dUx * one[:, None]).sum(0)
        ga[c, o2:o3] += (gc.T @ Zx).reshape(-1)
        ga[c, o3:o4] += (gc.T @ x).reshape(-1)
        ga[c, o4:] += (gc * one[:, None]).sum(0)


def qmp_fused_wgrad_tma(*args):
    qmp_fused_wgrad(*args)


def qmp_panel_wgrad(N, x, ldx, D, DC, g, ldg, Zs, dUs, gw):
    assert DC == 36
    qmp_fused_wgrad(N, None, 0, 0, 0, x, ldx, D, 1, 1, 0, _FC, g, ldg, None, None, Zs, dUs, None, gw)


def _cell_bwd_old_layout(N, in_ptr, in_src, ea, xa, lda, xb, ldb, image, usave, dP, lddp, logit, mstat, linv, ZsA, dUsA, ZsB, dUsB,
                         dxa, dxb, drop_p, seed):
    na, nb = 4 * _total(4) * 4, 4 * _total(32) * 4
    pa = _bwd_pack_from_image(image[:na].view(4, -1), 4, 4)
    pb = _bwd_pack_from_image(image[na:na + nb].view(4, -1), 4, 32)
    E, ti, sj = _edge_lists(N, in_ptr, in_src)
    ds = torch.zeros(max(E, 1), 8)
    qmp_fused_bwd_target(N, in_ptr, in_src, ea, xa, lda, 4, 4, pa, xb, ldb, 32, 4, 1, pb, 1, _FC, dP, lddp, logit, mstat, linv, ds, ZsA,
                         dUsA, ZsB, dUsB, dxa, dxb, drop_p, seed)
    lg, ms, li = flat(logit, max(E, 1) * 8).view(-1, 8), flat(mstat, N * 8).view(N, 8), flat(linv, N * 8).view(N, 8)
    us = flat(usave, N * 128).view(N, 4, 32)
    packs = {0: _unpack_bwd(pa, 4, 4), 1: _unpack_bwd(pb, 4, 32)}
    for (c, seg, g, xp, ld, off, D, DC) in _fused_convs(xa, lda, 4, 4, xb, ldb, 32, 4, 1):
        W1, b1, W2, W3 = (t[g] for t in packs[seg])
        x = _rows_pad(xp, N, ld, off, D, DC)
        g32 = _dP_of(dP, lddp, N, 1, _FC, c, 4)
        al = (lg[:E, c] - ms[ti, c]).exp() * li[ti, c]
        u = x @ W1[:DC].T + b1[:DC]
        if seg == 1:
            assert torch.allclose(us[:, g], u, rtol=1e-4, atol=1e-5), "usave must hold the logit projections of the H convs"
        contrib = al[:, None] * (g32 @ W2[:, :DC])[ti] + ds[:E, c][:, None] * u[ti]
        dst = win(flat(dxa if seg == 0 else dxb), (N, D), (ld, 1), off)
        dst.add_(torch.zeros(N, DC).index_add(0, sj, contrib)[:, :D])


def _tconv1_common(N, in_ptr, in_src, ea, x, ldx, P, drop_p):
    assert drop_p == 0.0
    E, ti, sj = _edge_lists(N, in_ptr, in_src)
    X = rows(x, N, ldx, 32)
    p = flat(P, 136)
    W4, b, we = p[:128].view(4, 32), p[128:132], p[132:134]
    S = X @ W4.T + b
    EA = flat(ea, 2 * E).view(E, 2) if ea is not None else torch.zeros(E, 2)
    e = EA @ we
    s = S[ti, 0] * (S[sj, 1] + e)
    m = torch.full((N,), -float("inf")).scatter_reduce(0, ti, s, "amax", include_self=True)
    pe = torch.exp(s - m[ti])
    l = torch.zeros(N).index_add(0, ti, pe)
    al = pe / l[ti]
    return E, ti, sj, X, W4, S, EA, e, al


def qmp_head_tail_fwd(N, in_ptr, in_src, ea, h, ldh, P, x, F, binary, drop_attn, seed_attn, drop_out, seed_out, s4, y, out, x_next):
    qmp_tconv1_fwd(N, in_ptr, in_src, ea, h, ldh, P, s4, y, drop_attn, seed_attn)
    qmp_head_finish_fwd(y, x, N, F, binary, drop_out, seed_out, out, x_next)


def qmp_head_tail_bwd(N, in_ptr, in_src, ea, h, ldh, P, s4, y, out, x, F, binary, drop_attn, seed_attn, drop_out, seed_out, d_out, d_xnext,
                      ds4, dh, lddh, relu_mask, dx, gP):
    dy = torch.zeros(N)
    qmp_head_finish_bwd(y, out, x, d_out, d_xnext, N, F, binary, drop_out, seed_out, dy, dx)
    qmp_tconv1_bwd(N, in_ptr, in_src, ea, h, ldh, P, s4, dy, ds4, dh, lddh, gP, drop_attn, seed_attn)
    if relu_mask and dh is not None:
        d = rows(dh, N, lddh, 32)
        d.mul_((rows(h, N, ldh, 32) > 0).float())


def _pack_params_of(tab):
    from quadtree_mpnnlstm_b200 import fused as FZ
    for table, params in FZ._ptr_tables.values():
        if table is tab:
            return params
    raise KeyError("unknown parameter table")


def qmp_pack_tconv_fwd(tab, G, D, DC, C, out):
    from quadtree_mpnnlstm_b200 import fused as FZ
    ps = _pack_params_of(tab)
    convs = [FZ._ParamView(ps[9 * g:9 * (g + 1)], C) for g in range(G)]
    with torch.no_grad():
        flat(out, G * FZ.conv_total(DC)).view(G, -1).copy_(FZ.pack_fused(convs, DC))


def qmp_pack_tconv_bwd(tab, G, D, DC, C, g, grads):
    from quadtree_mpnnlstm_b200 import fused as FZ
    ps = [p.detach().clone().requires_grad_(True) for p in _pack_params_of(tab)]
    convs = [FZ._ParamView(ps[9 * i:9 * (i + 1)], C) for i in range(G)]
    with torch.enable_grad():
        pack = FZ.pack_fused(convs, DC)
        gs = torch.autograd.grad(pack, ps, flat(g, pack.numel()).view_as(pack), allow_unused=True)
    rows_ = []
    for i in range(G):
        rows_.append(torch.cat([(gs[9 * i + k] if gs[9 * i + k] is not None else torch.zeros_like(ps[9 * i + k])).reshape(-1)
                                for k in range(9)]))
    flat(grads, G * rows_[0].numel()).view(G, -1).copy_(torch.stack(rows_))


def _gat_logits(N, C, mode, ti, sj, EA, XL, as_, ad, we, XR, We, att, slope):
    lr = lambda v: torch.where(v > 0, v, slope * v)
    if mode == 1:
        pre = flat(as_, N)[sj] + flat(ad, N)[ti] + EA @ flat(we, 2)
        return pre, lr(pre)
    m = XL[sj] + XR[ti] + EA @ flat(We, 2 * C).view(C, 2).T
    return m, (lr(m) * flat(att, C)).sum(1)


def qmp_gat_fwd(N, C, mode, in_ptr, in_src, ea, XL, ldl, as_, ad, we, XR, ldr, We, att, slope, out, ldo, alpha):
    E, ti, sj = _edge_lists(N, in_ptr, in_src)
    EA = flat(ea, 2 * E).view(E, 2) if ea is not None else torch.zeros(E, 2)
    xl = rows(XL, N, ldl, C)
    xr = rows(XR, N, ldr, C) if XR is not None else None
    _, lg = _gat_logits(N, C, mode, ti, sj, EA, xl, as_, ad, we, xr, We, att, slope)
    mx = torch.full((N,), -float("inf")).scatter_reduce(0, ti, lg, "amax", include_self=True)
    ex = torch.exp(lg - mx[ti])
    al = ex / (torch.zeros(N).index_add(0, ti, ex) + 1e-16)[ti]
    flat(alpha, E).copy_(al)
    rows(out, N, ldo, C).copy_(torch.zeros(N, C).index_add(0, ti, al[:, None] * xl[sj]))


def qmp_gat_bwd(N, C, mode, in_ptr, in_src, ea, XL, ldl, as_, ad, we, XR, ldr, We, att, slope, alpha, dOut, lddo, dlog, dXL, das, dad, dwe,
                dXR, dWe, datt):
    E, ti, sj = _edge_lists(N, in_ptr, in_src)
    EA = flat(ea, 2 * E).view(E, 2) if ea is not None else torch.zeros(E, 2)
    xl = rows(XL, N, ldl, C)
    xr = rows(XR, N, ldr, C) if XR is not None else None
    g = rows(dOut, N, lddo, C)
    al = flat(alpha, E)
    dal = (g[ti] * xl[sj]).sum(1)
    t = torch.zeros(N).index_add(0, ti, al * dal)
    dl = al * (dal - t[ti])
    gxl = torch.zeros(N, C).index_add(0, sj, al[:, None] * g[ti])
    pre, _ = _gat_logits(N, C, mode, ti, sj, EA, xl, as_, ad, we, xr, We, att, slope)
    d = torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, slope))
    if mode == 1:
        gg = dl * d
        flat(das, N).copy_(torch.zeros(N).index_add(0, sj, gg))
        flat(dad, N).copy_(torch.zeros(N).index_add(0, ti, gg))
        if dwe is not None:
            flat(dwe, 2).add_(gg @ EA)
    else:
        dm = dl[:, None] * flat(att, C) * d
        gxl = gxl + torch.zeros(N, C).index_add(0, sj, dm)
        rows(dXR, N, ldr, C).copy_(torch.zeros(N, C).index_add(0, ti, dm))
        if dWe is not None:
            flat(dWe, 2 * C).view(C, 2).add_(dm.T @ EA)
        if datt is not None:
            flat(datt, C).add_((dl[:, None] * torch.where(pre > 0, pre, slope * pre)).sum(0))
    rows(dXL, N, ldl, C).copy_(gxl)


def qmp_tconv1_fwd(N, in_ptr, in_src, ea, x, ldx, P, s4, out, drop_p, seed):
    E, ti, sj, X, W4, S, EA, e, al = _tconv1_common(N, in_ptr, in_src, ea, x, ldx, P, drop_p)
    flat(s4, 4 * N).view(N, 4).copy_(S)
    flat(out, N).copy_(torch.zeros(N).index_add(0, ti, al * (S[sj, 2] + e)) + S[:, 3])


def qmp_tconv1_bwd(N, in_ptr, in_src, ea, x, ldx, P, s4, g, ds4, dx, lddx, gP, drop_p, seed):
    E, ti, sj, X, W4, S, EA, e, al = _tconv1_common(N, in_ptr, in_src, ea, x, ldx, P, drop_p)
    gv = flat(g, N)
    key, val = S[sj, 1] + e, S[sj, 2] + e
    dal = gv[ti] * val
    tsum = torch.zeros(N).index_add(0, ti, al * dal)
    dsv = al * (dal - tsum[ti])
    dkey, dval = dsv * S[ti, 0], al * gv[ti]
    D = torch.stack([torch.zeros(N).index_add(0, ti, dsv * key), torch.zeros(N).index_add(0, sj, dkey),
                     torch.zeros(N).index_add(0, sj, dval), gv], dim=1)
    flat(ds4, 4 * N).view(N, 4).copy_(D)
    if dx is not None:
        rows(dx, N, lddx, 32).copy_(D @ W4)
    if gP is not None:
        gp = flat(gP, 136)
        gp[:128] += (D.T @ X).reshape(-1)
        gp[128:132] += D.sum(0)
        gp[132:134] += ((dkey + dval)[:, None] * EA).sum(0)


def _bwd_pack_from_image(img, G, DC):
    """Backward pack (W1 | b1 | W1T | W2T | W3T) rebuilt from the forward pack stored in an emulated image."""
    if img is None:
        return None
    w = _pack_from_image(img, G, DC)
    o1 = (DC + 2) * DC
    o2 = o1 + DC + 4
    o3 = o2 + _FC * (DC + 4)
    o4 = o3 + _FC * DC
    W1, b1 = w[:, :o1].view(G, DC + 2, DC), w[:, o1:o2]
    W2, W3 = w[:, o2:o3].view(G, _FC, DC + 4), w[:, o3:o4].view(G, _FC, DC)
    W1T = torch.zeros(G, DC, DC + 4)
    W1T[:, :, :DC + 2] = W1.transpose(1, 2)
    return torch.cat([W1.flatten(1), b1, W1T.flatten(1), W2.transpose(1, 2).flatten(1), W3.transpose(1, 2).flatten(1)], 1).contiguous()


def qmp_fused_bwd_target_tc(N, in_ptr, in_src, ea, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, *rest):
    qmp_fused_bwd_target(N, in_ptr, in_src, ea, xa, lda, DA, GA, _bwd_pack_from_image(wa, GA, _cap(DA, True)) if GA else None, xb,
                         ldb, DB, GB, sharedB, _bwd_pack_from_image(wb, GB, _cap(DB, False)), *rest)


def qmp_fused_bwd_onepass_tc(N, in_ptr, in_src, ea, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, mode, C, dP, lddp, logit, mstat,
                             linv, ds, ZsA, dUsA, ZsB, dUsB, dxa, dxb, drop_p, seed):
    """Target side, then the source side over the out-CSR derived from the in-CSR (the kernel does both in one pass)."""
    for dxp in (dxa, dxb):
        if dxp is not None:
            dxp.zero_()
    qmp_fused_bwd_target_tc(N, in_ptr, in_src, ea, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, mode, C, dP, lddp, logit, mstat,
                            linv, ds, ZsA, dUsA, ZsB, dUsB, dxa, dxb, drop_p, seed)
    E, ti, sj = _edge_lists(N, in_ptr, in_src)
    order = torch.sort(sj, stable=True).indices
    out_ptr = torch.zeros(N + 1, dtype=torch.int32)
    out_ptr[1:] = torch.cumsum(torch.bincount(sj, minlength=N), 0).int()
    qmp_fused_bwd_source_tc(N, out_ptr, ti[order].int(), order.int(), xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, mode, C, dP,
                            lddp, logit, mstat, linv, ds, dxa, dxb, drop_p, seed)


def qmp_fused_bwd_source_tc(N, out_ptr, out_dst, out_kin, xa, lda, DA, GA, wa, xb, ldb, DB, GB, sharedB, wb, *rest):
    qmp_fused_bwd_source(N, out_ptr, out_dst, out_kin, xa, lda, DA, GA, _bwd_pack_from_image(wa, GA, _cap(DA, True)) if GA else None,
                         xb, ldb, DB, GB, sharedB, _bwd_pack_from_image(wb, GB, _cap(DB, False)), *rest)


def qmp_fused_wgrad(N, xa, lda, DA, GA, xb, ldb, DB, GB, sharedB, mode, C, dP, lddp, ZsA, dUsA, ZsB, dUsB, gwa, gwb):
    for (c, seg, g, xp, ld, off, D, DC) in _fused_convs(xa, lda, DA, GA, xb, ldb, DB, GB, sharedB):
        G = GA if seg == 0 else GB
        W = DC + 4
        Zs, dUs, gw = (ZsA, dUsA, gwa) if seg == 0 else (ZsB, dUsB, gwb)
        z = win(flat(Zs), (N, W), (G * W, 1), g * W)
        du = win(flat(dUs), (N, W), (G * W, 1), g * W)
        x = _rows_pad(xp, N, ld, off, D, DC)
        g32 = _dP_of(dP, lddp, N, mode, C, c, GA)
        total = (DC + 2) * DC + DC + 4 + _FC * (DC + 4) + _FC * DC + _FC
        o1 = (DC + 2) * DC
        o2 = o1 + DC + 4
        o3 = o2 + _FC * (DC + 4)
        o4 = o3 + _FC * DC
        row = flat(gw, G * total).view(G, total)[g]
        row[:o1] += (du[:, :DC + 2].T @ x).reshape(-1)
        row[o1:o1 + DC + 2] += du[:, :DC + 2].sum(0)
        row[o2:o3] += (g32.T @ z).reshape(-1)
        row[o3:o4] += (g32.T @ x).reshape(-1)
        row[o4:] += g32.sum(0)


# ------------------------------------------------------------------------------------------ install
class Emulated:
    """Context manager: route ``_lib.call`` to the functions above and let CPU tensors through."""

    def __enter__(self):
        from quadtree_mpnnlstm_b200 import _lib, graph_functions as gf
        self._lib, self._gf = _lib, gf
        self._saved = (_lib.call, gf._device, _lib.lib)
        self._stream_ptr = _lib.stream_ptr
        _lib.stream_ptr = lambda: 0
        table = globals()

        def call(name, *args):
            _lib.CALL_COUNTS[name] = _lib.CALL_COUNTS.get(name, 0) + 1
            table[name](*args)

        class _FakeLib:
            @staticmethod
            def qmp_quadtree_pyramid_cells(n, m, s):
                return 1

            @staticmethod
            def qmp_quadtree_graph_scratch_bytes(n, m, s, t, c):
                return 16

            @staticmethod
            def qmp_fused_cell_image_bytes():
                return 4 * 4 * (_total(4) + _total(32))

            @staticmethod
            def qmp_fused_cell_bwd_image_bytes():
                return 4 * 4 * (_total(4) + _total(32))

            @staticmethod
            def qmp_head_bwd_image_bytes():
                return 4 * _total(36)

        _lib.call = call
        _lib.lib = lambda: _FakeLib
        gf._device = lambda device=None: torch.device("cpu")
        self._is_cuda = torch.Tensor.is_cuda
        torch.Tensor.is_cuda = property(lambda self: True)
        return self

    def __exit__(self, *exc):
        self._lib.call, self._gf._device, self._lib.lib = self._saved
        self._lib.stream_ptr = self._stream_ptr
        torch.Tensor.is_cuda = self._is_cuda
        return False
