// Decoder-cell forward: one GConvLSTM step with 4 TransformerConvs on the 4-wide input X and 4 on the 32-wide hidden
// state H (the launch that runs once per forecast step: 90 of the 100 graph-frames of an ice sample), one persistent
// CTA per SM.  Same contract and outputs as qmp_fused_fwd_tc for that configuration; what differs is the mapping:
//
//  * dense contractions: the four gates side by side -- U = h W1^T (N = 160) and P = [h|x] W3^T (N = 128) are TWO
//    tcgen05 chains per 128-node tile instead of eight, P_g += [z_h|z_x] W2_g^T one chain per gate; 99 MMAs per tile
//    instead of 312 (an M = 128, K = 8 tf32 MMA costs ~96 issue cycles whatever its N: the narrow per-conv chains were
//    a tensor-pipe bound).  A operands live in tensor memory (thread = TMEM lane = node), 3xTF32 split.
//  * edge phase and gate epilogue in OCTET layout: 8 lanes per node, a float4 of the 32-wide row per lane, 4 nodes per
//    warp instruction.  Rows move with naturally coalesced 128-byte accesses (no shared-memory transposes), a thread
//    keeps 4 (not 32) values per row, so 16 warps per SM stay resident, and the four H convs share every gathered
//    neighbour row.  The 16 logits of a node quad (4 convs x 4 edges) are reduced with two 8-value transposing
//    butterflies (7 shuffles each), the softmax runs on (conv, edge) lanes.
//  * the two layouts meet in a shared-memory exchange buffer: U rows TMEM -> smem -> octets, z rows octets -> smem ->
//    TMEM, P rows TMEM -> smem -> octets.
//  * the X convs (4-wide rows) run one thread per (node, conv) in plain FFMA while the first MMA group is in flight;
//    their aggregates ride along as 8 extra K columns of the value contraction.
//
// Reference: GConvLSTM.forward (model/model.py:394-463) around PyG TransformerConv (model/model.py:51), Decoder norms
// and head input (model/seq2seq.py:138-165).
#include <stdlib.h>
#include "fused_fwd.inl"
#include "fused_cell.cuh"

namespace qmp {

constexpr uint32_t TM_P = 0;                  // P accumulators of the four gates, 128 columns
constexpr uint32_t TM_R0 = 128, TM_RW = 96;   // A operand regions R_g = TM_R0 + g * TM_RW (hi 48 | lo 48); R_0 also holds [h|x] (hi 40 | lo 40)
constexpr uint32_t TM_U = 320;                // U of the four H convs, 160 columns, aliases R_2 / R_3 (dead before they are staged)
constexpr size_t CELL_SMEM = CellLayout::BYTES + (13 * FC + 4 * XPLANE) * sizeof(float);

// per-thread state of the pipelined tile loop ------------------------------------------------------------------------
struct CellOwn {                // this thread's share of the [h | x] row of its node (cg 0: h[0..15], 1: h[16..23], 2: h[24..31], 3: x)
    float4 v[4];
};
struct CellIdx {                // edge-phase indices of the two warp passes: first in-edge, in-degree, source / attributes of edge e4
    int k0[2], deg[2], jj[2];
    float2 ev[2];
};

__device__ __forceinline__ void cell_load_own(CellOwn& ow, const FusedFwdArgs& a, int i, bool valid, int cg) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) ow.v[k] = zero;
    if (!valid) return;
    if (cg == 3) {
        ow.v[0] = __ldg(reinterpret_cast<const float4*>(a.xa + (size_t)i * a.lda));
    } else {
        const float4* hp = reinterpret_cast<const float4*>(a.xb + (size_t)i * a.ldb) + (cg == 0 ? 0 : 2 + 2 * cg);
        ow.v[0] = __ldg(hp);
        ow.v[1] = __ldg(hp + 1);
        if (cg == 0) {
            ow.v[2] = __ldg(hp + 2);
            ow.v[3] = __ldg(hp + 3);
        }
    }
}

__device__ __forceinline__ void cell_stage_own(const CellOwn& ow, uint32_t base, int cg) {
    float v[8];
    auto put = [&](const float4& p, const float4& q2, uint32_t k0) {
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q2.x; v[5] = q2.y; v[6] = q2.z; v[7] = q2.w;
        cell_stage8(base + k0, base + 40 + k0, v);
    };
    if (cg == 0) {
        put(ow.v[0], ow.v[1], 0);
        put(ow.v[2], ow.v[3], 8);
    } else if (cg < 3) {
        put(ow.v[0], ow.v[1], 8 + 8 * cg);
    } else {
        put(ow.v[0], make_float4(0.f, 0.f, 0.f, 0.f), 32);
    }
    tc::tmem_st_wait();
}

// X conv cg of node i (row nrow of the tile): plain FFMA, online segment softmax; aggregates -> exchange columns 36..43
struct CellXIdx {               // X conv of one node: in-edge range and the sources of its first four in-edges
    int k0, k1, jn[4];
};
__device__ __forceinline__ void cell_xconv_idx(CellXIdx& xi, const FusedFwdArgs& a, int i, bool valid) {
    xi.k0 = valid ? __ldg(a.ptr + i) : 0;
    xi.k1 = valid ? __ldg(a.ptr + i + 1) : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) xi.jn[e] = (xi.k0 + e < xi.k1) ? __ldg(a.nbr + xi.k0 + e) : -1;
}

__device__ __forceinline__ void cell_xconv(const FusedFwdArgs& a, const uint8_t* smem, float* exch, int i, bool valid, int nrow, int cg,
                                           const CellXIdx& xix) {
    using L = CellLayout;
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) xi = __ldg(reinterpret_cast<const float4*>(a.xa + (size_t)i * a.lda));
    const int k0 = xix.k0, k1 = xix.k1;
    int jn[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) jn[e] = xix.jn[e];
    float2 evn[4];
    float4 xn[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        evn[e] = make_float2(0.f, 0.f);
        xn[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jn[e] >= 0) {
            xn[e] = __ldg(reinterpret_cast<const float4*>(a.xa + (size_t)jn[e] * a.lda));
            if (a.ea) evn[e] = __ldg(reinterpret_cast<const float2*>(a.ea) + k0 + e);
        }
    }
    const float* w1x = reinterpret_cast<const float*>(smem + L::W1X) + cg * 24;
    const float* b1x = reinterpret_cast<const float*>(smem + L::B1X) + cg * 8;
    float u[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const float4 w = ld4(w1x + 4 * r);
        u[r] = fmaf(w.w, xi.w, fmaf(w.z, xi.z, fmaf(w.y, xi.y, fmaf(w.x, xi.x, b1x[r]))));
    }
    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f, ze0 = 0.f, ze1 = 0.f, zs = 0.f, m = -INFINITY, l = 0.f;
    auto edge = [&](int kk, const float4& xj, const float2& ev) {
        const float s = fmaf(u[3], xj.w, fmaf(u[2], xj.z, fmaf(u[1], xj.y, fmaf(u[0], xj.x, fmaf(u[4], ev.x, u[5] * ev.y)))));
        a.logit[(size_t)kk * 8 + cg] = s;
        const float mn = fmaxf(m, s);
        const float sc = fast_exp(m - mn), pe = fast_exp(s - mn);
        const float pk = pe * fdropout_scale(QMP_SEED_SM, (long long)kk * 8 + cg, a.drop_p);
        l = fmaf(l, sc, pe);
        zs = fmaf(zs, sc, pk);
        ze0 = fmaf(ze0, sc, pk * ev.x);
        ze1 = fmaf(ze1, sc, pk * ev.y);
        z0 = fmaf(z0, sc, pk * xj.x);
        z1 = fmaf(z1, sc, pk * xj.y);
        z2 = fmaf(z2, sc, pk * xj.z);
        z3 = fmaf(z3, sc, pk * xj.w);
        m = mn;
    };
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (jn[e] >= 0) edge(k0 + e, xn[e], evn[e]);
    for (int kk = k0 + 4; kk < k1; ++kk) {                     // larger in-degrees (quadtree meshes)
        const int j = __ldg(a.nbr + kk);
        const float4 xj = __ldg(reinterpret_cast<const float4*>(a.xa + (size_t)j * a.lda));
        float2 ev = make_float2(0.f, 0.f);
        if (a.ea) ev = __ldg(reinterpret_cast<const float2*>(a.ea) + kk);
        edge(kk, xj, ev);
    }
    const float li = (l > 0.f) ? fast_rcp(l) : 0.f;
    if (valid) {
        a.mstat[(size_t)i * 8 + cg] = m;
        a.linv[(size_t)i * 8 + cg] = li;
    }
    float* zr = exch + cg * XPLANE + nrow * XS + 36;
    st4(zr, z0 * li, z1 * li, z2 * li, z3 * li);
    st4(zr + 4, ze0 * li, ze1 * li, zs * li, 0.f);
}

__device__ __forceinline__ void cell_load_idx(CellIdx& ix, const FusedFwdArgs& a, int tile0, int tcount, int warp, int o8, int e4) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int ln = 4 * (warp + 16 * p) + o8;
        const bool valid = ln < tcount;
        const int i = tile0 + ln;
        ix.k0[p] = valid ? __ldg(a.ptr + i) : 0;
        ix.deg[p] = valid ? __ldg(a.ptr + i + 1) - ix.k0[p] : 0;
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const bool on = e4 < ix.deg[p];
        ix.jj[p] = on ? __ldg(a.nbr + ix.k0[p] + e4) : -1;
        ix.ev[p] = make_float2(0.f, 0.f);
        if (on && a.ea) ix.ev[p] = __ldg(reinterpret_cast<const float2*>(a.ea) + ix.k0[p] + e4);
    }
}

__device__ __forceinline__ void cell_issue_g1(uint32_t tmem, const uint8_t* smem, uint64_t* bar) {
    using L = CellLayout;
    tc::fence_after_sync();
    tc_mma3_at(0, tmem + TM_U, tmem + TM_R0, tmem + TM_R0 + 40, tc::smem_u32(smem + L::W1H), tc::smem_u32(smem + L::W1L), L::NU,
               L::KU, false);
    tc_mma3_at(0, tmem + TM_P, tmem + TM_R0, tmem + TM_R0 + 40, tc::smem_u32(smem + L::W3H), tc::smem_u32(smem + L::W3L), L::NS,
               L::KS, false);
    tc::commit(bar);
}

// Software pipeline of a CTA over its tiles (tensor pipe and SIMT phases of consecutive tiles overlap):
//   ... | U dump(t) | edge phase(t) | stage z(t) | G2(t) issued || X convs(t+1), row / index prefetch(t+1) || P dump(t) |
//       stage [h|x](t+1) | G1(t+1) issued || gate epilogue(t) || U dump(t+1) | ...
// tiles of a CTA: it owns the Q consecutive nodes from blockIdx.x * Q and walks them in R tiles of T0 (+ / - stagger) nodes
struct CellTiling {
    int Q, R, T0, stagger;
};
// Tile r of this CTA: first node and node count.  With R == 3 the tile sizes are T0 + d, T0, T0 - d rotated by blockIdx.x % 3,
// so neighbouring SMs drift out of phase and their store-heavy epilogues do not all hit L2 / HBM in the same microsecond.
__device__ __forceinline__ void cell_tile(const CellTiling& tl, int N, int r, int& start, int& count) {
    const int beg = (int)blockIdx.x * tl.Q;
    int end = beg + tl.Q;
    if (end > N) end = N;
    int s0 = beg;
    int size = tl.T0;
    if (tl.R == 3 && tl.stagger > 0) {
        const int rot = (int)(blockIdx.x % 3);
        const int d0 = (rot == 0) ? tl.stagger : (rot == 1 ? 0 : -tl.stagger);           // sizes of tiles 0, 1, 2:
        const int d1 = (rot == 0) ? 0 : (rot == 1 ? -tl.stagger : tl.stagger);            // rot 0: +d 0 -d, rot 1: 0 -d +d,
        const int d2 = -d0 - d1;                                                          // rot 2: -d +d 0
        if (r >= 1) s0 += tl.T0 + d0;
        if (r >= 2) s0 += tl.T0 + d1;
        size += (r == 0) ? d0 : (r == 1 ? d1 : d2);
    } else {
        s0 += r * tl.T0;
    }
    start = s0;
    count = end - s0;
    if (count > size) count = size;
    if (count < 0) count = 0;
}

__global__ void __launch_bounds__(CELL_THREADS, 1) fused_cell_fwd_kernel(const __grid_constant__ FusedFwdArgs a,
                                                                         const uint8_t* __restrict__ img, const CellTiling tl,
                                                                         float* __restrict__ usave) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[2];                 // 0: MMA groups (every commit is waited once by every thread), 1: image landed
    __shared__ uint32_t tmem_slot;
    using L = CellLayout;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    qmp_seed_init(a.seed, a.salt);
    float* prm = reinterpret_cast<float*>(smem + L::BYTES);
    float* exch = prm + 13 * FC;
    CELL_CTA(0);
    if (t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_mbar_init();
        // the weight image (130 KB, ~4 us from L2) starts moving before the tensor-memory allocation and the first barrier
        tc::mbar_expect_tx(&bars[1], (uint32_t)L::BYTES);
        for (int off = 0; off < L::BYTES; off += 16384)
            tc::bulk_g2s(smem + off, img + off, (uint32_t)(L::BYTES - off < 16384 ? L::BYTES - off : 16384), &bars[1]);
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    for (int idx = t; idx < 13 * FC; idx += CELL_THREADS) prm[idx] = a.params[idx];
    pdl_wait();            // everything above reads step constants only (weight image, gate parameters)
    pdl_launch();
    tc::fence_before_sync();
    cell_sync();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int q = warp & 3, cg = warp >> 2;                    // TMEM lane quarter; column group = conv = gate of the TMEM-side phases
    const int nrow = q * 32 + lane;                            // node row this thread owns in the TMEM-side phases
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int o8 = lane >> 3, l8 = lane & 7, obase = lane & ~7;  // octet of the warp, lane in the octet
    const int cc = l8 >> 2, e4 = l8 & 3;                       // (conv of the pair, edge of the quad) role in the softmax stage
    const float* b1h = reinterpret_cast<const float*>(smem + L::B1H);
    const float* b3s = reinterpret_cast<const float*>(smem + L::B3S);
    uint32_t par = 0;
    int tile = 0;                                              // tile (round) index of this CTA, -1: none left
    {
        int s0, c0;
        cell_tile(tl, a.N, 0, s0, c0);
        if (c0 <= 0) tile = -1;
    }
    if (warp == CELL_WORKERS / 32) {
        // ---- the MMA warp: one lane issues every tcgen05.mma of the CTA, in step with the workers' barriers (an issuing
        // thread is held for the ~90 cycles each MMA occupies the tensor pipe; the workers must not be)
        tc::mbar_wait(&bars[1], 0);
        cell_sync();                                           // first [h|x] rows staged
        if (lane == 0 && tile >= 0) cell_issue_g1(tmem, smem, &bars[0]);
        __syncwarp();
        while (tile >= 0) {
            int next = tile + 1, ns0, nc0;
            cell_tile(tl, a.N, next, ns0, nc0);
            if (next >= tl.R || nc0 <= 0) next = -1;
            cell_sync();                                       // U dumped
            cell_sync();                                       // edge phase done
            cell_sync();                                       // z rows staged
            if (lane == 0) {
                tc::fence_after_sync();
#pragma unroll 1
                for (int g = 0; g < 4; ++g)
                    tc_mma3_at(0, tmem + TM_P + 32 * g, tmem + TM_R0 + TM_RW * g, tmem + TM_R0 + TM_RW * g + 48,
                               tc::smem_u32(smem + L::W2H + g * L::W2G), tc::smem_u32(smem + L::W2H + g * L::W2G + L::NZ * L::KZ * 4),
                               L::NZ, L::KZ, true);
                tc::commit(&bars[0]);
            }
            __syncwarp();
            CELL_MARK(12);
            cell_sync();                                       // P dumped, next [h|x] rows staged
            if (lane == 0 && next >= 0) cell_issue_g1(tmem, smem, &bars[0]);
            __syncwarp();
            CELL_MARK(4);
            cell_sync();                                       // epilogue done
            tile = next;
        }
    } else {
    CellIdx ix;
    {   // prologue: the first tile's X convs, [h|x] rows and first contraction
        CellOwn ow;
        int tile0 = 0, tcount = 0;
        if (tile >= 0) cell_tile(tl, a.N, 0, tile0, tcount);
        cell_load_own(ow, a, tile0 + nrow, nrow < tcount, cg);
        cell_load_idx(ix, a, tile0, tcount, warp, o8, e4);
        CellXIdx xix;
        cell_xconv_idx(xix, a, tile0 + nrow, nrow < tcount);
        tc::mbar_wait(&bars[1], 0);                            // weights in shared memory
        cell_xconv(a, smem, exch, tile0 + nrow, nrow < tcount, nrow, cg, xix);
        cell_stage_own(ow, lane_addr + TM_R0, cg);
        tc::fence_before_sync();
        cell_sync();
        CELL_CTA(1);
    }

    while (tile >= 0) {
        int tile0, tcount, next = tile + 1, next0, ncount;
        cell_tile(tl, a.N, tile, tile0, tcount);
        cell_tile(tl, a.N, next, next0, ncount);
        if (next >= tl.R || ncount <= 0) {
            next = -1;
            ncount = 0;
        }
        CELL_MARK(1);
        CellXIdx xix;                                          // next tile's X convs: indices now, rows and math under G2
        cell_xconv_idx(xix, a, next0 + nrow, nrow < ncount);
        tc::mbar_wait(&bars[0], par);                          // G1(tile): U and the skip part of P
        par ^= 1;
        tc::fence_after_sync();
        CELL_MARK(6);

        // ---- U block of conv cg, row nrow: tensor memory -> exchange plane cg (+ logit bias)
        {
            uint32_t r[5][8];
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) tc::tmem_ld8_nowait(lane_addr + TM_U + (uint32_t)(L::UB * cg + 8 * c8), r[c8]);
            tc::tmem_ld_wait();
            float* row = exch + cg * XPLANE + nrow * XS;
            const float* b = b1h + cg * L::UB;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                const float4 b0 = ld4(b + 8 * c8), b1 = ld4(b + 8 * c8 + 4);
                st4(row + 8 * c8, __uint_as_float(r[c8][0]) + b0.x, __uint_as_float(r[c8][1]) + b0.y, __uint_as_float(r[c8][2]) + b0.z,
                    __uint_as_float(r[c8][3]) + b0.w);
                st4(row + 8 * c8 + 4, __uint_as_float(r[c8][4]) + b1.x, __uint_as_float(r[c8][5]) + b1.y, __uint_as_float(r[c8][6]) + b1.z,
                    __uint_as_float(r[c8][7]) + b1.w);
            }
            const float4 b0 = ld4(b + 32);
            st4(row + 32, __uint_as_float(r[4][0]) + b0.x, __uint_as_float(r[4][1]) + b0.y, 0.f, 0.f);
        }
        // first edge quad of both passes: neighbour rows in flight across the barrier
        float4 hq[2][4];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int jx = __shfl_sync(0xffffffffu, ix.jj[p], obase + x);
                hq[p][x] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (jx >= 0) hq[p][x] = __ldg(reinterpret_cast<const float4*>(a.xb + (size_t)jx * a.ldb) + l8);
            }
        cell_sync();
        CELL_MARK(8);

        // ---- edge phase of the four H convs, octet layout, 4 nodes per warp pass
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int ln = 4 * (warp + 16 * p) + o8;
            const bool valid = ln < tcount;
            const int i = tile0 + ln;
            const int k0 = ix.k0[p], deg = ix.deg[p];
            float* xrow = exch + ln * XS;
            float4 u[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) u[c] = ld4(xrow + c * XPLANE + 4 * l8);
            if (usave && valid) {               // logit projections u_c = W1_c h + b1_c of the four H convs: the backward kernel's
#pragma unroll                                  // source-side term ds_e u_i needs them (fused_cell_bwd.cu)
                for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(usave + (size_t)i * 128 + 32 * c + 4 * l8) = u[c];
            }
            float2 w01[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) w01[r] = *reinterpret_cast<const float2*>(xrow + (2 * r + cc) * XPLANE + 32);
            float4 z[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, ze0[2] = {0.f, 0.f}, ze1[2] = {0.f, 0.f}, zs[2] = {0.f, 0.f};
            float4 hr[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) hr[x] = hq[p][x];
            int jj = ix.jj[p];
            float2 ev = ix.ev[p];
            for (int qd = 0; qd == 0 || __any_sync(0xffffffffu, 4 * qd < deg); ++qd) {
                const bool on = 4 * qd + e4 < deg;
                const int kk = k0 + 4 * qd + e4;
                if (qd > 0) {                                   // larger in-degrees (quadtree meshes): next quad of edges
                    jj = on ? __ldg(a.nbr + kk) : -1;
                    ev = make_float2(0.f, 0.f);
                    if (on && a.ea) ev = __ldg(reinterpret_cast<const float2*>(a.ea) + kk);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int jx = __shfl_sync(0xffffffffu, jj, obase + x);
                        hr[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (jx >= 0) hr[x] = __ldg(reinterpret_cast<const float4*>(a.xb + (size_t)jx * a.ldb) + l8);
                    }
                }
                float sc[2], pk[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float v[8];
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
                        for (int x = 0; x < 4; ++x) v[4 * c2 + x] = dot4(u[2 * r + c2], hr[x]);
                    const float tot = octet_reduce8(v, l8);                       // lane (cc, e4): conv 2r + cc, edge e4
                    const float s = on ? fmaf(w01[r].x, ev.x, fmaf(w01[r].y, ev.y, tot)) : -INFINITY;
                    const int c = 4 + 2 * r + cc;
                    if (on) a.logit[(size_t)kk * 8 + c] = s;
                    float gm = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, 1));
                    gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, 2));
                    const float mn = fmaxf(m[r], gm);
                    sc[r] = (mn == -INFINITY) ? 1.f : fast_exp(m[r] - mn);
                    const float pe = on ? fast_exp(s - mn) : 0.f;
                    pk[r] = pe * fdropout_scale(QMP_SEED_SM, (long long)kk * 8 + c, a.drop_p);
                    l[r] = fmaf(l[r], sc[r], quad_sum(pe));
                    zs[r] = fmaf(zs[r], sc[r], quad_sum(pk[r]));
                    ze0[r] = fmaf(ze0[r], sc[r], quad_sum(pk[r] * ev.x));
                    ze1[r] = fmaf(ze1[r], sc[r], quad_sum(pk[r] * ev.y));
                    m[r] = mn;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int src = obase + 4 * (c & 1);
                    if (qd > 0) {
                        const float f = __shfl_sync(0xffffffffu, sc[c >> 1], src);
                        z[c].x *= f; z[c].y *= f; z[c].z *= f; z[c].w *= f;
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float pb = __shfl_sync(0xffffffffu, pk[c >> 1], src + x);
                        z[c].x = fmaf(pb, hr[x].x, z[c].x);
                        z[c].y = fmaf(pb, hr[x].y, z[c].y);
                        z[c].z = fmaf(pb, hr[x].z, z[c].z);
                        z[c].w = fmaf(pb, hr[x].w, z[c].w);
                    }
                }
            }
            float li[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                li[r] = (l[r] > 0.f) ? fast_rcp(l[r]) : 0.f;
                if (valid && e4 == 0) {
                    a.mstat[(size_t)i * 8 + 4 + 2 * r + cc] = m[r];
                    a.linv[(size_t)i * 8 + 4 + 2 * r + cc] = li[r];
                }
            }
            __syncwarp();                                      // every lane of the octet has read u / w of this row
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float f = __shfl_sync(0xffffffffu, li[c >> 1], obase + 4 * (c & 1));
                st4(xrow + c * XPLANE + 4 * l8, z[c].x * f, z[c].y * f, z[c].z * f, z[c].w * f);
            }
            if (e4 == 0) {
#pragma unroll
                for (int r = 0; r < 2; ++r) st4(xrow + (2 * r + cc) * XPLANE + 32, ze0[r] * li[r], ze1[r] * li[r], zs[r] * li[r], 0.f);
            }
        }
        CELL_MARK(9);
        cell_sync();
        CELL_MARK(10);

        // ---- [z_h | ze zs | z_x] of gate cg, row nrow -> tensor memory region R_cg; then the four value contractions (G2)
        {
            const float* row = exch + cg * XPLANE + nrow * XS;
            const uint32_t base = lane_addr + TM_R0 + TM_RW * (uint32_t)cg;
            float v[8];
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                ld8(v, row + 8 * c8);
                cell_stage8(base + 8 * c8, base + 48 + 8 * c8, v);
            }
            const float4 zt = ld4(row + 32);
            v[0] = zt.x; v[1] = zt.y; v[2] = zt.z; v[3] = v[4] = v[5] = v[6] = v[7] = 0.f;
            cell_stage8(base + 32, base + 80, v);
            ld8(v, row + 36);
            cell_stage8(base + 40, base + 88, v);
            tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        cell_sync();
        CELL_MARK(11);

        // ---- while G2 runs: the next tile's rows, indices and X convs; this tile's previous cell state
        CellOwn ow;
        cell_load_own(ow, a, next0 + nrow, nrow < ncount, cg);
        cell_load_idx(ix, a, next0, ncount, warp, o8, e4);
        float4 cp4[2];
        float cct[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int ln = 4 * (warp + 16 * p) + o8;
            cp4[p] = make_float4(0.f, 0.f, 0.f, 0.f);
            cct[p] = 0.f;
            if (ln < tcount && a.Cprev) cp4[p] = __ldg(reinterpret_cast<const float4*>(a.Cprev + (size_t)(tile0 + ln) * FC) + l8);
            if (ln < tcount && a.concat && l8 == 0) cct[p] = __ldg(a.concat + tile0 + ln);
        }
        if (next >= 0) cell_xconv(a, smem, exch, next0 + nrow, nrow < ncount, nrow, cg, xix);
        CELL_MARK(5);
        tc::mbar_wait(&bars[0], par);                          // G2(tile)
        par ^= 1;
        tc::fence_after_sync();
        CELL_MARK(13);

        // ---- P block of gate cg, row nrow: tensor memory -> exchange plane cg (+ skip biases); next tile's [h|x] -> tensor memory
        {
            uint32_t r[4][8];
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) tc::tmem_ld8_nowait(lane_addr + TM_P + (uint32_t)(32 * cg + 8 * c8), r[c8]);
            tc::tmem_ld_wait();
            float* row = exch + cg * XPLANE + nrow * XS;
            const float* b = b3s + cg * FC;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                const float4 b0 = ld4(b + 8 * c8), b1 = ld4(b + 8 * c8 + 4);
                st4(row + 8 * c8, __uint_as_float(r[c8][0]) + b0.x, __uint_as_float(r[c8][1]) + b0.y, __uint_as_float(r[c8][2]) + b0.z,
                    __uint_as_float(r[c8][3]) + b0.w);
                st4(row + 8 * c8 + 4, __uint_as_float(r[c8][4]) + b1.x, __uint_as_float(r[c8][5]) + b1.y, __uint_as_float(r[c8][6]) + b1.z,
                    __uint_as_float(r[c8][7]) + b1.w);
            }
        }
        if (next >= 0) cell_stage_own(ow, lane_addr + TM_R0, cg);
        tc::fence_before_sync();
        cell_sync();
        CELL_MARK(14);                                         // the MMA warp issues G1(next): it runs under this tile's epilogue

        // ---- gate epilogue in octet layout (model/model.py:430-463, model/seq2seq.py:138-165), both passes interleaved
        // params rows (lstm.cu): 0 wci 1 wcf 2 wco 3 bi 4 bf 5 bc 6 bo 7 gh 8 bh 9 gc 10 bc 11 go 12 bo
        {
            const float* pr = prm + 4 * l8;
            const float4 wci = ld4(pr), wcf = ld4(pr + FC), wco = ld4(pr + 2 * FC), bi = ld4(pr + 3 * FC), bf = ld4(pr + 4 * FC),
                         bc = ld4(pr + 5 * FC), bo = ld4(pr + 6 * FC);
            const float wci_[4] = {wci.x, wci.y, wci.z, wci.w}, wcf_[4] = {wcf.x, wcf.y, wcf.z, wcf.w}, wco_[4] = {wco.x, wco.y, wco.z, wco.w};
            const float bi_[4] = {bi.x, bi.y, bi.z, bi.w}, bf_[4] = {bf.x, bf.y, bf.z, bf.w}, bc_[4] = {bc.x, bc.y, bc.z, bc.w},
                        bo_[4] = {bo.x, bo.y, bo.z, bo.w};
            float I[2][4], F[2][4], Tg[2][4], Cn[2][4], O[2][4], Hh[2][4];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const float* xrow = exch + (4 * (warp + 16 * p) + o8) * XS + 4 * l8;
                const float4 Pi = ld4(xrow), Pf = ld4(xrow + XPLANE), Pc = ld4(xrow + 2 * XPLANE), Po = ld4(xrow + 3 * XPLANE);
                const float cp[4] = {cp4[p].x, cp4[p].y, cp4[p].z, cp4[p].w};
                const float pi[4] = {Pi.x, Pi.y, Pi.z, Pi.w}, pf[4] = {Pf.x, Pf.y, Pf.z, Pf.w}, pc[4] = {Pc.x, Pc.y, Pc.z, Pc.w},
                            po[4] = {Po.x, Po.y, Po.z, Po.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    I[p][k] = sigm(pi[k] + wci_[k] * cp[k] + bi_[k]);
                    F[p][k] = sigm(pf[k] + wcf_[k] * cp[k] + bf_[k]);
                    Tg[p][k] = ftanh(pc[k] + bc_[k]);
                    Cn[p][k] = fmaf(F[p][k], cp[k], I[p][k] * Tg[p][k]);
                    O[p][k] = sigm(po[k] + wco_[k] * Cn[p][k] + bo_[k]);
                    Hh[p][k] = O[p][k] * ftanh(Cn[p][k]);
                }
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int ln = 4 * (warp + 16 * p) + o8;
                const size_t i = (size_t)(tile0 + ln);
#ifdef QMP_EXP_NOSTORE      // timing experiment only: how much of the epilogue is store back-pressure
                if (ln < tcount && a.eps < 0.f) {
#else
                if (ln < tcount) {
#endif
                    float* gs = a.gates + i * (4 * FC) + 4 * l8;
                    st4(gs, I[p][0], I[p][1], I[p][2], I[p][3]);
                    st4(gs + FC, F[p][0], F[p][1], F[p][2], F[p][3]);
                    st4(gs + 2 * FC, Tg[p][0], Tg[p][1], Tg[p][2], Tg[p][3]);
                    st4(gs + 3 * FC, O[p][0], O[p][1], O[p][2], O[p][3]);
                    st4(a.Craw + i * FC + 4 * l8, Cn[p][0], Cn[p][1], Cn[p][2], Cn[p][3]);
                    if (a.Oout) st4(a.Oout + i * FC + 4 * l8, O[p][0], O[p][1], O[p][2], O[p][3]);
                }
            }
            if (a.norm_h) {
                const float4 g = ld4(pr + 7 * FC), b = ld4(pr + 8 * FC);
                octet_layer_norm(Hh[0], a.eps, g, b);
                octet_layer_norm(Hh[1], a.eps, g, b);
            }
            if (a.norm_c) {
                const float4 g = ld4(pr + 9 * FC), b = ld4(pr + 10 * FC);
                octet_layer_norm(Cn[0], a.eps, g, b);
                octet_layer_norm(Cn[1], a.eps, g, b);
            }
            if (a.head_in && a.norm_o) {
                const float4 g = ld4(pr + 11 * FC), b = ld4(pr + 12 * FC);
                octet_layer_norm(O[0], a.eps, g, b);
                octet_layer_norm(O[1], a.eps, g, b);
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int ln = 4 * (warp + 16 * p) + o8;
                const size_t i = (size_t)(tile0 + ln);
#ifdef QMP_EXP_NOSTORE
                if (ln < tcount && a.eps < 0.f) {
#else
                if (ln < tcount) {
#endif
                    st4(a.Hout + i * FC + 4 * l8, Hh[p][0], Hh[p][1], Hh[p][2], Hh[p][3]);
                    st4(a.Cout + i * FC + 4 * l8, Cn[p][0], Cn[p][1], Cn[p][2], Cn[p][3]);
                    if (a.head_in) {
                        float* hr = a.head_in + i * a.ldh;
                        st4(hr + 4 * l8, fmaxf(O[p][0], 0.f), fmaxf(O[p][1], 0.f), fmaxf(O[p][2], 0.f), fmaxf(O[p][3], 0.f));
                        if (l8 == 0) {                    // column 32: concat layer (left alone when absent), then zero pads
                            if (a.ldh == FC + 4 && a.concat) {
                                st4(hr + FC, cct[p], 0.f, 0.f, 0.f);
                            } else {
                                if (a.concat) hr[FC] = cct[p];
                                for (int k = FC + 1; k < a.ldh; ++k) hr[k] = 0.f;
                            }
                        }
                    }
                }
            }
        }
        CELL_MARK(15);
        cell_sync();                                       // exchange planes free for the next tile's U
        CELL_MARK(16);
        tile = next;
    }
    }
    CELL_CTA(2);
    tc::fence_before_sync();
    cell_sync();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
    CELL_CTA(3);
}

// ---- weight image ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cell_put(uint8_t* img, int off_hi, int off_lo, int n, int k, int K, float v) {
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    *reinterpret_cast<float*>(img + off_hi + img_off(n, k, K)) = hi;
    *reinterpret_cast<float*>(img + off_lo + img_off(n, k, K)) = lo;
}

__global__ void __launch_bounds__(256) fused_pack_cell_kernel(const float* __restrict__ packA, const float* __restrict__ packB,
                                                              uint8_t* __restrict__ img) {
    using L = CellLayout;
    using SA = ConvSizes<4>;
    using SB = ConvSizes<32>;
    const int tid = blockIdx.x * 256 + threadIdx.x, nth = gridDim.x * 256;
    auto A = [&](int g) { return packA + (size_t)g * SA::TOTAL; };
    auto B = [&](int g) { return packB + (size_t)g * SB::TOTAL; };
    for (int idx = tid; idx < L::NU * L::KU; idx += nth) {             // W1: block g rows u (32) | w (2) | 0
        const int n = idx / L::KU, k = idx % L::KU, g = n / L::UB, r = n % L::UB;
        cell_put(img, L::W1H, L::W1L, n, k, L::KU, r < 34 ? B(g)[r * 32 + k] : 0.f);
    }
    for (int idx = tid; idx < L::NS * L::KS; idx += nth) {             // W3: row g*32+o, columns h | x | 0
        const int n = idx / L::KS, k = idx % L::KS, g = n / FC, o = n % FC;
        float v = 0.f;
        if (k < 32) v = (B(g) + SB::W1 + SB::B1 + SB::W2)[o * 32 + k];
        else if (k < 36) v = (A(g) + SA::W1 + SA::B1 + SA::W2)[o * 4 + (k - 32)];
        cell_put(img, L::W3H, L::W3L, n, k, L::KS, v);
    }
    for (int idx = tid; idx < 4 * L::NZ * L::KZ; idx += nth) {         // W2_g: columns z_h | ze zs 0 | 0 | z_x | ze zs 0
        const int g = idx / (L::NZ * L::KZ), rem = idx % (L::NZ * L::KZ), o = rem / L::KZ, k = rem % L::KZ;
        float v = 0.f;
        if (k < 36) v = (B(g) + SB::W1 + SB::B1)[o * 36 + k];
        else if (k >= 40) v = (A(g) + SA::W1 + SA::B1)[o * 8 + (k - 40)];
        cell_put(img, L::W2H + g * L::W2G, L::W2H + g * L::W2G + L::NZ * L::KZ * 4, o, k, L::KZ, v);
    }
    float* b1h = reinterpret_cast<float*>(img + L::B1H);
    for (int idx = tid; idx < 4 * L::UB; idx += nth) {
        const int g = idx / L::UB, r = idx % L::UB;
        b1h[idx] = r < 36 ? (B(g) + SB::W1)[r] : 0.f;
    }
    float* w1x = reinterpret_cast<float*>(img + L::W1X);
    for (int idx = tid; idx < 4 * 24; idx += nth) w1x[idx] = A(idx / 24)[idx % 24];
    float* b1x = reinterpret_cast<float*>(img + L::B1X);
    for (int idx = tid; idx < 4 * 8; idx += nth) b1x[idx] = (A(idx / 8) + SA::W1)[idx % 8];
    float* b3 = reinterpret_cast<float*>(img + L::B3S);
    for (int idx = tid; idx < 4 * FC; idx += nth) {
        const int g = idx / FC, o = idx % FC;
        b3[idx] = (A(g) + SA::W1 + SA::B1 + SA::W2 + SA::W3)[o] + (B(g) + SB::W1 + SB::B1 + SB::W2 + SB::W3)[o];
    }
}

}  // namespace qmp
using namespace qmp;

#ifdef QMP_CELL_TRACE
extern "C" __attribute__((visibility("default"))) int qmpx_cell_trace_dump(float* host_out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out, g_cell_trace, sizeof(float) * 2 * 2048);
    if (reset) {
        static float zeros[2 * 2048];
        cudaMemcpyToSymbol(g_cell_trace, zeros, sizeof(zeros));
    }
    return 0;
}
extern "C" __attribute__((visibility("default"))) int qmpx_cell_cta_dump(unsigned long long* host_out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out, g_cell_cta, sizeof(unsigned long long) * 256 * 4);
    return 0;
}
#endif

static int g_cell_stagger = -1;        // nodes; QMP_CELL_STAGGER overrides the default (timing experiments)

// Bytes of the decoder-cell weight image.
QMP_API long long qmp_fused_cell_image_bytes(void) { return CellLayout::BYTES; }

// packA [4, TOTAL(4)], packB [4, TOTAL(32)] (fused.cuh layout: the X convs and the H convs of gates i, f, c, o) ->
// out [qmp_fused_cell_image_bytes()]
QMP_API int qmp_fused_pack_cell(const float* packA, const float* packB, void* out, void* stream) {
    fused_pack_cell_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(packA, packB, (uint8_t*)out);
    QMP_LAUNCH_CHECK("fused_pack_cell_kernel");
    qmp::after_producer();
    return 0;
}

// One decoder-cell step (qmp_fused_fwd_tc with DA = 4, GA = 4, DB = 32, GB = 4, shared H input, gate mode, C = 32) from
// the cell image built by qmp_fused_pack_cell.  Rows of xa (4 floats), xb (32 floats), head_in must be 16-byte aligned.
// usave [N, 128] (may be NULL): receives the logit projections u of the four H convs, which qmp_fused_cell_bwd reads.
QMP_API int qmp_fused_cell_fwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* xa, int lda,
                               const float* xb, int ldb, const void* image, const float* Cprev, const float* params,
                               int norm_h, int norm_c, int norm_o, float eps, float* gates, float* Craw, float* Oout,
                               float* Hout, float* Cout, float* head_in, int ldh, const float* concat, float* logit,
                               float* mstat, float* linv, float* usave, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldb >= 32 && al16(xa) && al16(xb) && al16(image) && al16(params),
                "qmp_fused_cell_fwd: input rows must be 16-byte aligned");
    QMP_REQUIRE(al16(gates) && al16(Craw) && al16(Hout) && al16(Cout) && (!Oout || al16(Oout)) && (!Cprev || al16(Cprev)) &&
                    (!head_in || (al16(head_in) && ldh % 4 == 0 && ldh > FC)),
                "qmp_fused_cell_fwd: output rows must be 16-byte aligned");
    QMP_REQUIRE(!usave || al16(usave), "qmp_fused_cell_fwd: usave must be 16-byte aligned");
    QMP_REQUIRE(!ea || (reinterpret_cast<uintptr_t>(ea) & 7) == 0, "qmp_fused_cell_fwd: edge attributes must be 8-byte aligned");
    FusedFwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.xa = xa; a.lda = lda; a.DA = 4; a.GA = 4; a.xb = xb; a.ldb = ldb;
    a.DB = 32; a.GB = 4; a.sharedB = 1; a.NC = 8; a.mode = 1; a.C = FC; a.Cprev = Cprev; a.params = params; a.norm_h = norm_h;
    a.norm_c = norm_c; a.norm_o = norm_o; a.eps = eps; a.gates = gates; a.Craw = Craw; a.Oout = Oout; a.Hout = Hout;
    a.Cout = Cout; a.head_in = head_in; a.ldh = ldh; a.concat = concat; a.logit = logit; a.mstat = mstat; a.linv = linv;
    a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QMP_CUDA(cudaFuncSetAttribute(fused_cell_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CELL_SMEM));
    }
    // every CTA owns Q consecutive nodes and walks them in R equally full tiles of <= 128 nodes (cell_tile)
    if (g_cell_stagger < 0) {
        const char* e = getenv("QMP_CELL_STAGGER");
        g_cell_stagger = e ? atoi(e) : 0;      // measured: 0, 12 and 20 nodes give the same 58 us (the epilogue is not a chip-wide store burst)
    }
    const int G = cdiv(N, 128) < n_sm ? cdiv(N, 128) : n_sm;
    CellTiling tl;
    tl.Q = (cdiv(N, G) + 3) & ~3;
    tl.R = cdiv(tl.Q, 128);
    tl.T0 = (cdiv(tl.Q, tl.R) + 3) & ~3;
    tl.stagger = 128 - tl.T0 < g_cell_stagger ? (128 - tl.T0) & ~3 : g_cell_stagger;
    QMP_CUDA(launch_pdl(fused_cell_fwd_kernel, dim3(cdiv(N, tl.Q)), dim3(CELL_THREADS), CELL_SMEM, (cudaStream_t)stream, a,
                        reinterpret_cast<const uint8_t*>(image), tl, usave));
    QMP_LAUNCH_CHECK("fused_cell_fwd_kernel");
    return 0;
}
