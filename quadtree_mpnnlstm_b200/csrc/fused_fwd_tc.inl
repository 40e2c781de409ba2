// tcgen05 version of the fused forward of a conv layer group (same contract, arguments and outputs as
// fused_fwd.inl, which stays as the fp32-FFMA cross-check): per 128-node tile and per conv
//
//   x_i  --tcgen05.st-->  TMEM A   --mma-->  U = [u | w] = W1 x + b1       (logit projection)
//                                  --mma-->  P(gate) (+)= W3 x              (skip projection, accumulated per gate)
//   U    --tcgen05.ld-->  registers;  SIMT edge phase over the in-CSR: s_ij = u_i . x_j + w_i . e_ij, online segment
//                          softmax, z_i = sum_j alpha_ij x_j (fp32, rows gathered straight from global / L2)
//   z_i  --tcgen05.st-->  TMEM A   --mma-->  P(gate) += W2 [z | ze | zs]
//   P    --tcgen05.ld-->  registers;  gate epilogue (sigmoid / tanh / peepholes / LayerNorms / head input) or plain
//                          output row.
// Weights arrive as pre-split images (fused_pack_tc.cu) through a double-buffered shared-memory slot.
// Reference: GConvLSTM.forward (model/model.py:394-463) around PyG TransformerConv (model/model.py:51),
// Encoder/Decoder norms (model/seq2seq.py:59-66, 138-165).
#pragma once
#include "fused_fwd.inl"
#include "fused_tc.cuh"

namespace qmp {

struct TcStep {
    const uint8_t* img; uint32_t bytes;      // this conv's weight image
    const float* xin; int ld, D;             // its input rows
    int c;                                   // conv index in the group (column of logit / mstat / linv)
    uint32_t pcol;                           // column inside the P block it accumulates into
    int slot;                                // gate (gate mode) / conv (plain mode) the block belongs to
    bool segA, first, last;                  // first / last conv feeding that block
};

// one edge of the online segment softmax
template <int DC>
__device__ __forceinline__ void tc_edge(const FusedFwdArgs& a, int kk, int c, const float (&u)[DC], const float (&w01)[2],
                                        const float (&xj)[DC], float e0, float e1, float (&z)[DC], float& m, float& l, float& zs,
                                        float& ze0, float& ze1) {
    float s = fmaf(w01[0], e0, w01[1] * e1);
#pragma unroll
    for (int k = 0; k < DC; ++k) s = fmaf(u[k], xj[k], s);
    a.logit[(size_t)kk * a.NC + c] = s;
    const float mn = fmaxf(m, s);
    const float sc = __expf(m - mn), p = __expf(s - mn);
    const float pk = p * fdropout_scale(QMP_SEED_SM, (long long)kk * a.NC + c, a.drop_p);
    l = fmaf(l, sc, p);
    zs = fmaf(zs, sc, pk);
    ze0 = fmaf(ze0, sc, pk * e0);
    ze1 = fmaf(ze1, sc, pk * e1);
#pragma unroll
    for (int k = 0; k < DC; ++k) z[k] = fmaf(z[k], sc, pk * xj[k]);
    m = mn;
}

// `active` = false: this thread only takes part in the barriers / waits (paired-warp kernel: the second warp of a
// pair idles through the narrow X convs); it must then be called with valid = false and an empty edge list.
template <int DC>
__device__ __forceinline__ void conv_fwd_tc(TcCtx& cx, const FusedFwdArgs& a, int i, bool valid, const TcEdges& te,
                                            const TcStep& st, const TcStep& nx, bool has_next, bool active = true) {
    constexpr TcFwdLayout L(DC);
    const int t = threadIdx.x;
    const int buf = cx.toggle;
    uint8_t* wb = cx.wbase + (size_t)buf * cx.wslot;
    const float* __restrict__ xin = st.xin;
    const int ld = st.ld, D = st.D, c = st.c;
    constexpr bool vec = true;                 // the C entry point requires 16-byte aligned rows with D % 4 == 0
    // (1) own row (in flight while the previous conv's second contraction drains)
    float x[L.K1];
    constexpr bool co = DC == 32;              // coalesced row I/O (fused_tc.cuh) for the 32-float rows (D == 32 required)
    if constexpr (co) {
        warp_load_rows32(cx.rtile, xin, ld, valid ? i : -1, x);
    } else {
        tc_load_row<L.K1>(x, xin + (size_t)i * ld, D, vec, valid);
    }
    // (2) the previous conv's second contraction must be done with the A columns and with the other weight slot
    if (cx.pending) tc_wait(cx);
    if (t == 0 && has_next) tc_prefetch_image(cx, buf ^ 1, nx.img, nx.bytes);      // next conv's image, one conv ahead
    if (active) tc_stage_a_at<L.K1>(cx.lane_off, cx.ah_col, cx.al_col, x);
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tc::mbar_wait(cx.wfull + buf, (cx.wpar >> buf) & 1u);
        tc::fence_after_sync();
        tc_mma3_at(0, cx.u_col, cx.ah_col, cx.al_col, tc::smem_u32(wb + L.W1H), tc::smem_u32(wb + L.W1L), L.N1, L.K1, false);
        tc_mma3_at(0, cx.p_base + st.pcol, cx.ah_col, cx.al_col, tc::smem_u32(wb + L.W3H), tc::smem_u32(wb + L.W3L), FC, L.K1,
                   !st.first);
        tc::commit(cx.bar);
    }
    // (3) first pair of neighbour rows: issued before the wait for U
    const int k0 = te.k0, k1 = te.k1;
    float xa[DC], xb[DC];
    if constexpr (co) {
        if (__any_sync(0xffffffffu, k0 < k1))
            warp_load_rows32x2(cx.rtile, xin, ld, k0 < k1 ? te.j[0] : -1, k0 + 1 < k1 ? te.j[1] : -1, xa, xb);
    } else if (k0 < k1) {
        load_row<DC>(xa, xin + (size_t)te.j[0] * ld, D, vec);
        if (k0 + 1 < k1) load_row<DC>(xb, xin + (size_t)te.j[1] * ld, D, vec);
    }
    tc::mbar_wait(cx.wfull + buf, (cx.wpar >> buf) & 1u);       // biases below are read from the image
    tc::mbar_wait(cx.bar, cx.parity);
    cx.parity ^= 1;
    tc::fence_after_sync();
    // (4) u, w of this node
    float u[DC], w01[2];
    {
        constexpr int N8 = (DC + 2 + 7) / 8;
        float tmp[N8 * 8];
        if (active) tc_load_cols<N8>(cx.lane_off, cx.u_col, tmp);
        const float* b1 = reinterpret_cast<const float*>(wb + L.B1);
#pragma unroll
        for (int k = 0; k < DC; ++k) u[k] = tmp[k] + b1[k];
        w01[0] = tmp[DC] + b1[DC];
        w01[1] = tmp[DC + 1] + b1[DC + 1];
    }
    // (5) edge phase: online segment softmax over the in-edges, two rows in flight
    float z[DC];
#pragma unroll
    for (int k = 0; k < DC; ++k) z[k] = 0.f;
    float m = -INFINITY, l = 0.f, zs = 0.f, ze0 = 0.f, ze1 = 0.f;
    if constexpr (co) {
        {                                   // warp-cooperative gathers: every lane takes part, absent edges load nothing
            if (k0 < k1) tc_edge<DC>(a, k0, c, u, w01, xa, te.e0[0], te.e1[0], z, m, l, zs, ze0, ze1);
            if (k0 + 1 < k1) tc_edge<DC>(a, k0 + 1, c, u, w01, xb, te.e0[1], te.e1[1], z, m, l, zs, ze0, ze1);
            if (__any_sync(0xffffffffu, k0 + 2 < k1)) {
                warp_load_rows32x2(cx.rtile, xin, ld, k0 + 2 < k1 ? te.j[2] : -1, k0 + 3 < k1 ? te.j[3] : -1, xa, xb);
                if (k0 + 2 < k1) tc_edge<DC>(a, k0 + 2, c, u, w01, xa, te.e0[2], te.e1[2], z, m, l, zs, ze0, ze1);
                if (k0 + 3 < k1) tc_edge<DC>(a, k0 + 3, c, u, w01, xb, te.e0[3], te.e1[3], z, m, l, zs, ze0, ze1);
            }
            for (int kk = k0 + 4; __any_sync(0xffffffffu, kk < k1); ++kk) {      // larger in-degrees (quadtree meshes)
                const bool on = kk < k1;
                warp_load_rows32(cx.rtile, xin, ld, on ? a.nbr[kk] : -1, xa);
                if (on) {
                    const float e0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, e1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
                    tc_edge<DC>(a, kk, c, u, w01, xa, e0, e1, z, m, l, zs, ze0, ze1);
                }
            }
        }
    } else {
        if (k0 < k1) {                          // edges 0..3: indices and attributes were loaded once per tile
            tc_edge<DC>(a, k0, c, u, w01, xa, te.e0[0], te.e1[0], z, m, l, zs, ze0, ze1);
            if (k0 + 2 < k1) load_row<DC>(xa, xin + (size_t)te.j[2] * ld, D, vec);
            if (k0 + 1 < k1) tc_edge<DC>(a, k0 + 1, c, u, w01, xb, te.e0[1], te.e1[1], z, m, l, zs, ze0, ze1);
            if (k0 + 3 < k1) load_row<DC>(xb, xin + (size_t)te.j[3] * ld, D, vec);
            if (k0 + 2 < k1) tc_edge<DC>(a, k0 + 2, c, u, w01, xa, te.e0[2], te.e1[2], z, m, l, zs, ze0, ze1);
            if (k0 + 3 < k1) tc_edge<DC>(a, k0 + 3, c, u, w01, xb, te.e0[3], te.e1[3], z, m, l, zs, ze0, ze1);
        }
        for (int kk = k0 + 4; kk < k1; ++kk) {   // larger in-degrees (quadtree meshes)
            load_row<DC>(xa, xin + (size_t)a.nbr[kk] * ld, D, vec);
            const float e0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, e1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
            tc_edge<DC>(a, kk, c, u, w01, xa, e0, e1, z, m, l, zs, ze0, ze1);
        }
    }
    const float li = (l > 0.f) ? 1.f / l : 0.f;
    if (valid) {
        a.mstat[(size_t)i * a.NC + c] = m;
        a.linv[(size_t)i * a.NC + c] = li;
    }
    // (6) [z | ze | zs] -> A operand, second contraction accumulates into the P block
    {
        float zz[L.K2];
#pragma unroll
        for (int k = 0; k < L.K2; ++k) zz[k] = 0.f;
#pragma unroll
        for (int k = 0; k < DC; ++k) zz[k] = z[k] * li;
        zz[DC] = ze0 * li;
        zz[DC + 1] = ze1 * li;
        zz[DC + 2] = zs * li;
        if (active) tc_stage_a_at<L.K2>(cx.lane_off, cx.ah_col, cx.al_col, zz);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tc::fence_after_sync();
        tc_mma3_at(0, cx.p_base + st.pcol, cx.ah_col, cx.al_col, tc::smem_u32(wb + L.W2H), tc::smem_u32(wb + L.W2L), FC, L.K2, true);
        tc::commit(cx.bar);
    }
    cx.pending = true;
    cx.wpar ^= 1u << buf;
    cx.toggle ^= 1;
}

// P block -> registers, plus the skip biases of the convs that fed it
__device__ __forceinline__ void tc_collect(TcCtx& cx, uint32_t pcol, float (&P)[FC]) {
    if (cx.pending) tc_wait(cx);
    tc_load_cols<FC / 8>(cx.lane_off, cx.p_base + pcol, P);
}
__device__ __forceinline__ void tc_add_bias(float (&P)[FC], const uint8_t* __restrict__ img, int b3_off) {
    const float4* b = reinterpret_cast<const float4*>(img + b3_off);
#pragma unroll
    for (int o = 0; o < FC; o += 4) {
        const float4 v = __ldg(b + o / 4);
        P[o] += v.x; P[o + 1] += v.y; P[o + 2] += v.z; P[o + 3] += v.w;
    }
}

// Gate epilogue of slot s (same math as fused_fwd.inl gate_epilogue) with coalesced row I/O through the warp's row tile
// and the I / F / C' rows handed from slot to slot through 96 spare TMEM columns instead of global memory.
// params rows (lstm.cu): 0 wci 1 wcf 2 wco 3 bi 4 bf 5 bc 6 bo 7 gh 8 bh 9 gc 10 bc 11 go 12 bo
__device__ __forceinline__ void gate_epilogue_tc(TcCtx& cx, const FusedFwdArgs& a, int row0, int i, bool valid, int s,
                                                 const float* __restrict__ prm, float (&P)[FC]) {
    const uint32_t stash = cx.stash_col;
    float* tile = cx.rtile;
    if (s <= 2) {
        float cp[FC];
        if (a.Cprev) warp_load_rows32(tile, a.Cprev, FC, valid ? i : -1, cp);
        else {
#pragma unroll
            for (int o = 0; o < FC; ++o) cp[o] = 0.f;
        }
        if (s < 2) {                 // I, F
            const float* wc = prm + (s == 0 ? 0 : 1) * FC;
            const float* bb = prm + (s == 0 ? 3 : 4) * FC;
#pragma unroll
            for (int o = 0; o < FC; ++o) P[o] = sigm(P[o] + wc[o] * cp[o] + bb[o]);
            warp_store_rows32(tile, a.gates + s * FC, 4 * FC, row0, a.N, P);
            tc_store_cols<FC / 8>(cx.lane_off, stash + (uint32_t)s * FC, P);
            return;
        }
        // s == 2: T, then C' = F C + I T
        float I[FC], Fg[FC];
        tc_load_cols<FC / 8>(cx.lane_off, stash, I);
        tc_load_cols<FC / 8>(cx.lane_off, stash + FC, Fg);
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = ftanh(P[o] + prm[5 * FC + o]);
        warp_store_rows32(tile, a.gates + 2 * FC, 4 * FC, row0, a.N, P);
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = fmaf(Fg[o], cp[o], I[o] * P[o]);
        warp_store_rows32(tile, a.Craw, FC, row0, a.N, P);
        tc_store_cols<FC / 8>(cx.lane_off, stash + 2 * FC, P);
        return;
    }
    // s == 3: O, H', norms, head
    float Cn[FC];
    tc_load_cols<FC / 8>(cx.lane_off, stash + 2 * FC, Cn);
#pragma unroll
    for (int o = 0; o < FC; ++o) P[o] = sigm(P[o] + prm[2 * FC + o] * Cn[o] + prm[6 * FC + o]);   // O
    warp_store_rows32(tile, a.gates + 3 * FC, 4 * FC, row0, a.N, P);
    if (a.Oout) warp_store_rows32(tile, a.Oout, FC, row0, a.N, P);
    float mean, rstd;
    {
        float Hh[FC];
#pragma unroll
        for (int o = 0; o < FC; ++o) Hh[o] = P[o] * ftanh(Cn[o]);
        if (a.norm_h) {
            ln_stats(Hh, a.eps, mean, rstd);
#pragma unroll
            for (int o = 0; o < FC; ++o) Hh[o] = (Hh[o] - mean) * rstd * prm[7 * FC + o] + prm[8 * FC + o];
        }
        warp_store_rows32(tile, a.Hout, FC, row0, a.N, Hh);
    }
    if (a.norm_c) {
        ln_stats(Cn, a.eps, mean, rstd);
#pragma unroll
        for (int o = 0; o < FC; ++o) Cn[o] = (Cn[o] - mean) * rstd * prm[9 * FC + o] + prm[10 * FC + o];
    }
    warp_store_rows32(tile, a.Cout, FC, row0, a.N, Cn);
    if (a.head_in) {
        if (a.norm_o) {
            ln_stats(P, a.eps, mean, rstd);
#pragma unroll
            for (int o = 0; o < FC; ++o) P[o] = (P[o] - mean) * rstd * prm[11 * FC + o] + prm[12 * FC + o];
        }
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = fmaxf(P[o], 0.f);
        if (a.ldh % 4 == 0) warp_store_rows32(tile, a.head_in, a.ldh, row0, a.N, P);
        if (valid) {
            float* hr = a.head_in + (size_t)i * a.ldh;
            if (a.ldh % 4 != 0) store_row<FC>(hr, P, false);
            if (a.concat) hr[FC] = a.concat[i];
            for (int k = FC + 1; k < a.ldh; ++k) hr[k] = 0.f;
        }
    }
}

// conv schedule of a tile.  Gate mode: slot s = gate; its convs are [A[s]], B[s], [B[4+s]].  Plain mode: conv k.
template <int DA_, int DBC>
__device__ __forceinline__ void tc_fwd_step(const FusedFwdArgs& a, int k, TcStep& st) {
    constexpr TcFwdLayout LA(DA_), LB(DBC);
    const uint8_t* imgA = reinterpret_cast<const uint8_t*>(a.wa);
    const uint8_t* imgB = reinterpret_cast<const uint8_t*>(a.wb);
    int g;
    if (a.mode == 1) {
        const int hasA = a.GA ? 1 : 0, cps = hasA + (a.GB == 8 ? 2 : 1);
        const int s = k / cps, r = k - s * cps;
        st.segA = hasA && r == 0;
        g = st.segA ? s : s + 4 * (r - hasA);
        st.first = r == 0;
        st.last = r == cps - 1;
        st.pcol = 0;
        st.slot = s;
    } else {
        st.segA = k < a.GA;
        g = st.segA ? k : k - a.GA;
        st.first = st.last = true;
        st.pcol = 0;
        st.slot = k;
    }
    if (st.segA) {
        st.img = imgA + (size_t)g * LA.BYTES; st.bytes = LA.BYTES; st.xin = a.xa; st.ld = a.lda; st.D = a.DA; st.c = g;
    } else {
        st.img = imgB + (size_t)g * LB.BYTES; st.bytes = LB.BYTES; st.xin = a.xb + (a.sharedB ? 0 : g * a.DB); st.ld = a.ldb;
        st.D = a.DB; st.c = a.GA + g;
    }
}

template <int DAC, int DBC>
__global__ void __launch_bounds__(128, 2) fused_fwd_tc_kernel(const __grid_constant__ FusedFwdArgs a) {
    qmp_seed_init(a.seed, a.salt);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[3];
    __shared__ uint32_t tmem_slot[2];
    constexpr int DA_ = DAC > 0 ? DAC : 4;
    constexpr TcFwdLayout LA(DA_), LB(DBC);
    constexpr int SLOT = (LA.BYTES > LB.BYTES ? LA.BYTES : LB.BYTES);
    const int t = threadIdx.x, warp = t >> 5;
    float* prm = reinterpret_cast<float*>(smem + 2 * SLOT);
    if (t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_init(&bars[2], 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot[0], TC_COLS);    // U 48 | A_hi 40 | A_lo 40 | P 32 | stash 96 (I, F, C')
    if (a.mode == 1)
        for (int idx = t; idx < 13 * FC; idx += 128) prm[idx] = a.params[idx];
    pdl_wait();
    pdl_launch();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    TcCtx cx;
    cx.wbase = smem;
    cx.wslot = SLOT;
    cx.wfull = &bars[1];
    cx.wpar = 0;
    cx.toggle = 0;
    cx.bar = &bars[0];
    cx.parity = 0;
    cx.pending = false;
    cx.tmem = tmem_slot[0];
    cx.lane_off = (uint32_t)(warp * 32) << 16;
    cx.lane_base = cx.tmem + cx.lane_off;
    cx.u_col = cx.tmem + 0;
    cx.ah_col = cx.tmem + 48;
    cx.al_col = cx.tmem + 88;
    cx.p_base = cx.tmem + 128;
    cx.stash_col = cx.tmem + 160;
    cx.rtile = reinterpret_cast<float*>(smem + 2 * SLOT) + 13 * FC + warp * 2 * TC_ROWTILE;

    const int ntiles = (a.N + 127) / 128;
    const int nsteps = (a.mode == 1) ? 4 * ((a.GA ? 1 : 0) + (a.GB == 8 ? 2 : 1)) : a.NC;
    TcStep st, nx;
    if ((int)blockIdx.x < ntiles) {
        tc_fwd_step<DA_, DBC>(a, 0, st);
        if (t == 0) tc_prefetch_image(cx, 0, st.img, st.bytes);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * 128 + t;
        const bool valid = i < a.N;
        TcEdges te;
        tc_load_edges(te, a.ptr, a.nbr, a.ea, i, valid);
        for (int k = 0; k < nsteps; ++k) {
            tc_fwd_step<DA_, DBC>(a, k, st);
            const bool has_next = (k + 1 < nsteps) || (tile + (int)gridDim.x < ntiles);
            tc_fwd_step<DA_, DBC>(a, (k + 1 < nsteps) ? k + 1 : 0, nx);
            bool ranA = false;
            if constexpr (DAC > 0) {
                if (st.segA) {
                    conv_fwd_tc<DA_>(cx, a, i, valid, te, st, nx, has_next);
                    ranA = true;
                }
            }
            if (!ranA) conv_fwd_tc<DBC>(cx, a, i, valid, te, st, nx, has_next);
            if (!st.last) continue;
            float P[FC];
            tc_collect(cx, st.pcol, P);
            if (a.mode == 1) {
                const int s = st.slot;
                if constexpr (DAC > 0) {
                    if (a.GA) tc_add_bias(P, reinterpret_cast<const uint8_t*>(a.wa) + (size_t)s * LA.BYTES, LA.B3);
                }
                tc_add_bias(P, reinterpret_cast<const uint8_t*>(a.wb) + (size_t)s * LB.BYTES, LB.B3);
                if (a.GB == 8) tc_add_bias(P, reinterpret_cast<const uint8_t*>(a.wb) + (size_t)(4 + s) * LB.BYTES, LB.B3);
                gate_epilogue_tc(cx, a, tile * 128 + warp * 32, i, valid, s, prm, P);
            } else {
                tc_add_bias(P, st.img, st.segA ? LA.B3 : LB.B3);
                if (valid) {
                    float* orow = a.out + (size_t)i * a.ldo + (size_t)k * a.C;
                    if (a.relu_out) {
#pragma unroll
                        for (int o = 0; o < FC; ++o) P[o] = fmaxf(P[o], 0.f);
                    }
                    if (a.C == FC) {
                        store_row<FC>(orow, P, ((a.ldo % 4) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0));
                    } else {
#pragma unroll
                        for (int o = 0; o < FC; ++o)
                            if (o < a.C) orow[o] = P[o];
                    }
                }
            }
        }
    }
    if (cx.pending) tc_wait(cx);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(cx.tmem, TC_COLS);
}

template <int DAC, int DBC>
int launch_fwd_tc(const FusedFwdArgs& a, cudaStream_t st) {
    constexpr int DA_ = DAC > 0 ? DAC : 4;
    constexpr TcFwdLayout LA(DA_), LB(DBC);
    constexpr int SLOT = (LA.BYTES > LB.BYTES ? LA.BYTES : LB.BYTES);
    const size_t smem = 2 * (size_t)SLOT + (13 * FC + 4 * 2 * TC_ROWTILE) * sizeof(float);
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    auto kern = fused_fwd_tc_kernel<DAC, DBC>;
    QMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = cdiv(a.N, 128);
    const int grid = ntiles < 2 * n_sm ? ntiles : 2 * n_sm;
    QMP_CUDA(launch_pdl(kern, dim3(grid), dim3(128), smem, st, a));
    QMP_LAUNCH_CHECK("fused_fwd_tc_kernel");
    return 0;
}

}  // namespace qmp
