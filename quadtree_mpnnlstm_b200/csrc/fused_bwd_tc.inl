// tcgen05 versions of the fused backward kernels (same contract, arguments and outputs as fused_bwd.inl, which stays
// as the fp32-FFMA cross-check).  Thread = node = TMEM lane as in fused_fwd_tc.inl; per conv
//
//   target side:  g_i --st--> A --mma--> dz = W2^T g (logit-side gradient of [z | ze | zs]),  dx_i (+)= W3^T g
//                 dz --ld--> registers; two SIMT passes over the in-edges (d alpha, then ds / du / z);
//                 [du | dw] --st--> A --mma--> dx_i += W1^T [du | dw];  Zs / dUs rows written for qmp_fused_wgrad.
//   source side:  SIMT pass over the out-edges: av = sum alpha g_i, bv = sum ds x_i, sds = sum ds;
//                 [av | bv | sds] --st--> A --mma--> dx_j += W2[:, :D]^T av + W1[:D] bv + b1 sds.
// dx accumulates in TMEM over the convs that share an input and is written (target) / added (source) once.
#pragma once
#include "fused_bwd.inl"
#include "fused_tc.cuh"

namespace qmp {

constexpr uint32_t TCB_DXB = 0, TCB_DXA = 64;                           // target kernel: dx accumulators inside the P block
constexpr uint32_t TCB_UST = 80;                                        // one-pass mode: u_i stash (<= 40 columns of the P block)

// ONE-PASS mode (FusedBwdArgs::onepass, entry point qmp_fused_bwd_onepass_tc): the target kernel also does the source side
// of every in-edge j -> i, as in fused_cell_bwd.cu: the contribution ds_e u_i + alpha_e dz_i to dx_j is formed from
// target-side quantities only (u_i = W1[:DC] x_i + b1 recomputed with FFMAs from the plain copy in the image, dz_i read
// back from tensor memory) and added to row j with 16-byte vector reductions; the self part of dx_i goes the same way.
// dxa / dxb are zeroed by the entry point.  No out-CSR pass, no second launch, no second weight image.
__device__ __forceinline__ void tc_red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// lane's row x is ADDED to base + j * ld (j < 0: skipped) with coalesced reductions: 8 lanes per row, 4 rows per instruction
__device__ __forceinline__ void warp_red_rows32(float* tile, float* __restrict__ base, int ld, int j, const float (&x)[32]) {
    const int lane = threadIdx.x & 31, c = lane & 7;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(tc_tile_chunk(tile, lane, k)) = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = 4 * q + (lane >> 3);
        const int jr = __shfl_sync(0xffffffffu, j, r);
        if (jr >= 0) {
            const float4 v = *reinterpret_cast<const float4*>(tc_tile_chunk(tile, r, c));
            tc_red4(base + (size_t)jr * ld + 4 * c, v.x, v.y, v.z, v.w);
        }
    }
    __syncwarp();
}
constexpr uint32_t TCS_DXB = 0, TCS_DXA = 48, TCS_AH = 64, TCS_AL = 144;   // source kernel map (A up to 80 columns)

struct TcBStep {
    const uint8_t* img; uint32_t bytes;
    const float* xin; int ld, D;
    int c, gseg, G;                 // conv index in the group, index inside its segment, convs in the segment
    float* Zs; float* dUs;          // [N, G, cap+4]
    float* dx; int lddx;            // gradient rows of this conv's input (nullptr: not needed)
    uint32_t pcol;
    bool segA, first, last;
};

template <int DA_, int DBC, int KIND>
__device__ __forceinline__ void tc_bwd_step(const FusedBwdArgs& a, int k, TcBStep& st) {
    const uint8_t* imgA = reinterpret_cast<const uint8_t*>(a.wa);
    const uint8_t* imgB = reinterpret_cast<const uint8_t*>(a.wb);
    constexpr int BA = KIND == 1 ? TcBwdTLayout(DA_).BYTES : TcBwdSLayout(DA_).BYTES;
    constexpr int BB = KIND == 1 ? TcBwdTLayout(DBC).BYTES : TcBwdSLayout(DBC).BYTES;
    st.segA = k < a.GA;
    const int g = st.segA ? k : k - a.GA;
    st.c = k;
    st.gseg = g;
    if (st.segA) {
        st.img = imgA + (size_t)g * BA; st.bytes = BA; st.xin = a.xa; st.ld = a.lda; st.D = a.DA; st.G = a.GA;
        st.Zs = a.ZsA; st.dUs = a.dUsA; st.dx = a.need_dxa ? a.dxa : nullptr; st.lddx = a.lda;
        st.first = g == 0; st.last = g == a.GA - 1;
        st.pcol = KIND == 1 ? TCB_DXA : TCS_DXA;
    } else {
        const int off = a.sharedB ? 0 : g * a.DB;
        st.img = imgB + (size_t)g * BB; st.bytes = BB; st.xin = a.xb + off; st.ld = a.ldb; st.D = a.DB; st.G = a.GB;
        st.Zs = a.ZsB; st.dUs = a.dUsB; st.dx = a.need_dxb ? a.dxb + off : nullptr; st.lddx = a.ldb;
        st.first = a.sharedB ? g == 0 : true; st.last = a.sharedB ? g == a.GB - 1 : true;
        st.pcol = KIND == 1 ? TCB_DXB : TCS_DXB;
    }
}

// target kernel: (tile, conv) work items when no two convs share an input (encoder conv layers 1.., 8 blocks of 32)
__host__ __device__ __forceinline__ bool tc_bwd_conv_items(const FusedBwdArgs& a) { return a.GA == 0 && !a.sharedB && a.GB > 1; }

// ------------------------------------------------------------------------------------------------ target side
template <int DC>
__device__ __forceinline__ void conv_bwd_target_tc(TcCtx& cx, const FusedBwdArgs& a, int i, bool valid, const TcEdges& te,
                                                   const TcBStep& st, const TcBStep& nx, bool has_next) {
    constexpr TcBwdTLayout L(DC);
    const int t = threadIdx.x;
    const int buf = cx.toggle;
    uint8_t* wb = cx.wbase + (size_t)buf * cx.wslot;
    const float* __restrict__ xin = st.xin;
    const int ld = st.ld, D = st.D, c = st.c;
    constexpr bool vec = true;                 // the C entry point requires 16-byte aligned rows with D % 4 == 0
    constexpr bool co = DC == 32;              // coalesced row I/O (fused_tc.cuh) for the 32-float rows
    const bool need_dx = st.dx != nullptr;
    // neighbour rows: two in flight (xa, xb); absent edges are index -1
    float xa[DC], xb[DC];
    auto load_pair = [&](int ja, int jb) {
        if constexpr (co) {
            if (__any_sync(0xffffffffu, ja >= 0)) warp_load_rows32x2(cx.rtile, xin, ld, ja, jb, xa, xb);
        } else {
            if (ja >= 0) load_row<DC>(xa, xin + (size_t)ja * ld, D, vec);
            if (jb >= 0) load_row<DC>(xb, xin + (size_t)jb * ld, D, vec);
        }
    };
    auto load_one = [&](int j) {
        if constexpr (co) warp_load_rows32(cx.rtile, xin, ld, j, xa);
        else if (j >= 0) load_row<DC>(xa, xin + (size_t)j * ld, D, vec);
    };
    auto more = [&](bool pred) {               // loop condition: warp-uniform when the loads are cooperative
        if constexpr (co) return __any_sync(0xffffffffu, pred) != 0;
        else return pred;
    };
    {
        float g[FC];
        if (a.mode == 1 || a.C == FC) {        // full 32-float gradient rows: coalesced
            const int slot = (a.mode == 1) ? ((c < a.GA) ? c : ((c - a.GA) & 3)) : c;
            warp_load_rows32(cx.rtile_g, a.dP + (size_t)slot * FC, a.lddp, valid ? i : -1, g);
        } else if (valid) load_dP(g, a, i, c);
        else {
#pragma unroll
            for (int o = 0; o < FC; ++o) g[o] = 0.f;
        }
        if (cx.pending) tc_wait(cx);
        if (t == 0 && has_next) tc_prefetch_image(cx, buf ^ 1, nx.img, nx.bytes);
        tc_stage_a<FC>(cx.lane_base, g);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tc::mbar_wait(cx.wfull + buf, (cx.wpar >> buf) & 1u);
        tc::fence_after_sync();
        tc_mma3(cx.tmem, TC_U, tc::smem_u32(wb + L.W2TH), tc::smem_u32(wb + L.W2TL), L.N2, FC, false);
        if (need_dx) tc_mma3(cx.tmem, st.pcol, tc::smem_u32(wb + L.W3TH), tc::smem_u32(wb + L.W3TL), L.N1P, FC, !st.first);
        tc::commit(cx.bar);
    }
    const bool onep = need_dx && a.onepass;
    if (onep) {     // u_i = W1[:DC] x_i + b1 (what d logit_e / d x_j is for every in-edge e of this node) -> TMEM stash
        tc::mbar_wait(cx.wfull + buf, (cx.wpar >> buf) & 1u);
        float u[L.K1];
        const float* b1p = reinterpret_cast<const float*>(wb + L.B1P);
#pragma unroll
        for (int k = 0; k < L.K1; ++k) u[k] = b1p[k];
        const float4* w1p = reinterpret_cast<const float4*>(wb + L.W1P);
        const float4* xrow = reinterpret_cast<const float4*>(xin + (size_t)(valid ? i : 0) * ld);
#pragma unroll 1
        for (int m4 = 0; m4 < DC / 4; ++m4) {
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && 4 * m4 < D) xv = __ldg(xrow + m4);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const float4* wr = w1p + (4 * m4 + s4) * (L.K1 / 4);
#pragma unroll
                for (int q = 0; q < L.K1 / 4; ++q) {
                    const float4 w = wr[q];
                    u[4 * q] = fmaf(xs[s4], w.x, u[4 * q]);
                    u[4 * q + 1] = fmaf(xs[s4], w.y, u[4 * q + 1]);
                    u[4 * q + 2] = fmaf(xs[s4], w.z, u[4 * q + 2]);
                    u[4 * q + 3] = fmaf(xs[s4], w.w, u[4 * q + 3]);
                }
            }
        }
        tc_store_cols<L.K1 / 8>(cx.lane_base, TCB_UST, u);
    }
    const int k0 = te.k0, k1 = te.k1;
    const int j0 = k0 < k1 ? te.j[0] : -1, j1 = k0 + 1 < k1 ? te.j[1] : -1, j2 = k0 + 2 < k1 ? te.j[2] : -1,
              j3 = k0 + 3 < k1 ? te.j[3] : -1;
    load_pair(j0, j1);
    const float m = valid ? a.mstat[(size_t)i * a.NC + c] : 0.f, li = valid ? a.linv[(size_t)i * a.NC + c] : 0.f;
    float al4[4], dal4[4];                   // alpha and d alpha of edges 0..3 stay in registers between the passes
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        al4[e] = (k0 + e < k1) ? __expf(a.logit[(size_t)(k0 + e) * a.NC + c] - m) * li : 0.f;
        dal4[e] = 0.f;
    }
    tc::mbar_wait(cx.bar, cx.parity);
    cx.parity ^= 1;
    tc::fence_after_sync();
    float dz[DC + 4];
    {
        constexpr int N8 = (DC + 3 + 7) / 8;
        float tmp[N8 * 8];
        tc_load_cols<N8>(cx.lane_base, TC_U, tmp);
#pragma unroll
        for (int k = 0; k < DC + 3; ++k) dz[k] = tmp[k];
        dz[DC + 3] = 0.f;
    }
    // pass 1: d alpha per edge, t = sum alpha d alpha
    float tsum = 0.f;
    auto dalpha = [&](int e, const float(&xj)[DC], float a0, float a1) {
        float dal = fmaf(dz[DC], a0, fmaf(dz[DC + 1], a1, dz[DC + 2]));
#pragma unroll
        for (int k = 0; k < DC; ++k) dal = fmaf(dz[k], xj[k], dal);
        return dal * fdropout_scale(QMP_SEED_SM, (long long)e * a.NC + c, a.drop_p);
    };
    if (j0 >= 0) dal4[0] = dalpha(k0, xa, te.e0[0], te.e1[0]);
    if (j1 >= 0) dal4[1] = dalpha(k0 + 1, xb, te.e0[1], te.e1[1]);
    load_pair(j2, j3);
    if (j2 >= 0) dal4[2] = dalpha(k0 + 2, xa, te.e0[2], te.e1[2]);
    if (j3 >= 0) dal4[3] = dalpha(k0 + 3, xb, te.e0[3], te.e1[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e) tsum = fmaf(al4[e], dal4[e], tsum);
    for (int kk = k0 + 4; more(kk < k1); ++kk) {   // larger in-degrees: d alpha stashed in ds
        const bool on = kk < k1;
        load_one(on ? a.nbr[kk] : -1);
        if (!on) continue;
        const float a0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, a1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        const float dal = dalpha(kk, xa, a0, a1);
        tsum = fmaf(__expf(a.logit[(size_t)kk * a.NC + c] - m) * li, dal, tsum);
        a.ds[(size_t)kk * a.NC + c] = dal;
    }
    // pass 2: ds, du = sum ds x_j, z = sum alpha x_j
    float du[L.K2], z[DC + 4];
#pragma unroll
    for (int k = 0; k < L.K2; ++k) du[k] = 0.f;
#pragma unroll
    for (int k = 0; k < DC + 4; ++k) z[k] = 0.f;
    auto accum = [&](int e, const float(&xj)[DC], float a0, float a1, float al, float dal) {
        const float dsv = al * (dal - tsum);
        a.ds[(size_t)e * a.NC + c] = dsv;
        const float alk = al * fdropout_scale(QMP_SEED_SM, (long long)e * a.NC + c, a.drop_p);
#pragma unroll
        for (int k = 0; k < DC; ++k) {
            du[k] = fmaf(dsv, xj[k], du[k]);
            z[k] = fmaf(alk, xj[k], z[k]);
        }
        du[DC] = fmaf(dsv, a0, du[DC]);
        du[DC + 1] = fmaf(dsv, a1, du[DC + 1]);
        z[DC] = fmaf(alk, a0, z[DC]);
        z[DC + 1] = fmaf(alk, a1, z[DC + 1]);
        z[DC + 2] += alk;
    };
    // xa / xb hold edges 2, 3 (when they exist and no long-degree loop reused xa): finish them, then re-gather 0, 1
    if (more(k0 + 4 < k1)) load_pair(j2, j3);
    if (j2 >= 0) accum(k0 + 2, xa, te.e0[2], te.e1[2], al4[2], dal4[2]);
    if (j3 >= 0) accum(k0 + 3, xb, te.e0[3], te.e1[3], al4[3], dal4[3]);
    if (more(j2 >= 0)) load_pair(j0, j1);     // nodes with at most two in-edges still hold edges 0, 1
    if (j0 >= 0) accum(k0, xa, te.e0[0], te.e1[0], al4[0], dal4[0]);
    if (j1 >= 0) accum(k0 + 1, xb, te.e0[1], te.e1[1], al4[1], dal4[1]);
    for (int kk = k0 + 4; more(kk < k1); ++kk) {
        const bool on = kk < k1;
        load_one(on ? a.nbr[kk] : -1);
        if (!on) continue;
        const float a0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, a1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        accum(kk, xa, a0, a1, __expf(a.logit[(size_t)kk * a.NC + c] - m) * li, a.ds[(size_t)kk * a.NC + c]);
    }
    if (valid) {
        store_row<DC + 4>(st.Zs + ((size_t)i * st.G + st.gseg) * (DC + 4), z, true);
        float* dr = st.dUs + ((size_t)i * st.G + st.gseg) * (DC + 4);
#pragma unroll
        for (int k = 0; k < DC + 4; k += 4) *reinterpret_cast<float4*>(dr + k) = make_float4(du[k], du[k + 1], du[k + 2], du[k + 3]);
    }
    if (need_dx) {
        tc_stage_a<L.K2>(cx.lane_base, du);
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncthreads();
        if (t == 0) {
            tc::fence_after_sync();
            tc_mma3(cx.tmem, st.pcol, tc::smem_u32(wb + L.W1TH), tc::smem_u32(wb + L.W1TL), L.N1P, L.K2, true);
            tc::commit(cx.bar);
        }
        cx.pending = true;
        if (onep) {             // source side of this node's in-edges, under the MMA: dx_j += ds_e u_i + alpha_e dz_i
            float u[L.K1], dzr[L.K1];
            tc_load_cols<L.K1 / 8>(cx.lane_base, TCB_UST, u);
            tc_load_cols<L.K1 / 8>(cx.lane_base, TC_U, dzr);
            auto push = [&](int j, float dsv, float alk) {
                if constexpr (co) {
                    float r[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) r[k] = fmaf(dsv, u[k], alk * dzr[k]);
                    warp_red_rows32(cx.rtile, st.dx, st.lddx, j, r);
                } else if (j >= 0) {
                    float* dr = st.dx + (size_t)j * st.lddx;
#pragma unroll
                    for (int k = 0; k < DC; k += 4)
                        if (k < D)
                            tc_red4(dr + k, fmaf(dsv, u[k], alk * dzr[k]), fmaf(dsv, u[k + 1], alk * dzr[k + 1]),
                                    fmaf(dsv, u[k + 2], alk * dzr[k + 2]), fmaf(dsv, u[k + 3], alk * dzr[k + 3]));
                }
            };
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = e == 0 ? j0 : e == 1 ? j1 : e == 2 ? j2 : j3;
                if (more(j >= 0))
                    push(j, al4[e] * (dal4[e] - tsum), al4[e] * fdropout_scale(QMP_SEED_SM, (long long)(k0 + e) * a.NC + c, a.drop_p));
            }
            for (int kk = k0 + 4; more(kk < k1); ++kk) {
                const bool on = kk < k1;
                float dsv = 0.f, alk = 0.f;
                if (on) {
                    dsv = a.ds[(size_t)kk * a.NC + c];
                    alk = __expf(a.logit[(size_t)kk * a.NC + c] - m) * li * fdropout_scale(QMP_SEED_SM, (long long)kk * a.NC + c, a.drop_p);
                }
                push(on ? a.nbr[kk] : -1, dsv, alk);
            }
        }
        if (st.last && onep) {  // self part added like every other contribution (rows are zero on entry, any CTA order)
            tc_wait(cx);
            float dx[L.K1];
            tc_load_cols<L.K1 / 8>(cx.lane_base, st.pcol, dx);
            if constexpr (co) {
                warp_red_rows32(cx.rtile, st.dx, st.lddx, valid ? i : -1, dx);
            } else if (valid) {
                float* row = st.dx + (size_t)i * st.lddx;
#pragma unroll
                for (int k = 0; k < L.K1; k += 4)
                    if (k < D) tc_red4(row + k, dx[k], dx[k + 1], dx[k + 2], dx[k + 3]);
            }
        } else if (st.last) {   // dx_i (self part) = accumulated over the convs sharing this input
            tc_wait(cx);
            float dx[L.K1];
            tc_load_cols<L.K1 / 8>(cx.lane_base, st.pcol, dx);
            if constexpr (co) {
                warp_store_rows32(cx.rtile, st.dx, st.lddx, (i & ~31), a.N, dx);
            } else if (valid) {
                float* row = st.dx + (size_t)i * st.lddx;
#pragma unroll
                for (int k = 0; k < L.K1; ++k)
                    if (k < D) row[k] = dx[k];
            }
        }
    }
    cx.wpar ^= 1u << buf;
    cx.toggle ^= 1;
}

// ------------------------------------------------------------------------------------------------ source side
template <int DC>
__device__ __forceinline__ void conv_bwd_source_tc(TcCtx& cx, const FusedBwdArgs& a, int j, bool valid, const TcEdges& te,
                                                   const TcBStep& st, const TcBStep& nx, bool has_next) {
    constexpr TcBwdSLayout L(DC);
    const int t = threadIdx.x;
    const int buf = cx.toggle;
    uint8_t* wb = cx.wbase + (size_t)buf * cx.wslot;
    const float* __restrict__ xin = st.xin;
    const int ld = st.ld, D = st.D, c = st.c;
    constexpr bool vec = true;
    constexpr bool co = DC == 32;
    const bool gfull = a.mode == 1 || a.C == FC;                 // full 32-float gradient rows
    const int gslot = (a.mode == 1) ? ((c < a.GA) ? c : ((c - a.GA) & 3)) : c;
    float A[L.KS];
#pragma unroll
    for (int k = 0; k < L.KS; ++k) A[k] = 0.f;
    const int k0 = te.k0, k1 = te.k1;
    auto edge = [&](int i, int kin, const float(&g)[FC], const float(&xi)[DC], float al, float dsv) {
#pragma unroll
        for (int o = 0; o < FC; ++o) A[o] = fmaf(al, g[o], A[o]);
#pragma unroll
        for (int k = 0; k < DC; ++k) A[FC + k] = fmaf(dsv, xi[k], A[FC + k]);
        A[FC + L.K1] += dsv;
    };
    auto coef = [&](int i, int kin, float& al, float& dsv) {
        al = __expf(a.logit[(size_t)kin * a.NC + c] - a.mstat[(size_t)i * a.NC + c]) * a.linv[(size_t)i * a.NC + c] *
             fdropout_scale(QMP_SEED_SM, (long long)kin * a.NC + c, a.drop_p);
        dsv = a.ds[(size_t)kin * a.NC + c];
    };
    // out-edges 0..3: targets and in-CSR slots were loaded once per tile (te.j = target, te.e0 = slot as int bits)
    // one out-edge at a time: the gradient row of its target and the target's input row travel together
    auto one_edge = [&](bool on, int i, int kin) {
        float g[FC], xi[DC];
        if (co && gfull) {
            if constexpr (co) warp_load_rows32_ab(cx.rtile, a.dP + (size_t)gslot * FC, a.lddp, xin, ld, i, g, xi);
        } else {
            if (on) load_dP(g, a, i, c);
            if constexpr (co) warp_load_rows32(cx.rtile, xin, ld, i, xi);
            else if (on) load_row<DC>(xi, xin + (size_t)i * ld, D, vec);
        }
        if (on) {
            float al, dsv;
            coef(i, kin, al, dsv);
            edge(i, kin, g, xi, al, dsv);
        }
    };
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool on = k0 + e < k1;
        bool any = on;
        if constexpr (co) any = __any_sync(0xffffffffu, on) != 0;
        if (any) one_edge(on, on ? te.j[e] : -1, __float_as_int(te.e0[e]));
    }
    for (int kk = k0 + 4; co ? (__any_sync(0xffffffffu, kk < k1) != 0) : (kk < k1); ++kk) {
        const bool on = kk < k1;
        one_edge(on, on ? a.nbr[kk] : -1, on ? a.kin[kk] : 0);
    }
    if (cx.pending) tc_wait(cx);
    if (t == 0 && has_next) tc_prefetch_image(cx, buf ^ 1, nx.img, nx.bytes);
    tc_stage_a_at<L.KS>(cx.lane_base, TCS_AH, TCS_AL, A);
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tc::mbar_wait(cx.wfull + buf, (cx.wpar >> buf) & 1u);
        tc::fence_after_sync();
        tc_mma3_at(cx.tmem, st.pcol, TCS_AH, TCS_AL, tc::smem_u32(wb + L.BSH), tc::smem_u32(wb + L.BSL), L.N1P, L.KS, !st.first);
        tc::commit(cx.bar);
    }
    cx.pending = true;
    if (st.last) {
        tc_wait(cx);
        float dx[L.K1];
        tc_load_cols<L.K1 / 8>(cx.lane_base, st.pcol, dx);
        if constexpr (co) {                      // dx_j += ... : coalesced read-modify-write of the warp's 32 rows
            float cur[L.K1];
            warp_load_rows32(cx.rtile, st.dx, st.lddx, valid ? j : -1, cur);
#pragma unroll
            for (int k = 0; k < L.K1; ++k) dx[k] += cur[k];
            warp_store_rows32(cx.rtile, st.dx, st.lddx, (j & ~31), a.N, dx);
        } else if (valid) {
            float* row = st.dx + (size_t)j * st.lddx;
#pragma unroll
            for (int k = 0; k < L.K1; ++k)
                if (k < D) row[k] += dx[k];
        }
    }
    cx.wpar ^= 1u << buf;
    cx.toggle ^= 1;
}

template <int DAC, int DBC, int KIND>
__global__ void __launch_bounds__(128, 2) fused_bwd_tc_kernel(const __grid_constant__ FusedBwdArgs a) {
    qmp_seed_init(a.seed, a.salt);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[3];
    __shared__ uint32_t tmem_slot;
    constexpr int DA_ = DAC > 0 ? DAC : 4;
    constexpr int BA = KIND == 1 ? TcBwdTLayout(DA_).BYTES : TcBwdSLayout(DA_).BYTES;
    constexpr int BB = KIND == 1 ? TcBwdTLayout(DBC).BYTES : TcBwdSLayout(DBC).BYTES;
    constexpr int SLOT = BA > BB ? BA : BB;
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_init(&bars[2], 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, TC_COLS);
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
    cx.tmem = tmem_slot;
    cx.lane_base = cx.tmem + ((uint32_t)(warp * 32) << 16);
    cx.lane_off = (uint32_t)(warp * 32) << 16;
    constexpr int RT = DBC == 32 ? 2 : 1;          // the paired row loads exist for the 32-wide rows only
    cx.rtile = reinterpret_cast<float*>(smem + 2 * SLOT) + warp * RT * TC_ROWTILE;
    cx.rtile_g = cx.rtile;

    const int ntiles = (a.N + 127) / 128;
    // source side: only the convs whose input needs a gradient
    const int kfirst = (KIND == 2 && !a.need_dxa) ? a.GA : 0;
    const int kend = (KIND == 2 && !a.need_dxb) ? a.GA : a.NC;
    TcBStep st, nx;
    if constexpr (KIND == 1 && DAC == 0) {
        if (tc_bwd_conv_items(a)) {
            // every conv reads its own input block and owns its own gradient block: the work items are (tile, conv) pairs,
            // so 369 tiles x 8 convs spread over the 296 resident CTAs in 10 rounds of one conv instead of 2 rounds of 8
            const int items = ntiles * a.NC;
            if ((int)blockIdx.x < items) {
                tc_bwd_step<DA_, DBC, KIND>(a, (int)blockIdx.x % a.NC, st);
                if (t == 0) tc_prefetch_image(cx, 0, st.img, st.bytes);
            }
            for (int w = blockIdx.x; w < items; w += gridDim.x) {
                const int i = (w / a.NC) * 128 + t;
                const bool valid = i < a.N;
                TcEdges te;
                tc_load_edges(te, a.ptr, a.nbr, a.ea, i, valid);
                const int wn = w + (int)gridDim.x;
                tc_bwd_step<DA_, DBC, KIND>(a, w % a.NC, st);
                tc_bwd_step<DA_, DBC, KIND>(a, (wn < items ? wn : w) % a.NC, nx);
                conv_bwd_target_tc<DBC>(cx, a, i, valid, te, st, nx, wn < items);
            }
            if (cx.pending) tc_wait(cx);
            tc::fence_before_sync();
            __syncthreads();
            if (warp == 0) tc::tmem_dealloc(cx.tmem, TC_COLS);
            return;
        }
    }
    if ((int)blockIdx.x < ntiles && kfirst < kend) {
        tc_bwd_step<DA_, DBC, KIND>(a, kfirst, st);
        if (t == 0) tc_prefetch_image(cx, 0, st.img, st.bytes);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * 128 + t;
        const bool valid = i < a.N;
        TcEdges te;
        if constexpr (KIND == 1) tc_load_edges(te, a.ptr, a.nbr, a.ea, i, valid);
        else tc_load_out_edges(te, a.ptr, a.nbr, a.kin, i, valid);
        for (int k = kfirst; k < kend; ++k) {
            tc_bwd_step<DA_, DBC, KIND>(a, k, st);
            const bool has_next = (k + 1 < kend) || (tile + (int)gridDim.x < ntiles);
            tc_bwd_step<DA_, DBC, KIND>(a, (k + 1 < kend) ? k + 1 : kfirst, nx);
            bool ranA = false;
            if constexpr (DAC > 0) {
                if (st.segA) {
                    if constexpr (KIND == 1) conv_bwd_target_tc<DA_>(cx, a, i, valid, te, st, nx, has_next);
                    else conv_bwd_source_tc<DA_>(cx, a, i, valid, te, st, nx, has_next);
                    ranA = true;
                }
            }
            if (!ranA) {
                if constexpr (KIND == 1) conv_bwd_target_tc<DBC>(cx, a, i, valid, te, st, nx, has_next);
                else conv_bwd_source_tc<DBC>(cx, a, i, valid, te, st, nx, has_next);
            }
        }
    }
    if (cx.pending) tc_wait(cx);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(cx.tmem, TC_COLS);
}

template <int DAC, int DBC, int KIND>
int launch_bwd_tc(const FusedBwdArgs& a, cudaStream_t st) {
    constexpr int DA_ = DAC > 0 ? DAC : 4;
    constexpr int BA = KIND == 1 ? TcBwdTLayout(DA_).BYTES : TcBwdSLayout(DA_).BYTES;
    constexpr int BB = KIND == 1 ? TcBwdTLayout(DBC).BYTES : TcBwdSLayout(DBC).BYTES;
    constexpr int SLOT = BA > BB ? BA : BB;
    const size_t smem = 2 * (size_t)SLOT + 4 * (DBC == 32 ? 2 : 1) * TC_ROWTILE * sizeof(float);
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    auto kern = fused_bwd_tc_kernel<DAC, DBC, KIND>;
    QMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = cdiv(a.N, 128);
    const int work = (KIND == 1 && DAC == 0 && tc_bwd_conv_items(a)) ? ntiles * a.NC : ntiles;
    const int grid = work < 2 * n_sm ? work : 2 * n_sm;
    QMP_CUDA(launch_pdl(kern, dim3(grid), dim3(128), smem, st, a));
    QMP_LAUNCH_CHECK("fused_bwd_tc_kernel");
    return 0;
}

}  // namespace qmp
