// Paired-warp version of the tcgen05 fused forward (fused_fwd_tc.inl) for 32-wide inputs: TWO threads per node.
//
// ncu on the one-thread-per-node kernel (profiles/r01_g_ncu_fused_fwd_tc.txt): 7 warps / SM, ~16 k instructions per warp
// and tile at 0.25 IPC -- bound by single-warp instruction latency, and the 255 registers + 256 TMEM columns per tile
// allow no more warps.  Here a CTA has 256 threads: warps w and w + 4 own the SAME tensor-memory lane quarter (the same
// 32 nodes) and split the 32 feature columns in halves of 16.  Each thread stages / gathers / accumulates 16 columns
// (half the registers, so two 256-thread CTAs still fit an SM: 16 warps), reads its half of U and P from TMEM, and the
// two partial logits of an edge meet through shared memory (one 64-thread named barrier per pair of edges).  The gate
// epilogue works on 16 channels per thread; LayerNorm statistics are exchanged the same way.  The narrow X convs
// (D <= 8) run one-thread-per-node on the first warp of each pair (conv_fwd_tc with active = false on the second).
#pragma once
#include "fused_fwd_tc.inl"

namespace qmp {

constexpr int PW_H = 16;                        // columns per thread
constexpr int PW_TILE = 32 * PW_H;              // floats per half-row tile (32 rows x 16 floats)

// ---- coalesced half-row I/O: 4 lanes per 64-byte half row, 8 rows per instruction; chunk XOR-swizzle by (row >> 1) & 3
__device__ __forceinline__ float* pw_chunk(float* tile, int r, int c) { return tile + r * PW_H + ((c ^ ((r >> 1) & 3)) << 2); }

__device__ __forceinline__ void pw_load_rows(float* tile, const float* __restrict__ base, int ld, int j, float (&x)[PW_H]) {
    const int lane = threadIdx.x & 31, c = lane & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = 8 * q + (lane >> 2);
        const int jr = __shfl_sync(0xffffffffu, j, r);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jr >= 0) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)jr * ld) + c);
        *reinterpret_cast<float4*>(pw_chunk(tile, r, c)) = v;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(pw_chunk(tile, lane, k));
        x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
    }
    __syncwarp();
}

__device__ __forceinline__ void pw_load_rows_x2(float* tile, const float* __restrict__ base, int ld, int j0, int j1,
                                                float (&x0)[PW_H], float (&x1)[PW_H]) {
    const int lane = threadIdx.x & 31, c = lane & 3;
    float4 v0[4], v1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = 8 * q + (lane >> 2);
        const int ja = __shfl_sync(0xffffffffu, j0, r), jb = __shfl_sync(0xffffffffu, j1, r);
        v0[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ja >= 0) v0[q] = __ldg(reinterpret_cast<const float4*>(base + (size_t)ja * ld) + c);
        if (jb >= 0) v1[q] = __ldg(reinterpret_cast<const float4*>(base + (size_t)jb * ld) + c);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = 8 * q + (lane >> 2);
        *reinterpret_cast<float4*>(pw_chunk(tile, r, c)) = v0[q];
        *reinterpret_cast<float4*>(pw_chunk(tile + PW_TILE, r, c)) = v1[q];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(pw_chunk(tile, lane, k));
        const float4 b = *reinterpret_cast<const float4*>(pw_chunk(tile + PW_TILE, lane, k));
        x0[4 * k] = a.x; x0[4 * k + 1] = a.y; x0[4 * k + 2] = a.z; x0[4 * k + 3] = a.w;
        x1[4 * k] = b.x; x1[4 * k + 1] = b.y; x1[4 * k + 2] = b.z; x1[4 * k + 3] = b.w;
    }
    __syncwarp();
}

__device__ __forceinline__ void pw_store_rows(float* tile, float* __restrict__ base, int ld, int row0, int n_rows,
                                              const float (&x)[PW_H]) {
    const int lane = threadIdx.x & 31, c = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4*>(pw_chunk(tile, lane, k)) = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = 8 * q + (lane >> 2);
        if (row0 + r < n_rows)
            *(reinterpret_cast<float4*>(base + (size_t)(row0 + r) * ld) + c) = *reinterpret_cast<const float4*>(pw_chunk(tile, r, c));
    }
    __syncwarp();
}

#ifdef QMP_PW_TRACE
#define PW_MARK(tag)                                                                          \
    do {                                                                                      \
        if (a.mode == 1 && a.out && blockIdx.x == 0 && threadIdx.x == 0) {                    \
            int n__ = (int)a.out[0];                                                          \
            if (n__ < 1000) {                                                                 \
                a.out[1 + 2 * n__] = (float)(tag);                                            \
                a.out[2 + 2 * n__] = (float)((unsigned)clock64() & 0xFFFFFFu);                \
                a.out[0] = (float)(n__ + 1);                                                  \
            }                                                                                 \
        }                                                                                     \
    } while (0)
#else
#define PW_MARK(tag) do {} while (0)
#endif

struct PwCtx {
    int h, q;                 // half (0 / 1) and lane quarter (0..3) of this warp
    float* ex;                // [2][256] exchange slots (shared memory)
    float* htile;             // this warp's two half-row tiles
};

__device__ __forceinline__ void pw_pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); }

// sum of a per-thread partial with the partner thread's (same node, other half); all threads of the pair call it
__device__ __forceinline__ float pw_pair_sum(const PwCtx& pw, float v, int slot) {
    const int t = threadIdx.x;
    pw.ex[slot * 256 + t] = v;
    pw_pair_sync(pw.q);
    const float o = pw.ex[slot * 256 + (t ^ 128)];
    pw_pair_sync(pw.q);
    return v + o;
}

// one edge, this thread's 16 columns: s = full logit (already summed over the pair)
__device__ __forceinline__ void pw_edge(const FusedFwdArgs& a, int kk, int c, bool writer, float s, const float (&xj)[PW_H],
                                        float e0, float e1, float (&z)[PW_H], float& m, float& l, float& zs, float& ze0,
                                        float& ze1) {
    if (writer) a.logit[(size_t)kk * a.NC + c] = s;
    const float mn = fmaxf(m, s);
    const float sc = __expf(m - mn), p = __expf(s - mn);
    const float pk = p * fdropout_scale(QMP_SEED_SM, (long long)kk * a.NC + c, a.drop_p);
    l = fmaf(l, sc, p);
    zs = fmaf(zs, sc, pk);
    ze0 = fmaf(ze0, sc, pk * e0);
    ze1 = fmaf(ze1, sc, pk * e1);
#pragma unroll
    for (int k = 0; k < PW_H; ++k) z[k] = fmaf(z[k], sc, pk * xj[k]);
    m = mn;
}

__device__ __forceinline__ float pw_dot(const float (&u)[PW_H], const float (&x)[PW_H], float init) {
    float s = init;
#pragma unroll
    for (int k = 0; k < PW_H; ++k) s = fmaf(u[k], x[k], s);
    return s;
}

// a 32-wide conv, two threads per node
__device__ __forceinline__ void conv_fwd_pw(TcCtx& cx, const PwCtx& pw, const FusedFwdArgs& a, int i, bool valid,
                                            const TcEdges& te, const TcStep& st, const TcStep& nx, bool has_next) {
    constexpr TcFwdLayout L(32);
    const int t = threadIdx.x, h = pw.h;
    const int buf = cx.toggle;
    uint8_t* wb = cx.wbase + (size_t)buf * cx.wslot;
    const float* __restrict__ xin = st.xin + h * PW_H;          // this thread's half of every row
    const int ld = st.ld, c = st.c;
    float* tile = pw.htile;
    // (1) own half row
    float x[PW_H];
    PW_MARK(10);
    pw_load_rows(tile, xin, ld, valid ? i : -1, x);
    PW_MARK(11);
    if (cx.pending) tc_wait(cx);
    PW_MARK(12);
    if (t == 0 && has_next) tc_prefetch_image(cx, buf ^ 1, nx.img, nx.bytes);
    tc_stage_a_at<PW_H>(cx.lane_off, cx.ah_col + h * PW_H, cx.al_col + h * PW_H, x);
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    PW_MARK(13);
    if (t == 0) {
        tc::mbar_wait(cx.wfull + buf, (cx.wpar >> buf) & 1u);
        PW_MARK(14);
        tc::fence_after_sync();
        tc_mma3_at(0, cx.u_col, cx.ah_col, cx.al_col, tc::smem_u32(wb + L.W1H), tc::smem_u32(wb + L.W1L), L.N1, L.K1, false);
        tc_mma3_at(0, cx.p_base + st.pcol, cx.ah_col, cx.al_col, tc::smem_u32(wb + L.W3H), tc::smem_u32(wb + L.W3L), FC, L.K1,
                   !st.first);
        tc::commit(cx.bar);
    }
    PW_MARK(15);
    // (3) first pair of neighbour half rows, before the wait for U
    const int k0 = te.k0, k1 = te.k1;
    float xa[PW_H], xb[PW_H];
    if (__any_sync(0xffffffffu, k0 < k1)) pw_load_rows_x2(tile, xin, ld, k0 < k1 ? te.j[0] : -1, k0 + 1 < k1 ? te.j[1] : -1, xa, xb);
    tc::mbar_wait(cx.wfull + buf, (cx.wpar >> buf) & 1u);
    PW_MARK(16);
    tc::mbar_wait(cx.bar, cx.parity);
    cx.parity ^= 1;
    tc::fence_after_sync();
    PW_MARK(17);
    // (4) this thread's half of u, and w
    float u[PW_H], w01[2];
    {
        float tmp[PW_H], tw[8];
        tc_load_cols<2>(cx.lane_off, cx.u_col + h * PW_H, tmp);
        tc_load_cols<1>(cx.lane_off, cx.u_col + 32, tw);
        const float* b1 = reinterpret_cast<const float*>(wb + L.B1);
#pragma unroll
        for (int k = 0; k < PW_H; ++k) u[k] = tmp[k] + b1[h * PW_H + k];
        w01[0] = tw[0] + b1[32];
        w01[1] = tw[1] + b1[33];
    }
    // (5) edge phase: partial logits of two edges -> pair exchange -> online softmax on this thread's 16 columns
    float z[PW_H];
#pragma unroll
    for (int k = 0; k < PW_H; ++k) z[k] = 0.f;
    float m = -INFINITY, l = 0.f, zs = 0.f, ze0 = 0.f, ze1 = 0.f;
    const bool writer = h == 0;
    auto pair = [&](int e, bool on0, bool on1, float ea0, float ea1, float eb0, float eb1) {
        // the edge-attribute term is added once (by the first half)
        const float sa = pw_dot(u, xa, writer ? fmaf(w01[0], ea0, w01[1] * ea1) : 0.f);
        const float sb = pw_dot(u, xb, writer ? fmaf(w01[0], eb0, w01[1] * eb1) : 0.f);
        const int tt = threadIdx.x;
        pw.ex[tt] = sa;
        pw.ex[256 + tt] = sb;
        pw_pair_sync(pw.q);
        const float fa = sa + pw.ex[tt ^ 128], fb = sb + pw.ex[256 + (tt ^ 128)];
        pw_pair_sync(pw.q);
        if (on0) pw_edge(a, e, c, writer, fa, xa, ea0, ea1, z, m, l, zs, ze0, ze1);
        if (on1) pw_edge(a, e + 1, c, writer, fb, xb, eb0, eb1, z, m, l, zs, ze0, ze1);
    };
    if (__any_sync(0xffffffffu, k0 < k1)) pair(k0, k0 < k1, k0 + 1 < k1, te.e0[0], te.e1[0], te.e0[1], te.e1[1]);
    if (__any_sync(0xffffffffu, k0 + 2 < k1)) {
        pw_load_rows_x2(tile, xin, ld, k0 + 2 < k1 ? te.j[2] : -1, k0 + 3 < k1 ? te.j[3] : -1, xa, xb);
        pair(k0 + 2, k0 + 2 < k1, k0 + 3 < k1, te.e0[2], te.e1[2], te.e0[3], te.e1[3]);
    }
    for (int kk = k0 + 4; __any_sync(0xffffffffu, kk < k1); kk += 2) {      // larger in-degrees (quadtree meshes)
        const bool on0 = kk < k1, on1 = kk + 1 < k1;
        pw_load_rows_x2(tile, xin, ld, on0 ? a.nbr[kk] : -1, on1 ? a.nbr[kk + 1] : -1, xa, xb);
        const float ea0 = (on0 && a.ea) ? a.ea[(size_t)kk * 2] : 0.f, ea1 = (on0 && a.ea) ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        const float eb0 = (on1 && a.ea) ? a.ea[(size_t)kk * 2 + 2] : 0.f, eb1 = (on1 && a.ea) ? a.ea[(size_t)kk * 2 + 3] : 0.f;
        pair(kk, on0, on1, ea0, ea1, eb0, eb1);
    }
    PW_MARK(18);
    const float li = (l > 0.f) ? 1.f / l : 0.f;
    if (valid && writer) {
        a.mstat[(size_t)i * a.NC + c] = m;
        a.linv[(size_t)i * a.NC + c] = li;
    }
    // (6) z half (and, from the second half, [ze | zs | 0]) -> A operand
    {
        float zz[PW_H];
#pragma unroll
        for (int k = 0; k < PW_H; ++k) zz[k] = z[k] * li;
        tc_stage_a_at<PW_H>(cx.lane_off, cx.ah_col + h * PW_H, cx.al_col + h * PW_H, zz);
        if (h == 1) {
            float tail[8] = {ze0 * li, ze1 * li, zs * li, 0.f, 0.f, 0.f, 0.f, 0.f};
            tc_stage_a_at<8>(cx.lane_off, cx.ah_col + 32, cx.al_col + 32, tail);
        }
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    PW_MARK(19);
    if (t == 0) {
        tc::fence_after_sync();
        tc_mma3_at(0, cx.p_base + st.pcol, cx.ah_col, cx.al_col, tc::smem_u32(wb + L.W2H), tc::smem_u32(wb + L.W2L), FC, L.K2, true);
        tc::commit(cx.bar);
    }
    PW_MARK(20);
    cx.pending = true;
    cx.wpar ^= 1u << buf;
    cx.toggle ^= 1;
}

// LayerNorm over the node's 32 channels, 16 per thread (two-pass statistics, partial sums meet through shared memory)
__device__ __forceinline__ void pw_layer_norm(const PwCtx& pw, float (&x)[PW_H], float eps, const float* __restrict__ g,
                                              const float* __restrict__ b) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PW_H; ++k) s += x[k];
    const float mean = pw_pair_sum(pw, s, 0) * (1.f / FC);
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < PW_H; ++k) {
        const float d = x[k] - mean;
        v = fmaf(d, d, v);
    }
    const float rstd = rsqrtf(pw_pair_sum(pw, v, 1) * (1.f / FC) + eps);
#pragma unroll
    for (int k = 0; k < PW_H; ++k) x[k] = (x[k] - mean) * rstd * g[pw.h * PW_H + k] + b[pw.h * PW_H + k];
}

// gate epilogue of slot s on this thread's 16 channels (math: fused_fwd.inl gate_epilogue)
__device__ __forceinline__ void gate_epilogue_pw(TcCtx& cx, const PwCtx& pw, const FusedFwdArgs& a, int row0, int i, bool valid,
                                                 int s, const float* __restrict__ prm, float (&P)[PW_H]) {
    const int off = pw.h * PW_H;
    const uint32_t stash = cx.stash_col + off;
    float* tile = pw.htile;
    if (s <= 2) {
        float cp[PW_H];
        if (a.Cprev) pw_load_rows(tile, a.Cprev + off, FC, valid ? i : -1, cp);
        else {
#pragma unroll
            for (int o = 0; o < PW_H; ++o) cp[o] = 0.f;
        }
        if (s < 2) {                 // I, F
            const float* wc = prm + (s == 0 ? 0 : 1) * FC + off;
            const float* bb = prm + (s == 0 ? 3 : 4) * FC + off;
#pragma unroll
            for (int o = 0; o < PW_H; ++o) P[o] = sigm(P[o] + wc[o] * cp[o] + bb[o]);
            pw_store_rows(tile, a.gates + s * FC + off, 4 * FC, row0, a.N, P);
            tc_store_cols<2>(cx.lane_off, stash + (uint32_t)s * FC, P);
            return;
        }
        float I[PW_H], Fg[PW_H];
        tc_load_cols<2>(cx.lane_off, stash, I);
        tc_load_cols<2>(cx.lane_off, stash + FC, Fg);
#pragma unroll
        for (int o = 0; o < PW_H; ++o) P[o] = ftanh(P[o] + prm[5 * FC + off + o]);
        pw_store_rows(tile, a.gates + 2 * FC + off, 4 * FC, row0, a.N, P);
#pragma unroll
        for (int o = 0; o < PW_H; ++o) P[o] = fmaf(Fg[o], cp[o], I[o] * P[o]);
        pw_store_rows(tile, a.Craw + off, FC, row0, a.N, P);
        tc_store_cols<2>(cx.lane_off, stash + 2 * FC, P);
        return;
    }
    float Cn[PW_H];
    tc_load_cols<2>(cx.lane_off, stash + 2 * FC, Cn);
#pragma unroll
    for (int o = 0; o < PW_H; ++o) P[o] = sigm(P[o] + prm[2 * FC + off + o] * Cn[o] + prm[6 * FC + off + o]);   // O
    pw_store_rows(tile, a.gates + 3 * FC + off, 4 * FC, row0, a.N, P);
    if (a.Oout) pw_store_rows(tile, a.Oout + off, FC, row0, a.N, P);
    {
        float Hh[PW_H];
#pragma unroll
        for (int o = 0; o < PW_H; ++o) Hh[o] = P[o] * ftanh(Cn[o]);
        if (a.norm_h) pw_layer_norm(pw, Hh, a.eps, prm + 7 * FC, prm + 8 * FC);
        pw_store_rows(tile, a.Hout + off, FC, row0, a.N, Hh);
    }
    if (a.norm_c) pw_layer_norm(pw, Cn, a.eps, prm + 9 * FC, prm + 10 * FC);
    pw_store_rows(tile, a.Cout + off, FC, row0, a.N, Cn);
    if (a.head_in) {
        if (a.norm_o) pw_layer_norm(pw, P, a.eps, prm + 11 * FC, prm + 12 * FC);
#pragma unroll
        for (int o = 0; o < PW_H; ++o) P[o] = fmaxf(P[o], 0.f);
        if (a.ldh % 4 == 0) pw_store_rows(tile, a.head_in + off, a.ldh, row0, a.N, P);
        if (valid) {
            float* hr = a.head_in + (size_t)i * a.ldh;
            if (a.ldh % 4 != 0) {
#pragma unroll
                for (int o = 0; o < PW_H; ++o) hr[off + o] = P[o];
            }
            if (pw.h == 1) {
                if (a.concat) hr[FC] = a.concat[i];
                for (int k = FC + 1; k < a.ldh; ++k) hr[k] = 0.f;
            }
        }
    }
}

template <int DAC>
__global__ void __launch_bounds__(256, 2) fused_fwd_pw_kernel(const __grid_constant__ FusedFwdArgs a) {
    qmp_seed_init(a.seed, a.salt);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[3];
    __shared__ uint32_t tmem_slot;
    constexpr int DA_ = DAC > 0 ? DAC : 4;
    constexpr TcFwdLayout LA(DA_), LB(32);
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
    if (warp == 0) tc::tmem_alloc(&tmem_slot, TC_COLS);    // U 48 | A_hi 40 | A_lo 40 | P 32 | stash 96 (I, F, C')
    if (a.mode == 1)
        for (int idx = t; idx < 13 * FC; idx += 256) prm[idx] = a.params[idx];
    pdl_wait();
    pdl_launch();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    PwCtx pw;
    pw.h = warp >> 2;
    pw.q = warp & 3;
    pw.ex = prm + 13 * FC;
    pw.htile = pw.ex + 2 * 256 + warp * 2 * PW_TILE;
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
    cx.lane_off = (uint32_t)(pw.q * 32) << 16;
    cx.lane_base = cx.tmem + cx.lane_off;
    cx.u_col = cx.tmem + 0;
    cx.ah_col = cx.tmem + 48;
    cx.al_col = cx.tmem + 88;
    cx.p_base = cx.tmem + 128;
    cx.stash_col = cx.tmem + 160;
    cx.rtile = pw.htile;
    cx.rtile_g = pw.htile;

    const int ntiles = (a.N + 127) / 128;
    const int nsteps = (a.mode == 1) ? 4 * ((a.GA ? 1 : 0) + (a.GB == 8 ? 2 : 1)) : a.NC;
    TcStep st, nx;
    if ((int)blockIdx.x < ntiles) {
        tc_fwd_step<DA_, 32>(a, 0, st);
        if (t == 0) tc_prefetch_image(cx, 0, st.img, st.bytes);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * 128 + pw.q * 32 + (t & 31);
        const bool valid = i < a.N;
        const int row0 = tile * 128 + pw.q * 32;
        TcEdges te;
        tc_load_edges(te, a.ptr, a.nbr, a.ea, i, valid);
        TcEdges none = te;
        none.k0 = none.k1 = 0;
        for (int k = 0; k < nsteps; ++k) {
            tc_fwd_step<DA_, 32>(a, k, st);
            const bool has_next = (k + 1 < nsteps) || (tile + (int)gridDim.x < ntiles);
            tc_fwd_step<DA_, 32>(a, (k + 1 < nsteps) ? k + 1 : 0, nx);
            bool ranA = false;
            if constexpr (DAC > 0) {
                if (st.segA) {       // narrow conv: one thread per node (first warp of the pair), the other idles along
                    const bool act = pw.h == 0;
                    PW_MARK(1);
                    conv_fwd_tc<DA_>(cx, a, i, valid && act, act ? te : none, st, nx, has_next, act);
                    PW_MARK(2);
                    ranA = true;
                }
            }
            if (!ranA) conv_fwd_pw(cx, pw, a, i, valid, te, st, nx, has_next);
            if (!st.last) continue;
            PW_MARK(30);
            if (cx.pending) tc_wait(cx);
            PW_MARK(31);
            float P[PW_H];
            tc_load_cols<2>(cx.lane_off, cx.p_base + st.pcol + pw.h * PW_H, P);
            const int off = pw.h * PW_H;
            auto add_bias = [&](const uint8_t* img, int b3_off) {
                const float4* b = reinterpret_cast<const float4*>(img + b3_off) + off / 4;
#pragma unroll
                for (int o = 0; o < PW_H; o += 4) {
                    const float4 v = __ldg(b + o / 4);
                    P[o] += v.x; P[o + 1] += v.y; P[o + 2] += v.z; P[o + 3] += v.w;
                }
            };
            if (a.mode == 1) {
                const int s = st.slot;
                if constexpr (DAC > 0) {
                    if (a.GA) add_bias(reinterpret_cast<const uint8_t*>(a.wa) + (size_t)s * LA.BYTES, LA.B3);
                }
                add_bias(reinterpret_cast<const uint8_t*>(a.wb) + (size_t)s * LB.BYTES, LB.B3);
                if (a.GB == 8) add_bias(reinterpret_cast<const uint8_t*>(a.wb) + (size_t)(4 + s) * LB.BYTES, LB.B3);
                gate_epilogue_pw(cx, pw, a, row0, i, valid, s, prm, P);
                PW_MARK(32);
            } else {
                add_bias(st.img, st.segA ? LA.B3 : LB.B3);
                if (a.relu_out) {
#pragma unroll
                    for (int o = 0; o < PW_H; ++o) P[o] = fmaxf(P[o], 0.f);
                }
                if (a.C == FC && (a.ldo % 4) == 0) {
                    pw_store_rows(pw.htile, a.out + (size_t)k * a.C + off, a.ldo, row0, a.N, P);
                } else if (valid) {
                    float* orow = a.out + (size_t)i * a.ldo + (size_t)k * a.C;
#pragma unroll
                    for (int o = 0; o < PW_H; ++o)
                        if (off + o < a.C) orow[off + o] = P[o];
                }
            }
        }
    }
    if (cx.pending) tc_wait(cx);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(cx.tmem, TC_COLS);
}

template <int DAC>
int launch_fwd_pw(const FusedFwdArgs& a, cudaStream_t st) {
    constexpr int DA_ = DAC > 0 ? DAC : 4;
    constexpr TcFwdLayout LA(DA_), LB(32);
    constexpr int SLOT = (LA.BYTES > LB.BYTES ? LA.BYTES : LB.BYTES);
    const size_t smem = 2 * (size_t)SLOT + (13 * FC + 2 * 256 + 8 * 2 * PW_TILE) * sizeof(float);
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    auto kern = fused_fwd_pw_kernel<DAC>;
    QMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = cdiv(a.N, 128);
    const int grid = ntiles < 2 * n_sm ? ntiles : 2 * n_sm;
    QMP_CUDA(launch_pdl(kern, dim3(grid), dim3(256), smem, st, a));
    QMP_LAUNCH_CHECK("fused_fwd_pw_kernel");
    return 0;
}

}  // namespace qmp
