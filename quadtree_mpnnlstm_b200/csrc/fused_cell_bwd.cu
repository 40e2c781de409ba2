// Decoder-cell backward: the gradient of one GConvLSTM step (4 TransformerConvs on the 4-wide X, 4 on the 32-wide H, gate
// pre-activation gradients dP from qmp_lstm_gates_bwd) with respect to X and H, plus the per-node rows Zs / dUs that the
// weight-gradient kernel reduces -- ONE persistent launch in the mapping of fused_cell_fwd.cu, replacing the target-side
// and the source-side kernel of fused_bwd_tc.inl (two launches, eight narrow MMA chains each, one thread per node).
//
//   dense contractions (tcgen05, 3xTF32, A in tensor memory, one issuing warp):
//       G1  dz_g = g_g W2cat_g   (N = 48: [dz_h 32 | dze0 dze1 dzs 0 | 0 | dz_x 4 | dze0 dze1 dzs 0])   per gate g
//           dx  += g_g W3cat_g   (N = 48: [dH 32 | dX 4 | 0])                                            accumulated
//       G2  dH  += [du_c | dw_c] W1_c                                                                    per H conv c
//   edge phase (octet layout, 8 lanes per node row): softmax recomputed from the saved logits, d alpha, ds, du = sum ds h_j,
//       z = sum alpha h_j (for the weight gradients) -- and the SOURCE side of every edge in the same pass: the contribution
//       ds_e u_i + alpha_e dz_i of edge j -> i to dH_j is formed from target-side quantities only (u_i was saved by the
//       forward kernel, dz_i is in shared memory) and added to row j with one 16-byte vector reduction per lane
//       (red.global.add.v4.f32).  No out-CSR pass, no ds round trip through memory, no second launch.
//   X convs (4-wide rows): one thread per (node, conv), plain FFMA, same scheme with 16-byte reductions onto dX_j.
// dxa / dxb are zeroed by the entry point and receive every term through reductions (self terms included), so no CTA
// ordering is assumed.  Reference: autograd of GConvLSTM.forward (model/model.py:394-463) around PyG TransformerConv.
#include "fused_fwd.inl"
#include "fused_cell.cuh"
#include "lstm_oct.cuh"

namespace qmp {

struct CellBwdLayout {
    static constexpr int B1G = 2 * 48 * 32 * 4;           // per gate: W2cat^T, N = 48 rows x K = 32, hi then lo
    static constexpr int B3G = 2 * 48 * 32 * 4;           // per gate: W3cat^T
    static constexpr int B2G = 2 * 32 * 40 * 4;           // per H conv: W1^T, N = 32 rows x K = 40
    static constexpr int B1H = 0, B3H = 4 * B1G, B2H = B3H + 4 * B3G;
    static constexpr int MMA_BYTES = B2H + 4 * B2G;
    static constexpr int W1X = MMA_BYTES;                 // [4][6][4] logit weights of the X convs
    static constexpr int B1X = W1X + 4 * 24 * 4;          // [4][8]
    static constexpr int BYTES = B1X + 4 * 8 * 4;
};
static_assert(CellBwdLayout::BYTES % 16 == 0, "bulk copies move 16-byte units");

constexpr uint32_t TB_A = 0;          // g rows of gate g at TB_A + 64 g (hi 32 | lo 32); later [du|dw] of conv c at 80 c (hi 40 | lo 40)
constexpr uint32_t TB_D1 = 256;       // dz of gate g at TB_D1 + 48 g
constexpr uint32_t TB_D2 = 448;       // dx accumulators: dH 32 | dX 4 | pad
constexpr size_t CELLB_SMEM = CellBwdLayout::BYTES + 4 * XPLANE * sizeof(float);

struct CellBwdArgs {
    int N;
    const int* ptr; const int* nbr; const float* ea;
    const float* xa; int lda; const float* xb; int ldb;
    const float* usave;                                    // [N, 128] logit projections of the H convs (forward kernel)
    float* dP; int lddp;                                   // [N, 128] gate pre-activation gradients: input, or -- with `gates` --
                                                           // OUTPUT of the fused gate backward (the weight-gradient kernel reads it)
    // gate backward fused into the prologue of every tile (north_star (c); lstm_oct.cuh).  gates == nullptr: dP is an input.
    const float* gates; const float* Craw; const float* Cprev; const float* prm;      // [N, 128], [N, 32], [N, 32] or null, [13, 32]
    const float* dHout; const float* dCout; const float* dOdirect; const float* dHead; int lddh;   // each [N, 32] or null; [N, lddh]
    float* dCprev; float* dparams;                         // [N, 32] or null; [13, 32] accumulated with atomics, or null
    int norm_h, norm_c, norm_o; float eps;
    const float* logit; const float* mstat; const float* linv;     // [E, 8], [N, 8], [N, 8]
    float* zB; float* duB;                                 // [N, 128]: z / du of H conv c in columns 32 c .. 32 c + 31
    float* sd;                                             // [N, 64]: x(4) | 1 0 0 0 | (ze0 ze1 zs 0) x 4 H convs | 0 (8) | (z(4) | ze0 ze1 zs 0) x 4 X convs
    float* sg;                                             // [N, 32]: (dw0 dw1) x 4 H convs | du(4) x 4 X convs | (dw0 dw1) x 4 X convs
                                                           // (panel layout of the weight-gradient kernel, cell_wgrad.cu)
    float* dxa; float* dxb;                                // [N, lda], [N, ldb]: zero on entry
    float drop_p; unsigned long long seed; const unsigned long long* salt;
};

// 16 warps, no dedicated MMA warp: both MMA groups of a tile are issued at points where every warp waits for their result
// anyway, and 512 threads keep the 128-register budget (the 17-warp CTA of the forward kernel is capped at 96)
constexpr int CELLB_THREADS = CELL_WORKERS;
__device__ __forceinline__ void cellb_sync() { asm volatile("bar.sync 0, %0;" ::"n"(CELLB_THREADS) : "memory"); }

__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// everything the X conv of one (node, conv) thread reads from global memory for its first four in-edges: fetched while the
// first MMA group is in flight, so that only arithmetic is left once dz arrives
struct CellbXPre {
    int k0, k1, jn[4];
    float4 xi, xj[4];
    float2 ev[4];
    float lg[4], m, li;
};
__device__ __forceinline__ void cellb_xconv_load(CellbXPre& x, const CellBwdArgs& a, int i, bool valid, int c) {
    x.k0 = valid ? __ldg(a.ptr + i) : 0;
    x.k1 = valid ? __ldg(a.ptr + i + 1) : 0;
    x.xi = make_float4(0.f, 0.f, 0.f, 0.f);
    x.m = x.li = 0.f;
    if (valid) {
        x.xi = __ldg(reinterpret_cast<const float4*>(a.xa + (size_t)i * a.lda));
        x.m = __ldg(a.mstat + (size_t)i * 8 + c);
        x.li = __ldg(a.linv + (size_t)i * 8 + c);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) x.jn[e] = (x.k0 + e < x.k1) ? __ldg(a.nbr + x.k0 + e) : -1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        x.xj[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        x.ev[e] = make_float2(0.f, 0.f);
        x.lg[e] = 0.f;
        if (x.jn[e] >= 0) {
            x.xj[e] = __ldg(reinterpret_cast<const float4*>(a.xa + (size_t)x.jn[e] * a.lda));
            if (a.ea) x.ev[e] = __ldg(reinterpret_cast<const float2*>(a.ea) + x.k0 + e);
            x.lg[e] = __ldg(a.logit + (size_t)(x.k0 + e) * 8 + c);
        }
    }
}

// X conv c of node i (row nrow of the tile): target side, self term of dX into the exchange row, source side by reductions
__device__ __forceinline__ void cellb_xconv(const CellBwdArgs& a, const uint8_t* smem, float* exch, int i, bool valid, int nrow, int c,
                                            const CellbXPre& x) {
    using L = CellBwdLayout;
    float* xr = exch + c * XPLANE + nrow * XS + 36;        // dz_x (4) | dze0 dze1 dzs 0 of this conv and node
    const float4 dz = ld4(xr), dze = ld4(xr + 4);
    float4 self = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        const float* w1x = reinterpret_cast<const float*>(smem + L::W1X) + c * 24;
        const float* b1x = reinterpret_cast<const float*>(smem + L::B1X) + c * 8;
        const float4 xi = x.xi;
        float4 u;                                          // logit projection of this node (rows 0..3 of W1x)
        {
            const float4 w0 = ld4(w1x), w1 = ld4(w1x + 4), w2 = ld4(w1x + 8), w3 = ld4(w1x + 12);
            u.x = fmaf(w0.w, xi.w, fmaf(w0.z, xi.z, fmaf(w0.y, xi.y, fmaf(w0.x, xi.x, b1x[0]))));
            u.y = fmaf(w1.w, xi.w, fmaf(w1.z, xi.z, fmaf(w1.y, xi.y, fmaf(w1.x, xi.x, b1x[1]))));
            u.z = fmaf(w2.w, xi.w, fmaf(w2.z, xi.z, fmaf(w2.y, xi.y, fmaf(w2.x, xi.x, b1x[2]))));
            u.w = fmaf(w3.w, xi.w, fmaf(w3.z, xi.z, fmaf(w3.y, xi.y, fmaf(w3.x, xi.x, b1x[3]))));
        }
        const float m = x.m, li = x.li;
        const int k0 = x.k0, k1 = x.k1;
        auto dalpha = [&](int kk, const float4& xj, const float2& ev, float lg, float& al, float& keep) {
            al = fast_exp(lg - m) * li;
            keep = fdropout_scale(QMP_SEED_SM, (long long)kk * 8 + c, a.drop_p);
            return (fmaf(dz.w, xj.w, fmaf(dz.z, xj.z, fmaf(dz.y, xj.y, dz.x * xj.x))) + fmaf(dze.x, ev.x, fmaf(dze.y, ev.y, dze.z))) * keep;
        };
        auto fetch = [&](int kk, int& j, float4& xj, float2& ev, float& lg) {       // in-edges beyond the fourth (quadtree meshes)
            j = __ldg(a.nbr + kk);
            xj = __ldg(reinterpret_cast<const float4*>(a.xa + (size_t)j * a.lda));
            ev = make_float2(0.f, 0.f);
            if (a.ea) ev = __ldg(reinterpret_cast<const float2*>(a.ea) + kk);
            lg = __ldg(a.logit + (size_t)kk * 8 + c);
        };
        float al4[4], keep4[4], dal4[4], tsum = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            al4[e] = keep4[e] = dal4[e] = 0.f;
            if (x.jn[e] >= 0) {
                dal4[e] = dalpha(k0 + e, x.xj[e], x.ev[e], x.lg[e], al4[e], keep4[e]);
                tsum = fmaf(al4[e], dal4[e], tsum);
            }
        }
        for (int kk = k0 + 4; kk < k1; ++kk) {
            int j; float4 xj; float2 ev; float lg, al, keep;
            fetch(kk, j, xj, ev, lg);
            const float dal = dalpha(kk, xj, ev, lg, al, keep);
            tsum = fmaf(al, dal, tsum);
        }
        float4 du = make_float4(0.f, 0.f, 0.f, 0.f), z = du;
        float dw0 = 0.f, dw1 = 0.f, ze0 = 0.f, ze1 = 0.f, zs = 0.f;
        auto accumulate = [&](int j, const float4& xj, const float2& ev, float al, float keep, float dal) {
            const float dsv = al * (dal - tsum), alk = al * keep;
            fma4(du, dsv, xj);
            fma4(z, alk, xj);
            dw0 = fmaf(dsv, ev.x, dw0); dw1 = fmaf(dsv, ev.y, dw1);
            ze0 = fmaf(alk, ev.x, ze0); ze1 = fmaf(alk, ev.y, ze1); zs += alk;
            // source side of this edge: dX_j += ds u_i + alpha dz_i
            red4(a.dxa + (size_t)j * a.lda, fmaf(dsv, u.x, alk * dz.x), fmaf(dsv, u.y, alk * dz.y), fmaf(dsv, u.z, alk * dz.z),
                 fmaf(dsv, u.w, alk * dz.w));
        };
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (x.jn[e] >= 0) accumulate(x.jn[e], x.xj[e], x.ev[e], al4[e], keep4[e], dal4[e]);
        for (int kk = k0 + 4; kk < k1; ++kk) {
            int j; float4 xj; float2 ev; float lg, al, keep;
            fetch(kk, j, xj, ev, lg);
            const float dal = dalpha(kk, xj, ev, lg, al, keep);
            accumulate(j, xj, ev, al, keep, dal);
        }
        float* sdr = a.sd + (size_t)i * 64;
        st4(sdr + 32 + 8 * c, z.x, z.y, z.z, z.w);
        st4(sdr + 36 + 8 * c, ze0, ze1, zs, 0.f);
        if (c == 0) {                                      // the row's shared part: x | 1 0 0 0 ... 0 (8)
            st4(sdr, xi.x, xi.y, xi.z, xi.w);
            st4(sdr + 4, 1.f, 0.f, 0.f, 0.f);
            st4(sdr + 24, 0.f, 0.f, 0.f, 0.f);
            st4(sdr + 28, 0.f, 0.f, 0.f, 0.f);
        }
        float* sgr = a.sg + (size_t)i * 32;
        st4(sgr + 8 + 4 * c, du.x, du.y, du.z, du.w);
        *reinterpret_cast<float2*>(sgr + 24 + 2 * c) = make_float2(dw0, dw1);
        // self term: dX_i += W1x^T [du | dw]
        const float dU[6] = {du.x, du.y, du.z, du.w, dw0, dw1};
#pragma unroll
        for (int r = 0; r < 6; ++r) fma4(self, dU[r], ld4(w1x + 4 * r));
    }
    st4(xr, self.x, self.y, self.z, self.w);
}

// what one warp pass of the H-conv edge phase reads from global memory before any arithmetic: first edge quad (sources,
// attributes, rows, logits), the node's saved logit projections and softmax statistics
struct CellbHPre {
    int k0, deg, jj0, jx0[4];
    float2 ev0;
    float4 hr0[4];
    float m[2], li[2], lg[2];
};
__device__ __forceinline__ void cellb_h_load(CellbHPre& h, const CellBwdArgs& a, int tile0, int tcount, int ln, int l8, int obase) {
    const int cc = l8 >> 2, e4 = l8 & 3;
    const bool valid = ln < tcount;
    const int i = tile0 + ln;
    h.k0 = valid ? __ldg(a.ptr + i) : 0;
    h.deg = valid ? __ldg(a.ptr + i + 1) - h.k0 : 0;
    const bool on0 = e4 < h.deg;
    h.jj0 = on0 ? __ldg(a.nbr + h.k0 + e4) : -1;
    h.ev0 = make_float2(0.f, 0.f);
    if (on0 && a.ea) h.ev0 = __ldg(reinterpret_cast<const float2*>(a.ea) + h.k0 + e4);
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
        const int crole = 4 + 2 * r2 + cc;
        h.m[r2] = valid ? __ldg(a.mstat + (size_t)i * 8 + crole) : 0.f;
        h.li[r2] = valid ? __ldg(a.linv + (size_t)i * 8 + crole) : 0.f;
        h.lg[r2] = on0 ? __ldg(a.logit + (size_t)(h.k0 + e4) * 8 + crole) : 0.f;
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        h.jx0[x] = __shfl_sync(0xffffffffu, h.jj0, obase + x);
        h.hr0[x] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h.jx0[x] >= 0) h.hr0[x] = __ldg(reinterpret_cast<const float4*>(a.xb + (size_t)h.jx0[x] * a.ldb) + l8);
    }
}

// everything tile [n0, n0 + cnt) streams from global memory, requested into L2 by one thread (rows of h are gathered through
// the CSR and stay L2-resident anyway: 6 MB)
__device__ __forceinline__ void cellb_prefetch_tile(const CellBwdArgs& a, int n0, int cnt) {
    const size_t n = (size_t)n0;
    if (a.gates) {
        tc::l2_prefetch(a.gates + n * 128, (long long)cnt * 512);
        tc::l2_prefetch(a.Craw + n * 32, (long long)cnt * 128);
        if (a.Cprev) tc::l2_prefetch(a.Cprev + n * 32, (long long)cnt * 128);
        if (a.dHout) tc::l2_prefetch(a.dHout + n * 32, (long long)cnt * 128);
        if (a.dCout) tc::l2_prefetch(a.dCout + n * 32, (long long)cnt * 128);
        if (a.dOdirect) tc::l2_prefetch(a.dOdirect + n * 32, (long long)cnt * 128);
        if (a.dHead) tc::l2_prefetch(a.dHead + n * a.lddh, (long long)cnt * a.lddh * 4);
    } else {
        tc::l2_prefetch(a.dP + n * a.lddp, (long long)cnt * a.lddp * 4);
    }
    tc::l2_prefetch(a.usave + n * 128, (long long)cnt * 512);
    tc::l2_prefetch(a.mstat + n * 8, (long long)cnt * 32);
    tc::l2_prefetch(a.linv + n * 8, (long long)cnt * 32);
    tc::l2_prefetch(a.xa + n * a.lda, (long long)cnt * a.lda * 4);
    const int k0 = __ldg(a.ptr + n0), k1 = __ldg(a.ptr + n0 + cnt);
    tc::l2_prefetch(a.logit + (size_t)k0 * 8, (long long)(k1 - k0) * 32);
    tc::l2_prefetch(a.nbr + k0, (long long)(k1 - k0) * 4);
    if (a.ea) tc::l2_prefetch(a.ea + (size_t)k0 * 2, (long long)(k1 - k0) * 8);
}

__global__ void __launch_bounds__(CELLB_THREADS, 1) fused_cell_bwd_kernel(const __grid_constant__ CellBwdArgs a,
                                                                         const uint8_t* __restrict__ img, const int Q, const int R,
                                                                         const int T0) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[2];                 // 0: MMA groups, 1: image landed
    __shared__ uint32_t tmem_slot;
    __shared__ float s_dp[P_COUNT * 32];         // parameter gradients of the fused gate backward, summed over this CTA's nodes
    using L = CellBwdLayout;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    qmp_seed_init(a.seed, a.salt);
    if (t < P_COUNT * 32) s_dp[t] = 0.f;
    float* exch = reinterpret_cast<float*>(smem + L::BYTES);
    if (t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    cellb_sync();
    tc::fence_after_sync();
    if (t == 0) {
        tc::mbar_expect_tx(&bars[1], (uint32_t)L::BYTES);
        for (int off = 0; off < L::BYTES; off += 16384)
            tc::bulk_g2s(smem + off, img + off, (uint32_t)(L::BYTES - off < 16384 ? L::BYTES - off : 16384), &bars[1]);
    }
    pdl_wait();            // barrier init, tensor-memory allocation and the image copy run under the predecessor's tail
    pdl_launch();
    const uint32_t tmem = tmem_slot;
    const int beg = (int)blockIdx.x * Q;
    int end = beg + Q;
    if (end > a.N) end = a.N;
    if (t == 64 && beg < end) cellb_prefetch_tile(a, beg, end - beg < T0 ? end - beg : T0);      // first tile, under the image load
    // the weight image (139 KB per CTA, ~4 us from L2) is awaited where it is first used -- by the thread that issues G1 and by
    // every thread before the X convs -- so that the first tile's gate backward runs under the copy

    {
        const int q = warp & 3, cg = warp >> 2;
        const int nrow = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const int o8 = lane >> 3, l8 = lane & 7, obase = lane & ~7;
        const int cc = l8 >> 2, e4 = l8 & 3;
        uint32_t par = 0;
        for (int r = 0; r < R; ++r) {
            const int tile0 = beg + r * T0;
            if (tile0 >= end) break;
            const int tcount = (end - tile0 < T0) ? end - tile0 : T0;
            if (t == 64 && tile0 + T0 < end) cellb_prefetch_tile(a, tile0 + T0, end - tile0 - T0 < T0 ? end - tile0 - T0 : T0);

            CELL_MARK(1);
            if (a.gates) {
                // ---- gate backward (sigmoid / tanh / peepholes / LayerNorms / head input) in octet layout: dP rows to global
                // memory (for the weight-gradient kernel) and into exchange plane g (for the contraction below)
                float dprm[P_COUNT][4];
#pragma unroll
                for (int p = 0; p < P_COUNT; ++p)
#pragma unroll
                    for (int k = 0; k < 4; ++k) dprm[p][k] = 0.f;
                auto PRM = [&](int p, float (&v)[4]) { f4(v, ldg4(a.prm + p * 32 + 4 * l8)); };
#pragma unroll 1
                for (int p = 0; p < 2; ++p) {
                    if (4 * (warp + 16 * p) >= tcount) continue;           // warp-uniform
                    const int ln = 4 * (warp + 16 * p) + o8;
                    const bool valid = ln < tcount;
                    const int i = tile0 + ln;
                    const size_t r32 = (size_t)i * 32 + 4 * l8;
                    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                    float I[4], F[4], T[4], O[4], Cn[4], cp[4], dH[4], dC[4], dO[4], dhd[4];
                    const float* gs = a.gates + (size_t)i * 128 + 4 * l8;
                    f4(I, valid ? ldg4(gs) : zero);
                    f4(F, valid ? ldg4(gs + 32) : zero);
                    f4(T, valid ? ldg4(gs + 64) : zero);
                    f4(O, valid ? ldg4(gs + 96) : zero);
                    f4(Cn, valid ? ldg4(a.Craw + r32) : zero);
                    f4(cp, (valid && a.Cprev) ? ldg4(a.Cprev + r32) : zero);
                    f4(dH, (valid && a.dHout) ? ldg4(a.dHout + r32) : zero);
                    f4(dC, (valid && a.dCout) ? ldg4(a.dCout + r32) : zero);
                    f4(dO, (valid && a.dOdirect) ? ldg4(a.dOdirect + r32) : zero);
                    f4(dhd, (valid && a.dHead) ? ldg4(a.dHead + (size_t)i * a.lddh + 4 * l8) : zero);
                    float dI[4], dF[4], dT[4], dOp[4], dCp[4];
                    oct_gate_bwd(I, F, T, O, Cn, cp, dH, dC, dO, dhd, a.dHead != nullptr, a.norm_h, a.norm_c, a.norm_o, a.eps, PRM, dprm, dI,
                                 dF, dT, dOp, dCp);
                    if (valid) {
                        float* dpr = a.dP + (size_t)i * a.lddp + 4 * l8;
                        st4(dpr, dI[0], dI[1], dI[2], dI[3]);
                        st4(dpr + 32, dF[0], dF[1], dF[2], dF[3]);
                        st4(dpr + 64, dT[0], dT[1], dT[2], dT[3]);
                        st4(dpr + 96, dOp[0], dOp[1], dOp[2], dOp[3]);
                        if (a.dCprev) st4(a.dCprev + r32, dCp[0], dCp[1], dCp[2], dCp[3]);
                    }
                    float* xr = exch + ln * XS + 4 * l8;
                    st4(xr, dI[0], dI[1], dI[2], dI[3]);
                    st4(xr + XPLANE, dF[0], dF[1], dF[2], dF[3]);
                    st4(xr + 2 * XPLANE, dT[0], dT[1], dT[2], dT[3]);
                    st4(xr + 3 * XPLANE, dOp[0], dOp[1], dOp[2], dOp[3]);
                }
                if (a.dparams) {
                    // sum over the warp's 4 octets with a value-halving butterfly (39 shuffles for the 52 values instead of 104):
                    // afterwards lane (b4, b3, l8) holds row i, channel 4 l8 + 2 b3 + b4 of every parameter row i -- 32 distinct
                    // consecutive addresses per warp instruction (float atomics on shared memory are compare-and-swap loops)
                    const bool b4 = lane & 16, b3 = lane & 8;
                    float r1[2 * P_COUNT];
#pragma unroll
                    for (int i = 0; i < 2 * P_COUNT; ++i) {
                        const float e = dprm[i >> 1][2 * (i & 1)], o = dprm[i >> 1][2 * (i & 1) + 1];
                        r1[i] = (b4 ? o : e) + __shfl_xor_sync(0xffffffffu, b4 ? e : o, 16);
                    }
#pragma unroll
                    for (int i = 0; i < P_COUNT; ++i) {
                        const float e = r1[2 * i], o = r1[2 * i + 1];
                        const float v = (b3 ? o : e) + __shfl_xor_sync(0xffffffffu, b3 ? e : o, 8);
                        atomicAdd(&s_dp[i * 32 + 4 * l8 + (b3 ? 2 : 0) + (b4 ? 1 : 0)], v);
                    }
                }
                cellb_sync();
                const bool valid = nrow < tcount;
                const float* row = exch + cg * XPLANE + nrow * XS;
                const uint32_t base = lane_addr + TB_A + 64 * (uint32_t)cg;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float v[8];
                    ld8(v, row + 8 * c8);
                    if (!valid) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = 0.f;
                    }
                    cell_stage8(base + 8 * c8, base + 32 + 8 * c8, v);
                }
                tc::tmem_st_wait();
            } else {
                const int i = tile0 + nrow;
                const bool valid = nrow < tcount;
                const float* gp = a.dP + (size_t)i * a.lddp + 32 * cg;
                const uint32_t base = lane_addr + TB_A + 64 * (uint32_t)cg;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float v[8];
                    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
                    if (valid) {
                        p0 = __ldg(reinterpret_cast<const float4*>(gp) + 2 * c8);
                        p1 = __ldg(reinterpret_cast<const float4*>(gp) + 2 * c8 + 1);
                    }
                    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
                    cell_stage8(base + 8 * c8, base + 32 + 8 * c8, v);
                }
                tc::tmem_st_wait();
            }
            tc::fence_before_sync();
            CELL_MARK(2);
            cellb_sync();
            if (t == 0) {                                      // G1: dz of the four gates, skip-path part of dx
                tc::mbar_wait(&bars[1], 0);                    // weights in shared memory (returns at once after the first tile)
                tc::fence_after_sync();
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    const uint32_t ah = tmem + TB_A + 64 * g, al = ah + 32;
                    tc_mma3_at(0, tmem + TB_D1 + 48 * g, ah, al, tc::smem_u32(smem + L::B1H + g * L::B1G),
                               tc::smem_u32(smem + L::B1H + g * L::B1G + L::B1G / 2), 48, 32, false);
                    tc_mma3_at(0, tmem + TB_D2, ah, al, tc::smem_u32(smem + L::B3H + g * L::B3G),
                               tc::smem_u32(smem + L::B3H + g * L::B3G + L::B3G / 2), 48, 32, g > 0);
                }
                tc::commit(&bars[0]);
            }
            __syncwarp();
            CELL_MARK(3);
            // while G1 runs: every global read of the first edge-phase pass
            CellbHPre pre;
            cellb_h_load(pre, a, tile0, tcount, 4 * warp + o8, l8, obase);
            tc::mbar_wait(&bars[0], par);
            par ^= 1;
            tc::fence_after_sync();

            CELL_MARK(4);
            // ---- dz block of gate cg, row nrow: tensor memory -> exchange plane cg (columns 0..35 | 36..43 = the X conv's)
            {
                uint32_t rr[6][8];
#pragma unroll
                for (int c8 = 0; c8 < 6; ++c8) tc::tmem_ld8_nowait(lane_addr + TB_D1 + (uint32_t)(48 * cg + 8 * c8), rr[c8]);
                tc::tmem_ld_wait();
                float* row = exch + cg * XPLANE + nrow * XS;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    st4(row + 8 * c8, __uint_as_float(rr[c8][0]), __uint_as_float(rr[c8][1]), __uint_as_float(rr[c8][2]), __uint_as_float(rr[c8][3]));
                    st4(row + 8 * c8 + 4, __uint_as_float(rr[c8][4]), __uint_as_float(rr[c8][5]), __uint_as_float(rr[c8][6]), __uint_as_float(rr[c8][7]));
                }
                st4(row + 32, __uint_as_float(rr[4][0]), __uint_as_float(rr[4][1]), __uint_as_float(rr[4][2]), 0.f);
                st4(row + 36, __uint_as_float(rr[5][0]), __uint_as_float(rr[5][1]), __uint_as_float(rr[5][2]), __uint_as_float(rr[5][3]));
                st4(row + 40, __uint_as_float(rr[5][4]), __uint_as_float(rr[5][5]), __uint_as_float(rr[5][6]), 0.f);
            }
            CELL_MARK(5);
            cellb_sync();
            CELL_MARK(7);

            // ---- edge phase of the four H convs, octet layout, two conv pairs per pass (pass 0's reads were issued under G1)
#pragma unroll 1
            for (int p = 0; p < 2; ++p) {
                if (4 * (warp + 16 * p) >= tcount) continue;   // warp-uniform
                if (p == 1) cellb_h_load(pre, a, tile0, tcount, 4 * (warp + 16) + o8, l8, obase);
                const int ln = 4 * (warp + 16 * p) + o8;
                const bool valid = ln < tcount;
                const int i = tile0 + ln;
                const int k0 = pre.k0, deg = pre.deg;
                const int nq = __reduce_max_sync(0xffffffffu, (deg + 3) >> 2);      // edge quads of the largest in-degree of the pass
                float* xrow = exch + ln * XS;
                const bool on0 = e4 < deg;
                float4 con0[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) con0[x] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2) {               // conv pair: H convs 2 r2 and 2 r2 + 1; this lane's role conv = 2 r2 + cc
                    float4 dz[2], uu[2];
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        dz[c2] = ld4(xrow + (2 * r2 + c2) * XPLANE + 4 * l8);
                        // the saved logit projections u_i of this conv pair: read here (the tile's rows were prefetched into L2),
                        // not carried in registers from the top of the tile
                        uu[c2] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid) uu[c2] = __ldg(reinterpret_cast<const float4*>(a.usave + (size_t)i * 128 + 32 * (2 * r2 + c2)) + l8);
                    }
                    const int crole = 4 + 2 * r2 + cc;                                         // conv index in logit / mstat / linv
                    const float4 dzt = ld4(xrow + (2 * r2 + cc) * XPLANE + 32);                 // dze0 dze1 dzs of the role conv
                    const float m = pre.m[r2], li = pre.li[r2];
                    // (alpha, d alpha) of this lane's (conv, edge) for the quad held in hr
                    auto coef = [&](const float4 (&hr)[4], bool on, int kk, float lg, const float2& ev, float& al, float& keep) {
                        float v[8];
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
                            for (int x = 0; x < 4; ++x) v[4 * c2 + x] = dot4(dz[c2], hr[x]);
                        const float tot = octet_reduce8(v, l8);
                        al = 0.f;
                        keep = 0.f;
                        if (on) {
                            al = fast_exp(lg - m) * li;
                            keep = fdropout_scale(QMP_SEED_SM, (long long)kk * 8 + crole, a.drop_p);
                        }
                        return (tot + fmaf(dzt.x, ev.x, fmaf(dzt.y, ev.y, dzt.z))) * keep;
                    };
                    auto gather = [&](int qd, bool& on, int& kk, float& lg, float2& ev, float4 (&hr)[4], int (&jx)[4]) {
                        on = 4 * qd + e4 < deg;
                        kk = k0 + 4 * qd + e4;
                        const int jj = on ? __ldg(a.nbr + kk) : -1;
                        ev = make_float2(0.f, 0.f);
                        if (on && a.ea) ev = __ldg(reinterpret_cast<const float2*>(a.ea) + kk);
                        lg = on ? __ldg(a.logit + (size_t)kk * 8 + crole) : 0.f;
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            jx[x] = __shfl_sync(0xffffffffu, jj, obase + x);
                            hr[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (jx[x] >= 0) hr[x] = __ldg(reinterpret_cast<const float4*>(a.xb + (size_t)jx[x] * a.ldb) + l8);
                        }
                    };
                    // pass 1 over the quads: t = sum_e alpha_e d alpha_e of the role conv
                    float al0, keep0;
                    const float dal0 = coef(pre.hr0, on0, k0 + e4, pre.lg[r2], pre.ev0, al0, keep0);
                    float tsum = quad_sum(al0 * dal0);
                    for (int qd = 1; qd < nq; ++qd) {
                        bool on; int kk; float lg; float2 ev; float4 hr[4]; int jx[4];
                        gather(qd, on, kk, lg, ev, hr, jx);
                        float al, keep;
                        const float dal = coef(hr, on, kk, lg, ev, al, keep);
                        tsum += quad_sum(al * dal);
                    }
                    // pass 2: ds, du, z and the source-side contributions
                    float4 du[2], z[2];
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) du[c2] = z[c2] = make_float4(0.f, 0.f, 0.f, 0.f);
                    float dw0 = 0.f, dw1 = 0.f, ze0 = 0.f, ze1 = 0.f, zs = 0.f;
                    auto accumulate = [&](const float4 (&hr)[4], float al, float keep, float dal, const float2& ev, float4 (&con)[4]) {
                        const float dsv = al * (dal - tsum), alk = al * keep;
                        dw0 += quad_sum(dsv * ev.x); dw1 += quad_sum(dsv * ev.y);
                        ze0 += quad_sum(alk * ev.x); ze1 += quad_sum(alk * ev.y); zs += quad_sum(alk);
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                const float dsb = __shfl_sync(0xffffffffu, dsv, obase + 4 * c2 + x);
                                const float alb = __shfl_sync(0xffffffffu, alk, obase + 4 * c2 + x);
                                fma4(du[c2], dsb, hr[x]);
                                fma4(z[c2], alb, hr[x]);
                                fma4(con[x], dsb, uu[c2]);                 // source side: dH_j += ds u_i + alpha dz_i
                                fma4(con[x], alb, dz[c2]);
                            }
                    };
                    accumulate(pre.hr0, al0, keep0, dal0, pre.ev0, con0);
                    for (int qd = 1; qd < nq; ++qd) {
                        bool on; int kk; float lg; float2 ev; float4 hr[4], con[4]; int jx[4];
                        gather(qd, on, kk, lg, ev, hr, jx);
#pragma unroll
                        for (int x = 0; x < 4; ++x) con[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                        float al, keep;
                        const float dal = coef(hr, on, kk, lg, ev, al, keep);
                        accumulate(hr, al, keep, dal, ev, con);
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                            if (jx[x] >= 0) red4(a.dxb + (size_t)jx[x] * a.ldb + 4 * l8, con[x].x, con[x].y, con[x].z, con[x].w);
                    }
                    // rows for the weight-gradient kernel, and [du | dw] for the second contraction
                    __syncwarp();                              // every lane of the octet has read dz of this pair
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int c = 2 * r2 + c2;
                        st4(xrow + c * XPLANE + 4 * l8, du[c2].x, du[c2].y, du[c2].z, du[c2].w);
                        if (valid) {
                            st4(a.zB + (size_t)i * 128 + 32 * c + 4 * l8, z[c2].x, z[c2].y, z[c2].z, z[c2].w);
                            st4(a.duB + (size_t)i * 128 + 32 * c + 4 * l8, du[c2].x, du[c2].y, du[c2].z, du[c2].w);
                        }
                    }
                    if (e4 == 0) {
                        const int c = 2 * r2 + cc;
                        st4(xrow + c * XPLANE + 32, dw0, dw1, 0.f, 0.f);
                        if (valid) {
                            st4(a.sd + (size_t)i * 64 + 8 + 4 * c, ze0, ze1, zs, 0.f);
                            *reinterpret_cast<float2*>(a.sg + (size_t)i * 32 + 2 * c) = make_float2(dw0, dw1);
                        }
                    }
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)                    // first quad's source rows: both conv pairs summed
                    if (pre.jx0[x] >= 0) red4(a.dxb + (size_t)pre.jx0[x] * a.ldb + 4 * l8, con0[x].x, con0[x].y, con0[x].z, con0[x].w);
            }
            CELL_MARK(8);
            cellb_sync();
            CELL_MARK(9);
            CellbXPre xp;                                      // the X conv's global reads fly under the staging and the G2 issue
            cellb_xconv_load(xp, a, tile0 + nrow, nrow < tcount, cg);

            // ---- [du | dw] of H conv cg, row nrow -> tensor memory (K = 40), second contraction
            {
                const float* row = exch + cg * XPLANE + nrow * XS;
                const uint32_t base = lane_addr + TB_A + 80 * (uint32_t)cg;
                float v[8];
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    ld8(v, row + 8 * c8);
                    cell_stage8(base + 8 * c8, base + 40 + 8 * c8, v);
                }
                const float4 dt = ld4(row + 32);
                v[0] = dt.x; v[1] = dt.y; v[2] = v[3] = v[4] = v[5] = v[6] = v[7] = 0.f;
                cell_stage8(base + 32, base + 72, v);
                tc::tmem_st_wait();
            }
            tc::fence_before_sync();
            CELL_MARK(10);
            cellb_sync();
            if (t == 0) {                                      // G2: dH += [du | dw] W1 of the four H convs
                tc::fence_after_sync();
#pragma unroll 1
                for (int c = 0; c < 4; ++c)
                    tc_mma3_at(0, tmem + TB_D2, tmem + TB_A + 80 * c, tmem + TB_A + 80 * c + 40, tc::smem_u32(smem + L::B2H + c * L::B2G),
                               tc::smem_u32(smem + L::B2H + c * L::B2G + L::B2G / 2), 32, 40, true);
                tc::commit(&bars[0]);
            }
            __syncwarp();
            // ---- while G2 runs: X conv cg of node nrow (reads dz_x in columns 36..43 of its own exchange row, leaves its dX term there)
            tc::mbar_wait(&bars[1], 0);
            cellb_xconv(a, smem, exch, tile0 + nrow, nrow < tcount, nrow, cg, xp);
            CELL_MARK(6);
            tc::mbar_wait(&bars[0], par);
            par ^= 1;
            tc::fence_after_sync();
            cellb_sync();                                      // every X conv's dX term is in the exchange rows
            CELL_MARK(11);
            // ---- self terms: dH_i (8 columns per thread) and, for cg == 0, dX_i = W3x^T g + sum_c W1x_c^T [du | dw]
            {
                const int i = tile0 + nrow;
                const bool valid = nrow < tcount;
                float v[8];
                tc::tmem_ld8(lane_addr + TB_D2 + 8 * (uint32_t)cg, v);
                float w[8];
                if (cg == 0) tc::tmem_ld8(lane_addr + TB_D2 + 32, w);
                if (valid) {
                    float* d = a.dxb + (size_t)i * a.ldb + 8 * cg;
                    red4(d, v[0], v[1], v[2], v[3]);
                    red4(d + 4, v[4], v[5], v[6], v[7]);
                    if (cg == 0) {
                        float4 s = make_float4(w[0], w[1], w[2], w[3]);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 pz = ld4(exch + c * XPLANE + nrow * XS + 36);
                            s.x += pz.x; s.y += pz.y; s.z += pz.z; s.w += pz.w;
                        }
                        red4(a.dxa + (size_t)i * a.lda, s.x, s.y, s.z, s.w);
                    }
                }
            }
            tc::fence_before_sync();
            CELL_MARK(12);
            cellb_sync();                                       // exchange planes and tensor memory free for the next tile
            CELL_MARK(13);
        }
    }
    tc::fence_before_sync();
    cellb_sync();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
    if (a.gates && a.dparams && t < P_COUNT * 32) {
        const float v = s_dp[t];
        if (v != 0.f) atomicAdd(a.dparams + t, v);
    }
}

// ---- weight image ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cellb_put(uint8_t* img, int off, int half, int n, int k, int K, float v) {
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    *reinterpret_cast<float*>(img + off + img_off(n, k, K)) = hi;
    *reinterpret_cast<float*>(img + off + half + img_off(n, k, K)) = lo;
}

__global__ void __launch_bounds__(256) fused_pack_cell_bwd_kernel(const float* __restrict__ packA, const float* __restrict__ packB,
                                                                  uint8_t* __restrict__ img) {
    using L = CellBwdLayout;
    using SA = ConvSizes<4>;
    using SB = ConvSizes<32>;
    const int tid = blockIdx.x * 256 + threadIdx.x, nth = gridDim.x * 256;
    auto A = [&](int g) { return packA + (size_t)g * SA::TOTAL; };
    auto B = [&](int g) { return packB + (size_t)g * SB::TOTAL; };
    for (int idx = tid; idx < 4 * 48 * 32; idx += nth) {              // W2cat_g^T: row n = z index, column k = gate output
        const int g = idx / (48 * 32), n = (idx / 32) % 48, k = idx % 32;
        float v = 0.f;
        if (n < 36) v = (B(g) + SB::W1 + SB::B1)[k * 36 + n];
        else if (n >= 40) v = (A(g) + SA::W1 + SA::B1)[k * 8 + (n - 40)];
        cellb_put(img, L::B1H + g * L::B1G, L::B1G / 2, n, k, 32, v);
    }
    for (int idx = tid; idx < 4 * 48 * 32; idx += nth) {              // W3cat_g^T: row n = input index (h 32 | x 4), column k = gate output
        const int g = idx / (48 * 32), n = (idx / 32) % 48, k = idx % 32;
        float v = 0.f;
        if (n < 32) v = (B(g) + SB::W1 + SB::B1 + SB::W2)[k * 32 + n];
        else if (n < 36) v = (A(g) + SA::W1 + SA::B1 + SA::W2)[k * 4 + (n - 32)];
        cellb_put(img, L::B3H + g * L::B3G, L::B3G / 2, n, k, 32, v);
    }
    for (int idx = tid; idx < 4 * 32 * 40; idx += nth) {              // W1_c^T: row n = h index, column r = [du | dw] index
        const int c = idx / (32 * 40), n = (idx / 40) % 32, r = idx % 40;
        cellb_put(img, L::B2H + c * L::B2G, L::B2G / 2, n, r, 40, r < 34 ? B(c)[r * 32 + n] : 0.f);
    }
    float* w1x = reinterpret_cast<float*>(img + L::W1X);
    for (int idx = tid; idx < 4 * 24; idx += nth) w1x[idx] = A(idx / 24)[idx % 24];
    float* b1x = reinterpret_cast<float*>(img + L::B1X);
    for (int idx = tid; idx < 4 * 8; idx += nth) b1x[idx] = (A(idx / 8) + SA::W1)[idx % 8];
}

}  // namespace qmp
using namespace qmp;

#ifdef QMP_CELL_TRACE
extern "C" __attribute__((visibility("default"))) int qmpx_cellb_trace_dump(float* host_out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out, g_cell_trace, sizeof(float) * 2 * 2048);
    if (reset) {
        static float zeros[2 * 2048];
        cudaMemcpyToSymbol(g_cell_trace, zeros, sizeof(zeros));
    }
    return 0;
}
#endif

// Bytes of the decoder-cell backward weight image.
QMP_API long long qmp_fused_cell_bwd_image_bytes(void) { return CellBwdLayout::BYTES; }

// packA [4, TOTAL(4)], packB [4, TOTAL(32)] (forward packs, fused.cuh layout) -> out [qmp_fused_cell_bwd_image_bytes()]
QMP_API int qmp_fused_pack_cell_bwd(const float* packA, const float* packB, void* out, void* stream) {
    fused_pack_cell_bwd_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(packA, packB, (uint8_t*)out);
    QMP_LAUNCH_CHECK("fused_pack_cell_bwd_kernel");
    qmp::after_producer();
    return 0;
}

// Backward of qmp_fused_cell_fwd with respect to X (dxa [N, lda]) and H (dxb [N, ldb]) -- both are OVERWRITTEN (zeroed here,
// then accumulated with reductions) -- plus the rows zB / duB [N, 128], sd [N, 64], sg [N, 32] for qmp_cell_wgrad (layout there).
// usave [N, 128] is the forward kernel's output of that name; dP [N, lddp >= 128] the gate pre-activation gradients.
// With gates != NULL the gate backward (qmp_lstm_gates_bwd's arguments: gates, Craw, Cprev, params, norm flags, eps, dHout,
// dCout, dOdirect, dHead / lddh, dCprev, dparams) runs in the prologue of every tile and dP is an OUTPUT; with gates == NULL
// dP is an input and those arguments are ignored.
QMP_API int qmp_fused_cell_bwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* xa, int lda,
                               const float* xb, int ldb, const void* image, const float* usave, float* dP, int lddp,
                               const float* gates, const float* Craw, const float* Cprev, const float* params, int norm_h, int norm_c,
                               int norm_o, float eps, const float* dHout, const float* dCout, const float* dOdirect, const float* dHead,
                               int lddh, float* dCprev, float* dparams,
                               const float* logit, const float* mstat, const float* linv, float* zB, float* duB, float* sd,
                               float* sg, float* dxa, float* dxb, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldb >= 32 && lddp % 4 == 0 && lddp >= 128 && al16(xa) && al16(xb) && al16(image) &&
                    al16(usave) && al16(dP) && al16(zB) && al16(duB) && al16(sd) && al16(sg) && al16(dxa) && al16(dxb),
                "qmp_fused_cell_bwd: rows must be 16-byte aligned");
    QMP_REQUIRE(!ea || (reinterpret_cast<uintptr_t>(ea) & 7) == 0, "qmp_fused_cell_bwd: edge attributes must be 8-byte aligned");
    CellBwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.xa = xa; a.lda = lda; a.xb = xb; a.ldb = ldb; a.usave = usave;
    a.dP = dP; a.lddp = lddp; a.logit = logit; a.mstat = mstat; a.linv = linv; a.zB = zB; a.duB = duB; a.sd = sd;
    a.gates = gates; a.Craw = Craw; a.Cprev = Cprev; a.prm = params; a.norm_h = norm_h; a.norm_c = norm_c; a.norm_o = norm_o; a.eps = eps;
    a.dHout = dHout; a.dCout = dCout; a.dOdirect = dOdirect; a.dHead = dHead; a.lddh = lddh; a.dCprev = dCprev; a.dparams = dparams;
    if (gates)
        QMP_REQUIRE(Craw && params && al16(gates) && al16(Craw) && al16(Cprev) && al16(params) && al16(dHout) && al16(dCout) &&
                        al16(dOdirect) && al16(dHead) && al16(dCprev) && (!dHead || lddh % 4 == 0),
                    "qmp_fused_cell_bwd: gate-backward rows must be 16-byte aligned");
    a.sg = sg; a.dxa = dxa; a.dxb = dxb; a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QMP_CUDA(cudaFuncSetAttribute(fused_cell_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CELLB_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    QMP_CUDA(cudaMemsetAsync(dxa, 0, (size_t)N * lda * sizeof(float), st));
    QMP_CUDA(cudaMemsetAsync(dxb, 0, (size_t)N * ldb * sizeof(float), st));
    const int G = cdiv(N, 128) < n_sm ? cdiv(N, 128) : n_sm;
    const int Q = (cdiv(N, G) + 3) & ~3;
    const int R = cdiv(Q, 128);
    const int T0 = (cdiv(Q, R) + 3) & ~3;
    QMP_CUDA(launch_pdl(fused_cell_bwd_kernel, dim3(cdiv(N, Q)), dim3(CELLB_THREADS), CELLB_SMEM, st, a, reinterpret_cast<const uint8_t*>(image), Q,
                        R, T0));
    QMP_LAUNCH_CHECK("fused_cell_bwd_kernel");
    return 0;
}
