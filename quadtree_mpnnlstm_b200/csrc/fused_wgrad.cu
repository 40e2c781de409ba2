// Weight gradients of a fused conv layer group (counterpart of fused_fwd / fused_bwd) in ONE launch, on tcgen05.
//
// For every conv c of the group the gradient of its packed weights (layout: fused.cuh) is a reduction over the
// mesh nodes of outer products of per-node rows that the backward target kernel has already produced:
//
//     gW1 [(DC+2) x DC] = sum_i dU_i (x) x_i        gb1 = sum_i dU_i              dU = [du | dw]  (DC+4, zero padded)
//     gW2 [32 x (DC+4)] = sum_i g_i  (x) Z_i                                       Z  = [z | ze | zs | 0]
//     gW3 [32 x DC]     = sum_i g_i  (x) x_i        gb3 = sum_i g_i               g  = dP rows of the conv's gate
//
// i.e. one "TN" GEMM per conv with the node index as the reduction dimension K:
//     D [M x Nb] = A B^T,   A [m, i] = [dU_i | g_i][m]  (M = DC + 36 rows used of 128),   B [n, i] = [x_i | 1 0 0 0 | Z_i][n]
// Both operands are node-major in memory (the transposes of what a K-major MMA wants):
//   * A goes through TENSOR MEMORY: thread m reads component m of 64 consecutive nodes (a warp reads 128 contiguous
//     bytes per node) and writes them to TMEM lane m with tcgen05.st -- lane = M row, column = node -- which is the
//     K-major A operand of tcgen05.mma's TMEM-A form.  No shared memory, no transposition code.
//   * B goes through shared memory in the canonical K-major no-swizzle layout (tc.cuh): a thread loads the same
//     16-byte column chunk of 4 consecutive nodes, transposes the 4 x 4 block in registers and stores four 16-byte
//     (n, 4 nodes) pieces; the 8-row block stride is padded by 16 bytes so the stores of a quarter-warp hit 32 banks.
// (The MN-major descriptor form that would take both operands as they lie returns zeros for kind::tf32 with the
// no-swizzle layout on this part -- scripts/tc_probe2.py variants 1/3 -- so it is not used.)
// 3xTF32 (tc.cuh): A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in TMEM.  A CTA owns one conv and a strided
// set of 64-node tiles, accumulates all of them in TMEM and flushes once with atomics: the node rows are read exactly
// once and nothing else touches memory.
#include "common.cuh"
#include "fused.cuh"
#include "tc.cuh"

namespace qmp {

constexpr int WG_KT = 64;                       // nodes per tile (K of the tile)
constexpr int WG_KC = WG_KT / 4;                // 16-byte K chunks per B row
constexpr int WG_SBO = WG_KC * 128 + 16;        // bytes between consecutive 8-row blocks of B (padded)
constexpr int WG_MAXCONV = 12;
constexpr int WG_BIT = 2;                       // B items per thread: 16 node groups x (<= 32 chunks) / 256 threads
constexpr uint32_t WG_TMEM_COLS = 256, WG_AHI = 128, WG_ALO = 192;

// One reduction problem = one or two convs ("parts") that share the gradient rows g (in gate mode conv_x_g and conv_h_g
// feed the same gate, so their g is the same dP block): rows of A = [dU_0 | dU_1 | g], columns of B =
// [x_0 | 1 0 0 0 | Z_0 | x_1 | Z_1].  Merging halves the MMA count of the decoder cell (issue cost is flat in N).
struct WgPart {
    const float* x; const float* zs; const float* dus; float* gw;
    int ldx, ldz, D, DC;
};
struct WgConv {
    WgPart p[2];
    const float* g;
    int nparts, ldg, gvalid, cta0, nctas;
};
struct WgArgs {
    int N, nconv;
    WgConv c[WG_MAXCONV];
};

__device__ __forceinline__ float4 wg_load4(const float* __restrict__ p, int nvalid, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid >= 4 && vec) return __ldg(reinterpret_cast<const float4*>(p));
    if (nvalid > 0) v.x = __ldg(p);
    if (nvalid > 1) v.y = __ldg(p + 1);
    if (nvalid > 2) v.z = __ldg(p + 2);
    if (nvalid > 3) v.w = __ldg(p + 3);
    return v;
}

__device__ __forceinline__ void wg_st_split(uint8_t* hi, uint8_t* lo, uint32_t off, float a, float b, float c, float d) {
    float4 h, l;
    tc::split_tf32(a, h.x, l.x);
    tc::split_tf32(b, h.y, l.y);
    tc::split_tf32(c, h.z, l.z);
    tc::split_tf32(d, h.w, l.w);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

constexpr int WG_WORKERS = 256;                 // 8 staging warps: warps w and w + 4 share TMEM lane quarter w, each takes half of a tile's nodes
constexpr int WG_THREADS = WG_WORKERS + 32;     // + the warp that issues the MMAs
constexpr int WG_KH = WG_KT / 2;                // nodes per tile staged by one thread
__device__ __forceinline__ void wg_sync() { asm volatile("bar.sync 0, %0;" ::"n"(WG_THREADS) : "memory"); }
__device__ __forceinline__ void wg_sync_workers() { asm volatile("bar.sync 1, %0;" ::"n"(WG_WORKERS) : "memory"); }

#ifdef QMP_CELL_TRACE
static __device__ long long g_wg_trace[64];
#define WG_MARK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (k) < 64) g_wg_trace[k] = clock64(); } while (0)
#else
#define WG_MARK(k) do { } while (0)
#endif

// 96 registers: registers are handed out per warp PAIR, so two 9-warp CTAs per SM count as 20 warps (65536 / 20 / 32 = 102);
// __maxnreg__(112) removed the spills of the hoisted per-item state but left ONE CTA per SM (measured: 94 -> 129 us)
__global__ void __launch_bounds__(WG_THREADS, 2) fused_wgrad_kernel(const __grid_constant__ WgArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    WG_MARK(0);
    int ci = 0;
    while (ci + 1 < a.nconv && (int)blockIdx.x >= a.c[ci].cta0 + a.c[ci].nctas) ++ci;
    const WgConv& cv = a.c[ci];
    const bool two = cv.nparts == 2;
    const int DC0 = cv.p[0].DC, DC1 = two ? cv.p[1].DC : 0;
    // A rows
    const int RA1 = DC0 + 4, RG = RA1 + (two ? DC1 + 4 : 0);
    // B chunks (16 bytes = 4 columns): x_0 | ones | Z_0 | x_1 | Z_1
    const int q_one = DC0 / 4, q_z0 = q_one + 1, q_x1 = q_z0 + (DC0 + 4) / 4, q_z1 = q_x1 + DC1 / 4;
    const int qb = two ? q_z1 + (DC1 + 4) / 4 : q_x1;
    const int NB = (qb * 4 + 15) / 16 * 16;
    uint8_t* b_hi = smem;
    uint8_t* b_lo = smem + (NB / 8) * WG_SBO;

    if (t == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, WG_TMEM_COLS);
    for (int idx = t; idx < 2 * (NB / 8) * WG_SBO / 16; idx += WG_THREADS) reinterpret_cast<float4*>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_before_sync();
    wg_sync();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc = tc::make_idesc_tf32(128, NB);

    // A: this thread's component of [dU_0 | dU_1 | g]
    const int tr = t & 127, half = (t >> 7) & 1;          // A row (= TMEM lane) and node half of this thread
    const float* a_src = nullptr;
    int a_ld = 0;
    if (tr < RA1) { a_src = cv.p[0].dus + tr; a_ld = cv.p[0].ldz; }
    else if (tr < RG) { a_src = cv.p[1].dus + (tr - RA1); a_ld = cv.p[1].ldz; }
    else if (tr < RG + FC && tr - RG < cv.gvalid) { a_src = cv.g + (tr - RG); a_ld = cv.ldg; }
    const bool vx0 = (cv.p[0].ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(cv.p[0].x) & 15) == 0);
    const bool vx1 = two && (cv.p[1].ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(cv.p[1].x) & 15) == 0);
    const int ntiles = (a.N + WG_KT - 1) / WG_KT;
    const int items = (WG_KT / 4) * qb;

    // Everything about this thread's loads that does not depend on the tile is resolved once, here: the loop body below is
    // issue-bound (CTA timeline: 7 k of 8.9 k cycles per tile went into ISSUING the old fetch -- per-element bound checks,
    // a five-way column-range dispatch and two integer divisions per item and tile), not latency- or bandwidth-bound.
    float va[WG_KH];
    float4 vb[WG_BIT][4];
#pragma unroll
    for (int k = 0; k < WG_KH; ++k) va[k] = 0.f;           // rows without a source stay zero
    const float* b_ptr[WG_BIT];                            // fast items (whole 16-byte chunks of an aligned row): chunk of node 0
    int b_ld[WG_BIT], b_kg[WG_BIT];
    uint32_t b_off[WG_BIT];                                // byte offset of the item's 4 x 4 block in the B tile
#pragma unroll
    for (int it = 0; it < WG_BIT; ++it) {
        const int idx = t + WG_WORKERS * it;
        const bool on = idx < items;
        const int kg = on ? idx / qb : 0, q = on ? idx - kg * qb : 0;
        b_kg[it] = kg;
        b_off[it] = (uint32_t)(((q * 4) >> 3) * WG_SBO + kg * 128 + ((q * 4) & 7) * 16);
        b_ptr[it] = nullptr; b_ld[it] = 0;
        if (on && q != q_one) {
            if (q < q_one) { if (vx0 && cv.p[0].D - q * 4 >= 4) { b_ptr[it] = cv.p[0].x + q * 4; b_ld[it] = cv.p[0].ldx; } }
            else if (q < q_x1) { b_ptr[it] = cv.p[0].zs + (q - q_z0) * 4; b_ld[it] = cv.p[0].ldz; }
            else if (q < q_z1) { if (vx1 && cv.p[1].D - (q - q_x1) * 4 >= 4) { b_ptr[it] = cv.p[1].x + (q - q_x1) * 4; b_ld[it] = cv.p[1].ldx; } }
            else { b_ptr[it] = cv.p[1].zs + (q - q_z1) * 4; b_ld[it] = cv.p[1].ldz; }
        }
    }
    auto fetch = [&](int tile) {
        const int n0 = tile * WG_KT;
        const bool full = n0 + WG_KT <= a.N;               // all but the last tile: no bound checks
        if (a_src != nullptr) {
            const float* p = a_src + (size_t)(n0 + half * WG_KH) * a_ld;
            if (full) {
#pragma unroll
                for (int k = 0; k < WG_KH; ++k) va[k] = __ldg(p + (size_t)k * a_ld);
            } else {
#pragma unroll
                for (int k = 0; k < WG_KH; ++k) va[k] = (n0 + half * WG_KH + k < a.N) ? __ldg(p + (size_t)k * a_ld) : 0.f;
            }
        }
#pragma unroll
        for (int it = 0; it < WG_BIT; ++it) {
            if (b_ptr[it] != nullptr && full) {
                const float* p = b_ptr[it] + (size_t)(n0 + b_kg[it] * 4) * b_ld[it];
#pragma unroll
                for (int j = 0; j < 4; ++j) vb[it][j] = __ldg(reinterpret_cast<const float4*>(p + (size_t)j * b_ld[it]));
            } else {                                       // the ones column, ragged rows, the last tile, no item
                const int idx = t + WG_WORKERS * it;
#pragma unroll
                for (int j = 0; j < 4; ++j) vb[it][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < items) {
                    const int kg = b_kg[it], q = idx - kg * qb;
#pragma unroll 1
                    for (int j = 0; j < 4; ++j) {
                        const int i = n0 + kg * 4 + j;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i < a.N) {
                            if (q < q_one) v = wg_load4(cv.p[0].x + (size_t)i * cv.p[0].ldx + q * 4, cv.p[0].D - q * 4, vx0);
                            else if (q == q_one) v.x = 1.f;
                            else if (q < q_x1) v = wg_load4(cv.p[0].zs + (size_t)i * cv.p[0].ldz + (q - q_z0) * 4, 4, true);
                            else if (q < q_z1) v = wg_load4(cv.p[1].x + (size_t)i * cv.p[1].ldx + (q - q_x1) * 4, cv.p[1].D - (q - q_x1) * 4, vx1);
                            else v = wg_load4(cv.p[1].zs + (size_t)i * cv.p[1].ldz + (q - q_z1) * 4, 4, true);
                        }
                        if (j == 0) vb[it][0] = v; else if (j == 1) vb[it][1] = v; else if (j == 2) vb[it][2] = v; else vb[it][3] = v;
                    }
                }
            }
        }
    };
    auto stage = [&]() {
        // A -> TMEM lane t, columns = nodes (hi at WG_AHI, lo at WG_ALO)
#pragma unroll
        for (int k0 = 0; k0 < WG_KH; k0 += 8) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float hi, lo;
                tc::split_tf32(va[k0 + i], hi, lo);
                h[i] = __float_as_uint(hi);
                l[i] = __float_as_uint(lo);
            }
            tc::tmem_st8(lane_base + WG_AHI + (uint32_t)(half * WG_KH + k0), h);
            tc::tmem_st8(lane_base + WG_ALO + (uint32_t)(half * WG_KH + k0), l);
        }
        // B -> shared memory, K-major: element (n, node k) at (n/8)*SBO + (k/4)*128 + (n%8)*16 + (k%4)*4
#pragma unroll
        for (int it = 0; it < WG_BIT; ++it) {
            if (t + WG_WORKERS * it < items) {
                const uint32_t base = b_off[it];
                wg_st_split(b_hi, b_lo, base, vb[it][0].x, vb[it][1].x, vb[it][2].x, vb[it][3].x);
                wg_st_split(b_hi, b_lo, base + 16, vb[it][0].y, vb[it][1].y, vb[it][2].y, vb[it][3].y);
                wg_st_split(b_hi, b_lo, base + 32, vb[it][0].z, vb[it][1].z, vb[it][2].z, vb[it][3].z);
                wg_st_split(b_hi, b_lo, base + 48, vb[it][0].w, vb[it][1].w, vb[it][2].w, vb[it][3].w);
            }
        }
        tc::tmem_st_wait();
    };

    uint32_t parity = 0, acc = 0;
    int tile = (int)blockIdx.x - cv.cta0;
    if (warp == WG_WORKERS / 32) {
        // ---- the MMA warp: an issuing thread is held for the ~90 cycles each MMA occupies the tensor pipe; the staging
        // warps must not be (they fetch the next tile's rows meanwhile)
        for (; tile < ntiles; tile += cv.nctas) {
            wg_sync();                                         // operands of this tile staged
            if ((t & 31) == 0) {
                tc::fence_after_sync();
#pragma unroll 1
                for (int ks = 0; ks < WG_KT / 8; ++ks) {
                    const uint64_t dbh = tc::make_desc(tc::smem_u32(b_hi) + (uint32_t)ks * 256, 128, WG_SBO);
                    const uint64_t dbl = tc::make_desc(tc::smem_u32(b_lo) + (uint32_t)ks * 256, 128, WG_SBO);
                    const uint32_t ah = tmem + WG_AHI + (uint32_t)ks * 8, al = tmem + WG_ALO + (uint32_t)ks * 8;
                    tc::mma_tf32_ts(tmem, ah, dbh, idesc, (acc | ks) ? 1u : 0u);
                    tc::mma_tf32_ts(tmem, al, dbh, idesc, 1);
                    tc::mma_tf32_ts(tmem, ah, dbl, idesc, 1);
                }
                tc::commit(&bar);
            }
            __syncwarp();
            acc = 1;
        }
    } else {
        WG_MARK(1);
        if (tile < ntiles) fetch(tile);
        int wk = 2;
        for (; tile < ntiles; tile += cv.nctas) {
            if (acc) {                   // the previous tile's MMAs must have read the operands before they are overwritten
                tc::mbar_wait(&bar, parity);
                parity ^= 1;
                tc::fence_after_sync();
            }
            WG_MARK(wk); ++wk;
            stage();
            tc::fence_async_smem();
            tc::fence_before_sync();
            WG_MARK(wk); ++wk;
            wg_sync();
            acc = 1;
            if (tile + cv.nctas < ntiles) fetch(tile + cv.nctas);     // next tile's rows fly while the tensor core works
            WG_MARK(wk); ++wk;
        }
        WG_MARK(60);
        if (acc) {
            tc::mbar_wait(&bar, parity);
            tc::fence_after_sync();
            WG_MARK(61);
            // flush.  Accumulator row t (thread t's TMEM lane) -> shared memory, then warps walk the rows with their lanes
            // along the columns so that one reduction instruction touches 4 sectors instead of 32 (the L2 atomic units
            // see 8x fewer transactions).  Pack offsets (fused.cuh): W1 | b1 | W2 | W3 | b3
            const int c_one = q_one * 4, c_z0 = q_z0 * 4, c_x1 = q_x1 * 4, c_z1 = q_z1 * 4, c_end = qb * 4;
            float* ds = reinterpret_cast<float*>(smem);                     // [128][c_end + 1], the operand tiles are dead
            const int ldd = c_end + 1;
            for (int c0 = 8 * half; c0 < c_end; c0 += 16) {                 // the two warps of a lane quarter alternate column blocks
                float v[8];
                tc::tmem_ld8(lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (c0 + i < c_end) ds[tr * ldd + c0 + i] = v[i];
            }
            wg_sync_workers();
            const int o1a = (DC0 + 2) * DC0, o2a = o1a + DC0 + 4, o3a = o2a + FC * (DC0 + 4), o4a = o3a + FC * DC0;
            const int o1b = (DC1 + 2) * DC1, o2b = o1b + DC1 + 4, o3b = o2b + FC * (DC1 + 4), o4b = o3b + FC * DC1;
            float* gwa = cv.p[0].gw;
            float* gwb = two ? cv.p[1].gw : nullptr;
            const int lane = t & 31;
            for (int row = warp; row < RG + FC; row += WG_WORKERS / 32) {
                // row roles: 0 = dU row of part 0, 1 = dU row of part 1, 2 = g row, 3 = nothing
                int role = 3, r = 0;
                if (row < DC0 + 2) { role = 0; r = row; }
                else if (two && row >= RA1 && row < RA1 + DC1 + 2) { role = 1; r = row - RA1; }
                else if (row >= RG && row - RG < cv.gvalid) { role = 2; r = row - RG; }
                if (role == 3) continue;
                for (int col = lane; col < c_end; col += 32) {
                    const float v = ds[row * ldd + col];
                    if (col < c_one) {                                   // x_0
                        if (role == 0) atomicAdd(gwa + r * DC0 + col, v);
                        else if (role == 2) atomicAdd(gwa + o3a + r * DC0 + col, v);
                    } else if (col == c_one) {                           // ones
                        if (role == 0) atomicAdd(gwa + o1a + r, v);
                        else if (role == 1) atomicAdd(gwb + o1b + r, v);
                        else {
                            atomicAdd(gwa + o4a + r, v);
                            if (two) atomicAdd(gwb + o4b + r, v);
                        }
                    } else if (col >= c_z0 && col < c_x1) {              // Z_0
                        if (role == 2) atomicAdd(gwa + o2a + r * (DC0 + 4) + (col - c_z0), v);
                    } else if (col >= c_x1 && col < c_z1) {              // x_1
                        if (role == 1) atomicAdd(gwb + r * DC1 + (col - c_x1), v);
                        else if (role == 2) atomicAdd(gwb + o3b + r * DC1 + (col - c_x1), v);
                    } else if (col >= c_z1) {                            // Z_1
                        if (role == 2) atomicAdd(gwb + o2b + r * (DC1 + 4) + (col - c_z1), v);
                    }
                }
            }
        }
    }
    WG_MARK(62);
    tc::fence_before_sync();
    wg_sync();
    WG_MARK(63);
    if (warp == 0) tc::tmem_dealloc(tmem, WG_TMEM_COLS);
}

}  // namespace qmp
using namespace qmp;

#ifdef QMP_CELL_TRACE
extern "C" __attribute__((visibility("default"))) int qmpx_wg_trace_dump(long long* host_out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out, g_wg_trace, sizeof(long long) * 64);
    return 0;
}
#endif

// Weight gradients of one fused layer group (see the top of this file): accumulates into gwa [GA, TOTAL(cap DA)] /
// gwb [GB, TOTAL(cap DB)] (forward pack layout, caller zero-initialises).  Arguments as qmp_fused_bwd_target, whose
// outputs ZsA / dUsA [N, GA, capA+4] and ZsB / dUsB [N, GB, capB+4] are this kernel's inputs.
QMP_API int qmp_fused_wgrad(int N, const float* xa, int lda, int DA, int GA, const float* xb, int ldb, int DB, int GB,
                            int sharedB, int mode, int C, const float* dP, int lddp, const float* ZsA, const float* dUsA,
                            const float* ZsB, const float* dUsB, float* gwa, float* gwb, void* stream) {
    if (N <= 0) return 0;
    const int NC = GA + GB;
    QMP_REQUIRE(NC >= 1 && NC <= WG_MAXCONV, "qmp_fused_wgrad: at most %d convs per group", WG_MAXCONV);
    QMP_REQUIRE(DB >= 1 && DB <= 36 && DA >= 0 && DA <= 8 && C >= 1 && C <= FC, "qmp_fused_wgrad: unsupported sizes");
    const int dac = (GA == 0) ? 0 : (DA <= 4 ? 4 : 8);
    const int dbc = (DB <= 32) ? 32 : 36;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    WgArgs a{};
    a.N = N;
    const int ntiles = cdiv(N, WG_KT);
    auto fill_part = [&](WgPart& p, int c) {
        const bool segA = c < GA;
        const int g = segA ? c : c - GA;
        p.DC = segA ? dac : dbc;
        p.D = segA ? DA : DB;
        const int G = segA ? GA : GB, W = p.DC + 4;
        p.x = segA ? xa : xb + (sharedB ? 0 : (size_t)g * DB);
        p.ldx = segA ? lda : ldb;
        p.zs = (segA ? ZsA : ZsB) + (size_t)g * W;
        p.dus = (segA ? dUsA : dUsB) + (size_t)g * W;
        p.ldz = G * W;
        const int total = (p.DC + 2) * p.DC + (p.DC + 4) + FC * (p.DC + 4) + FC * p.DC + FC;
        p.gw = (segA ? gwa : gwb) + (size_t)g * total;
    };
    // gate mode with one narrow (X) and one wide (H) conv per gate: the pair shares the gate's gradient rows -> one problem
    const bool merge = mode == 1 && GA == 4 && GB == 4;
    int nprob = 0;
    int w[WG_MAXCONV], wsum = 0, nbmax = 16;
    if (merge) {
        for (int s = 0; s < 4; ++s) {
            WgConv& v = a.c[nprob++];
            v.nparts = 2;
            fill_part(v.p[0], s);
            fill_part(v.p[1], GA + s);
            v.g = dP + (size_t)s * FC;
            v.gvalid = FC;
            v.ldg = lddp;
        }
    } else {
        for (int c = 0; c < NC; ++c) {
            WgConv& v = a.c[nprob++];
            v.nparts = 1;
            fill_part(v.p[0], c);
            v.p[1] = v.p[0];
            if (mode == 1) {
                v.g = dP + (size_t)(c < GA ? c : ((c - GA) & 3)) * FC;
                v.gvalid = FC;
            } else {
                v.g = dP + (size_t)c * C;
                v.gvalid = C;
            }
            v.ldg = lddp;
        }
    }
    a.nconv = nprob;
    for (int k = 0; k < nprob; ++k) {
        const WgConv& v = a.c[k];
        int qb = v.p[0].DC / 4 + 1 + (v.p[0].DC + 4) / 4;
        if (v.nparts == 2) qb += v.p[1].DC / 4 + (v.p[1].DC + 4) / 4;
        w[k] = 16 + qb;
        wsum += w[k];
        const int nb = (qb * 4 + 15) / 16 * 16;
        nbmax = nb > nbmax ? nb : nbmax;
    }
    QMP_REQUIRE(nbmax <= 128, "qmp_fused_wgrad: problem too wide for the accumulator block");
    const int budget = 2 * n_sm;
    int cta = 0;
    for (int k = 0; k < nprob; ++k) {
        int n = (int)((long long)budget * w[k] / wsum);
        n = n < 1 ? 1 : (n > ntiles ? ntiles : n);
        a.c[k].cta0 = cta;
        a.c[k].nctas = n;
        cta += n;
    }
    size_t smem = (size_t)2 * (nbmax / 8) * WG_SBO;
    if (smem < (size_t)128 * (nbmax + 1) * sizeof(float)) smem = (size_t)128 * (nbmax + 1) * sizeof(float);   // flush tile
    static size_t smem_set = 0;
    if (smem > smem_set) {
        QMP_CUDA(cudaFuncSetAttribute(fused_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    fused_wgrad_kernel<<<cta, WG_THREADS, smem, (cudaStream_t)stream>>>(a);
    QMP_LAUNCH_CHECK("fused_wgrad_kernel");
    return 0;
}
