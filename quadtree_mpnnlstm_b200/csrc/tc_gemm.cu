// tcgen05 versions of the per-node dense contractions (see gemm.cu for the semantics and tc.cuh for the
// operand layout / 3xTF32 scheme).  Same batched, strided interface as the FFMA kernels, which stay as the
// fallback for shapes that do not fit a tile.
//
//   tc_gemm_rows : C[b][i, j] (+)= sum_k A[b][i, k] * W[b][j, k] (+ bias[j]) (relu)     i over N nodes
//       one CTA = 128 threads walks (batch, 128-node tile) work items; the weights of the current batch sit
//       in shared memory (hi / lo), the node rows are staged per tile with coalesced 16-byte loads, the
//       accumulator row of node i comes back through TMEM lane i to thread i.
//   tc_gemm_tn   : C[b][i, j] += sum_r A[b][r, i] * B[b][r, j]                            r over N nodes
//       the reduction dimension is the node index: both operands are staged TRANSPOSED (K = 128 nodes per
//       tile), a CTA accumulates all of its tiles in TMEM and flushes once with atomics.
//
// These contractions are HBM/L2-bound (K is 8..136); the tensor core is used so that the dense math costs
// nothing next to the row traffic, not to chase its peak.
#include "common.cuh"
#include "tc.cuh"

namespace qmp {

struct TcGemmArgs {
    const float* A; const float* B; const float* bias; float* C;
    int n, m, k;                  // C is n x m (rows kernel) / ma x mb (tn kernel: m = ma, k = mb logical)
    int lda, ldb, ldc;
    long long sA, sB, sC, sBias;
    int batch, b_is_kxm, accumulate, relu, b_ones;
    int Kp, Np;                   // padded K (multiple of 8) and N (multiple of 16)
    int tiles;                    // 128-row tiles per batch
    int chunks;                   // tn: CTAs per batch
    uint32_t tmem_cols;
};

__device__ __forceinline__ void st_split4(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
    float4 h, l;
    tc::split_tf32(v.x, h.x, l.x);
    tc::split_tf32(v.y, h.y, l.y);
    tc::split_tf32(v.z, h.z, l.z);
    tc::split_tf32(v.w, h.w, l.w);
    *(float4*)(hi_base + off) = h;
    *(float4*)(lo_base + off) = l;
}

__global__ void __launch_bounds__(128) tc_gemm_rows_kernel(TcGemmArgs g) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int Kp = g.Kp, Np = g.Np, KC = Kp / 4;
    const uint32_t a_bytes = 128u * Kp * 4, b_bytes = (uint32_t)Np * Kp * 4;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    const int t = threadIdx.x, warp = t >> 5;

    if (t == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, g.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc_tf32(128, Np);
    const uint32_t lbo = 128, sbo = 128u * KC;
    const bool a_vec = (g.lda % 4 == 0) && (g.sA % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);

    uint32_t parity = 0;
    int loaded_batch = -1;
    const long long total = (long long)g.batch * g.tiles;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = (int)(w / g.tiles), tile = (int)(w % g.tiles);
        if (b != loaded_batch) {       // weights of this batch -> shared memory (hi / lo), zero padded
            const float* B = g.B + b * g.sB;
            for (int idx = t; idx < Np * Kp; idx += 128) {
                const int j = idx / Kp, k = idx % Kp;
                float v = 0.f;
                if (j < g.m && k < g.k) v = g.b_is_kxm ? B[(size_t)k * g.ldb + j] : B[(size_t)j * g.ldb + k];
                float hi, lo;
                tc::split_tf32(v, hi, lo);
                const uint32_t off = tc::tile_off(j, k, KC);
                *(float*)(b_hi + off) = hi;
                *(float*)(b_lo + off) = lo;
            }
            loaded_batch = b;
        }
        // node rows of this tile: 16-byte chunks, consecutive threads along K
        const float* A = g.A + b * g.sA;
        const int row0 = tile * 128;
        for (int base = 0; base < 128 * KC; base += 128 * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {          // loads first (independent), stores after
                const int idx = base + u * 128 + t;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < 128 * KC) {
                    const int r = idx / KC, c = idx % KC;
                    const int gr = row0 + r, k0 = c * 4;
                    if (gr < g.n) {
                        const float* p = A + (size_t)gr * g.lda + k0;
                        if (a_vec && k0 + 3 < g.k) {
                            v[u] = *reinterpret_cast<const float4*>(p);
                        } else {
                            if (k0 < g.k) v[u].x = p[0];
                            if (k0 + 1 < g.k) v[u].y = p[1];
                            if (k0 + 2 < g.k) v[u].z = p[2];
                            if (k0 + 3 < g.k) v[u].w = p[3];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 128 + t;
                if (idx < 128 * KC) st_split4(a_hi, a_lo, tc::tile_off(idx / KC, (idx % KC) * 4, KC), v[u]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (t == 0) {
            uint32_t acc = 0;
            for (int ks = 0; ks < Kp / 8; ++ks) {
                const uint32_t koff = (uint32_t)ks * 2 * lbo;
                const uint64_t dah = tc::make_desc(tc::smem_u32(a_hi) + koff, lbo, sbo);
                const uint64_t dal = tc::make_desc(tc::smem_u32(a_lo) + koff, lbo, sbo);
                const uint64_t dbh = tc::make_desc(tc::smem_u32(b_hi) + koff, lbo, sbo);
                const uint64_t dbl = tc::make_desc(tc::smem_u32(b_lo) + koff, lbo, sbo);
                tc::mma_tf32(tmem, dah, dbh, idesc, acc);
                tc::mma_tf32(tmem, dal, dbh, idesc, 1);
                tc::mma_tf32(tmem, dah, dbl, idesc, 1);
                acc = 1;
            }
            tc::commit(&bar);
        }
        tc::mbar_wait(&bar, parity);
        parity ^= 1;
        tc::fence_after_sync();
        // epilogue: thread t owns row t
        const int gr = row0 + t;
        float* Crow = g.C + b * g.sC + (size_t)gr * g.ldc;
        const float* bias = g.bias ? g.bias + b * g.sBias : nullptr;
        for (int c0 = 0; c0 < g.m; c0 += 8) {
            float v[8];
            tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (gr < g.n) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = c0 + i;
                    if (j < g.m) {
                        float x = v[i];
                        if (bias) x += bias[j];
                        if (g.accumulate) x += Crow[j];
                        if (g.relu) x = fmaxf(x, 0.f);
                        Crow[j] = x;
                    }
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();          // TMEM and the A tile are free for the next work item
    }
    if (warp == 0) tc::tmem_dealloc(tmem, g.tmem_cols);
}

// C[i, j] += sum_r A[r, i] B[r, j]: grid = (chunks, m-blocks of 128, batch)
__global__ void __launch_bounds__(128) tc_gemm_tn_kernel(TcGemmArgs g) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int Np = g.Np;
    constexpr int KC = 32;                                   // K = 128 nodes per tile
    const uint32_t a_bytes = 128u * 128 * 4, b_bytes = (uint32_t)Np * 128 * 4;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    const int t = threadIdx.x, warp = t >> 5;
    const int b = blockIdx.z, i0 = blockIdx.y * 128;
    const int ma_blk = min(128, g.m - i0);                   // live rows of this M block
    const int mb_real = g.b_ones ? g.k - 1 : g.k;

    if (t == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, g.tmem_cols);
    // rows of the transposed tiles that never receive data stay zero for the whole kernel
    for (int idx = t; idx < (int)((2 * a_bytes + 2 * b_bytes) / 16); idx += 128)
        reinterpret_cast<float4*>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc_tf32(128, Np);
    const uint32_t lbo = 128, sbo = 128u * KC;
    const float* A = g.A + b * g.sA;
    const float* B = g.B + b * g.sB;

    uint32_t parity = 0, acc = 0;
    for (int tile = blockIdx.x; tile < g.tiles; tile += g.chunks) {
        const int row0 = tile * 128;
        // A^T: element (i, r) <- A[row0 + r, i0 + i]; consecutive threads along i (coalesced global reads)
        for (int base = 0; base < 128 * ma_blk; base += 128 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * 128 + t;
                v[u] = 0.f;
                if (idx < 128 * ma_blk) {
                    const int r = idx / ma_blk, i = idx % ma_blk;
                    if (row0 + r < g.n) v[u] = A[(size_t)(row0 + r) * g.lda + i0 + i];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * 128 + t;
                if (idx < 128 * ma_blk) {
                    float hi, lo;
                    tc::split_tf32(v[u], hi, lo);
                    const uint32_t off = tc::tile_off(idx % ma_blk, idx / ma_blk, KC);
                    *(float*)(a_hi + off) = hi;
                    *(float*)(a_lo + off) = lo;
                }
            }
        }
        for (int base = 0; base < 128 * g.k; base += 128 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * 128 + t;
                v[u] = 0.f;
                if (idx < 128 * g.k) {
                    const int r = idx / g.k, j = idx % g.k;
                    if (row0 + r < g.n) v[u] = (j < mb_real) ? B[(size_t)(row0 + r) * g.ldb + j] : 1.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * 128 + t;
                if (idx < 128 * g.k) {
                    float hi, lo;
                    tc::split_tf32(v[u], hi, lo);
                    const uint32_t off = tc::tile_off(idx % g.k, idx / g.k, KC);
                    *(float*)(b_hi + off) = hi;
                    *(float*)(b_lo + off) = lo;
                }
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (t == 0) {
            for (int ks = 0; ks < 16; ++ks) {
                const uint32_t koff = (uint32_t)ks * 2 * lbo;
                const uint64_t dah = tc::make_desc(tc::smem_u32(a_hi) + koff, lbo, sbo);
                const uint64_t dal = tc::make_desc(tc::smem_u32(a_lo) + koff, lbo, sbo);
                const uint64_t dbh = tc::make_desc(tc::smem_u32(b_hi) + koff, lbo, sbo);
                const uint64_t dbl = tc::make_desc(tc::smem_u32(b_lo) + koff, lbo, sbo);
                tc::mma_tf32(tmem, dah, dbh, idesc, acc);
                tc::mma_tf32(tmem, dal, dbh, idesc, 1);
                tc::mma_tf32(tmem, dah, dbl, idesc, 1);
                acc = 1;
            }
            tc::commit(&bar);
        }
        tc::mbar_wait(&bar, parity);       // operands may be overwritten once the MMAs have read them
        parity ^= 1;
        tc::fence_after_sync();
        acc = 1;
        __syncthreads();
    }
    if (acc) {   // at least one tile was accumulated: flush row i = t of this M block
        float* Crow = g.C + b * g.sC + (size_t)(i0 + t) * g.ldc;
        for (int c0 = 0; c0 < g.k; c0 += 8) {
            float v[8];
            tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (t < ma_blk) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (c0 + i < g.k) atomicAdd(Crow + c0 + i, v[i]);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, g.tmem_cols);
}

static uint32_t pow2_cols(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

// returns 1 when launched, 0 when the shape does not fit (caller falls back to the FFMA kernel)
int tc_gemm_rows_try(const float* A, const float* B, const float* bias, float* C, int n, int m, int k, int lda, int ldb,
                     int ldc, long long sA, long long sB, long long sC, long long sBias, int batch, int b_is_kxm,
                     int accumulate, int relu, cudaStream_t st, int* rc) {
    *rc = 0;
    const int Kp = (k + 7) / 8 * 8, Np = (m + 15) / 16 * 16;
    if (k < 1 || Np > 256 || n < 256) return 0;
    const size_t smem = 2 * (size_t)(128 + Np) * Kp * 4;
    if (smem > 180 * 1024) return 0;
    TcGemmArgs g{};
    g.A = A; g.B = B; g.bias = bias; g.C = C; g.n = n; g.m = m; g.k = k; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.sA = sA; g.sB = sB; g.sC = sC; g.sBias = sBias; g.batch = batch; g.b_is_kxm = b_is_kxm; g.accumulate = accumulate;
    g.relu = relu; g.Kp = Kp; g.Np = Np; g.tiles = cdiv(n, 128); g.tmem_cols = pow2_cols(Np);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_error("tc_gemm_rows: %s", cudaGetErrorString(e)); *rc = (int)e; return 1; }
        attr_set = true;
    }
    const long long total = (long long)batch * g.tiles;
    const int per_sm = (int)((200 * 1024) / (smem + 1024));
    int ctas = 148 * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    if (ctas > total) ctas = (int)total;
    tc_gemm_rows_kernel<<<ctas, 128, smem, st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc_gemm_rows: %s", cudaGetErrorString(e)); *rc = (int)e; }
    return 1;
}

int tc_gemm_tn_try(const float* A, const float* B, float* C, int n, int ma, int mb, int lda, int ldb, int ldc,
                   long long sA, long long sB, long long sC, int batch, int b_ones, cudaStream_t st, int* rc) {
    *rc = 0;
    const int Np = (mb + 15) / 16 * 16;
    if (Np > 192 || n < 256) return 0;
    const size_t smem = 2 * (size_t)(128 + Np) * 128 * 4;
    if (smem > 200 * 1024) return 0;
    TcGemmArgs g{};
    g.A = A; g.B = B; g.C = C; g.n = n; g.m = ma; g.k = mb; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.sA = sA; g.sB = sB;
    g.sC = sC; g.batch = batch; g.b_ones = b_ones; g.Np = Np; g.Kp = 128; g.tiles = cdiv(n, 128);
    g.tmem_cols = pow2_cols(Np);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_error("tc_gemm_tn: %s", cudaGetErrorString(e)); *rc = (int)e; return 1; }
        attr_set = true;
    }
    const int mblocks = cdiv(ma, 128);
    int chunks = (148 + batch * mblocks - 1) / (batch * mblocks);
    if (chunks > g.tiles) chunks = g.tiles;
    if (chunks < 1) chunks = 1;
    g.chunks = chunks;
    tc_gemm_tn_kernel<<<dim3(chunks, mblocks, batch), 128, smem, st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc_gemm_tn: %s", cudaGetErrorString(e)); *rc = (int)e; }
    return 1;
}

}  // namespace qmp
