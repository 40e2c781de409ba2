// Building blocks of the tcgen05 versions of the fused cell kernels (fused_fwd_tc / fused_bwd_tc).
//
// Thread t of a 128-thread CTA owns node t of a 128-node tile AND tensor-memory lane t.  The dense per-node
// contractions of a conv (logit projection, value / skip projection and their transposes in the backward pass) are
// tcgen05.mma kind::tf32 instructions, 3xTF32 split (tc.cuh), whose A operand lives in TENSOR MEMORY: a thread
// writes its node's row (x_i, z_i, g_i, ...) to its own lane with tcgen05.st and reads the accumulator row back
// with tcgen05.ld -- no shared-memory tile, no layout shuffling, and the per-node rows never leave the SM between
// the gather phase and the contraction.  The B operand (the conv's weights, pre-split into hi / lo and laid out in
// the canonical K-major no-swizzle form by qmp_fused_pack_tc) is streamed through a double-buffered shared-memory
// slot, one conv at a time.  The edge phase in between (gathers over the CSR, segment softmax, fp32 accumulation)
// is plain SIMT code on the same thread.
//
// TMEM column map of a CTA (256 columns, two CTAs per SM):
//     [  0,128)  P   accumulators of the conv outputs: 4 gates x 32 (gate mode) or one 32-column block
//     [128,176)  U   accumulator of the current conv's first contraction (<= 48 columns)
//     [176,216)  A_hi   the A operand, high parts (<= 40 columns = K)
//     [216,256)  A_lo   low parts
#pragma once
#include "fused.cuh"
#include "tc.cuh"

namespace qmp {

constexpr uint32_t TC_COLS = 256, TC_P = 0, TC_U = 128, TC_AH = 176, TC_AL = 216;

__host__ __device__ constexpr int tc_pad8(int v) { return (v + 7) / 8 * 8; }

// byte layout of one conv's FORWARD image (B operands in canonical K-major no-swizzle form, SBO = 32 K bytes... see img_off)
struct TcFwdLayout {
    int K1, K2, N1, W1H, W1L, W3H, W3L, W2H, W2L, B1, B3, BYTES;
    __host__ __device__ constexpr TcFwdLayout(int DC)
        : K1(tc_pad8(DC)), K2(tc_pad8(DC + 4)), N1(DC + 2 <= 16 ? 16 : 48), W1H(0), W1L(W1H + N1 * K1 * 4),
          W3H(W1L + N1 * K1 * 4), W3L(W3H + FC * K1 * 4), W2H(W3L + FC * K1 * 4), W2L(W2H + FC * K2 * 4),
          B1(W2L + FC * K2 * 4), B3(B1 + 48 * 4), BYTES(B3 + FC * 4) {}
};

__host__ __device__ constexpr int tc_pad16(int v) { return (v + 15) / 16 * 16; }

// BACKWARD, target side: g (32) -> dz = W2^T g (N2 rows), dx_self = W3^T g (+ W1^T du) (N1P rows)
// W1P / B1P: the logit weights once more as plain fp32, [DC rows m][K1 columns k] = W1[k][m] and b1[k] (K1 entries): the
// one-pass mode recomputes u_i = W1[:DC] x_i + b1 with FFMAs (the transposed MMA image cannot produce it)
struct TcBwdTLayout {
    int K1, K2, N2, N1P, W2TH, W2TL, W3TH, W3TL, W1TH, W1TL, W1P, B1P, BYTES;
    __host__ __device__ constexpr TcBwdTLayout(int DC)
        : K1(tc_pad8(DC)), K2(tc_pad8(DC + 4)), N2(tc_pad16(K2)), N1P(tc_pad16(K1)), W2TH(0), W2TL(W2TH + N2 * FC * 4),
          W3TH(W2TL + N2 * FC * 4), W3TL(W3TH + N1P * FC * 4), W1TH(W3TL + N1P * FC * 4), W1TL(W1TH + N1P * K2 * 4),
          W1P(W1TL + N1P * K2 * 4), B1P(W1P + DC * K1 * 4), BYTES(B1P + K1 * 4) {}
};

// BACKWARD, source side: [av (32) | bv (K1) | sum ds, 0 ...] (KS columns) -> dx += W2[:, :D]^T av + W1[:D] bv + b1 sum ds
struct TcBwdSLayout {
    int K1, KS, N1P, BSH, BSL, BYTES;
    __host__ __device__ constexpr TcBwdSLayout(int DC)
        : K1(tc_pad8(DC)), KS(FC + K1 + 8), N1P(tc_pad16(K1)), BSH(0), BSL(N1P * KS * 4), BYTES(2 * N1P * KS * 4) {}
};

// byte offset of element (n, k) of a K-major operand with K columns (K % 8 == 0): 8-row x 16-byte core matrices,
// consecutive K chunks 128 bytes apart (LBO), consecutive 8-row blocks 32 K bytes apart (SBO)
__host__ __device__ __forceinline__ int img_off(int n, int k, int K) {
    return (n >> 3) * (32 * K) + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;
}

// D[tmem + dcol] (+)= A (TMEM, K columns at TC_AH / TC_AL) * B^T (shared memory, N rows x K, hi / lo), 3xTF32.
// Called by ONE thread.
__device__ __forceinline__ void tc_mma3_at(uint32_t tmem, uint32_t dcol, uint32_t ah_col, uint32_t al_col, uint32_t bh_addr,
                                           uint32_t bl_addr, int N, int K, bool accumulate) {
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    const uint32_t sbo = 32u * (uint32_t)K;
#pragma unroll 1
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t dbh = tc::make_desc(bh_addr + (uint32_t)ks * 256, 128, sbo);
        const uint64_t dbl = tc::make_desc(bl_addr + (uint32_t)ks * 256, 128, sbo);
        const uint32_t ah = tmem + ah_col + (uint32_t)ks * 8, al = tmem + al_col + (uint32_t)ks * 8;
        tc::mma_tf32_ts(tmem + dcol, ah, dbh, idesc, (accumulate || ks > 0) ? 1u : 0u);
#ifndef QMP_TIMING_TF32X1      // timing experiment only (scripts/kernel_times.py): how much of the step is MMA issue
        tc::mma_tf32_ts(tmem + dcol, al, dbh, idesc, 1);
        tc::mma_tf32_ts(tmem + dcol, ah, dbl, idesc, 1);
#endif
    }
}

__device__ __forceinline__ void tc_mma3(uint32_t tmem, uint32_t dcol, uint32_t bh_addr, uint32_t bl_addr, int N, int K,
                                        bool accumulate) {
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    const uint32_t sbo = 32u * (uint32_t)K;
#pragma unroll 1
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t dbh = tc::make_desc(bh_addr + (uint32_t)ks * 256, 128, sbo);
        const uint64_t dbl = tc::make_desc(bl_addr + (uint32_t)ks * 256, 128, sbo);
        const uint32_t ah = tmem + TC_AH + (uint32_t)ks * 8, al = tmem + TC_AL + (uint32_t)ks * 8;
        tc::mma_tf32_ts(tmem + dcol, ah, dbh, idesc, (accumulate || ks > 0) ? 1u : 0u);
        tc::mma_tf32_ts(tmem + dcol, al, dbh, idesc, 1);
        tc::mma_tf32_ts(tmem + dcol, ah, dbl, idesc, 1);
    }
}

// this thread's row v[0..K) -> its TMEM lane, split into hi / lo (the A operand of the next tc_mma3)
template <int K>
__device__ __forceinline__ void tc_stage_a(uint32_t lane_base, const float (&v)[K]) {
    static_assert(K % 8 == 0 && K <= 40, "A operand: K multiple of 8, at most 40 columns");
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float hi, lo;
            tc::split_tf32(v[k0 + i], hi, lo);
            h[i] = __float_as_uint(hi);
            l[i] = __float_as_uint(lo);
        }
        tc::tmem_st8(lane_base + TC_AH + (uint32_t)k0, h);
        tc::tmem_st8(lane_base + TC_AL + (uint32_t)k0, l);
    }
}

template <int K>
__device__ __forceinline__ void tc_stage_a_at(uint32_t lane_base, uint32_t ah_col, uint32_t al_col, const float (&v)[K]) {
    static_assert(K % 8 == 0, "A operand: K multiple of 8");
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float hi, lo;
            tc::split_tf32(v[k0 + i], hi, lo);
            h[i] = __float_as_uint(hi);
            l[i] = __float_as_uint(lo);
        }
        tc::tmem_st8(lane_base + ah_col + (uint32_t)k0, h);
        tc::tmem_st8(lane_base + al_col + (uint32_t)k0, l);
    }
}

// N8 * 8 accumulator columns of this thread's lane starting at column col
template <int N8>
__device__ __forceinline__ void tc_load_cols(uint32_t lane_base, uint32_t col, float (&v)[N8 * 8]) {
    uint32_t r[N8][8];
#pragma unroll
    for (int q = 0; q < N8; ++q) tc::tmem_ld8_nowait(lane_base + col + (uint32_t)q * 8, r[q]);
    tc::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < N8; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[q * 8 + i] = __uint_as_float(r[q][i]);
}

// N8 * 8 values -> columns of this thread's lane (plain fp32, e.g. a stash)
template <int N8>
__device__ __forceinline__ void tc_store_cols(uint32_t lane_base, uint32_t col, const float (&v)[N8 * 8]) {
#pragma unroll
    for (int q = 0; q < N8; ++q) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[q * 8 + i]);
        tc::tmem_st8(lane_base + col + (uint32_t)q * 8, r);
    }
    tc::tmem_st_wait();
}

// zero-padded row load into K >= D registers; vec = 16-byte aligned rows with D % 4 == 0
template <int K>
__device__ __forceinline__ void tc_load_row(float (&x)[K], const float* __restrict__ p, int D, bool vec, bool valid) {
    if (vec) {
#pragma unroll
        for (int k = 0; k < K; k += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && k < D) v = __ldg(reinterpret_cast<const float4*>(p + k));
            x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = (valid && k < D) ? __ldg(p + k) : 0.f;
    }
}

// ---- coalesced row I/O for the thread-per-node layout --------------------------------------------------------------
// A thread that reads "its" 128-byte row with eight 16-byte loads makes the warp touch 32 different cache lines per
// instruction: 256 L1 tag lookups per 32 rows, and the L1 tag stage (one line per cycle) becomes the bound of the
// fused kernels (ncu: 31 M tag requests ~ 107 us of a 118 us launch).  These helpers move the same rows with
// COALESCED instructions -- 8 lanes per row, 4 rows (4 lines) per instruction -- and transpose through a per-warp
// shared-memory tile (32 rows x 32 floats, 16-byte chunks XOR-swizzled by the row so that both the row-wise writes and
// the lane-owns-a-row reads are bank-conflict free).  All 32 lanes of the warp must call them together.
constexpr int TC_ROWTILE = 32 * 32;             // floats per tile

__device__ __forceinline__ float* tc_tile_chunk(float* tile, int r, int c) { return tile + r * 32 + ((c ^ (r & 7)) << 2); }

// this lane's row = base + j * ld (j < 0: zeros); rows are 32 floats, 16-byte aligned
__device__ __forceinline__ void warp_load_rows32(float* tile, const float* __restrict__ base, int ld, int j, float (&x)[32]) {
    const int lane = threadIdx.x & 31, c = lane & 7;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = 4 * q + (lane >> 3);
        const int jr = __shfl_sync(0xffffffffu, j, r);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jr >= 0) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)jr * ld) + c);
        *reinterpret_cast<float4*>(tc_tile_chunk(tile, r, c)) = v;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(tc_tile_chunk(tile, lane, k));
        x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
    }
    __syncwarp();
}

// two independent row sets in one go (both sets of loads are in flight together); needs two tiles
__device__ __forceinline__ void warp_load_rows32x2(float* tile, const float* __restrict__ base, int ld, int j0, int j1,
                                                   float (&x0)[32], float (&x1)[32]) {
    const int lane = threadIdx.x & 31, c = lane & 7;
    float4 v0[8], v1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = 4 * q + (lane >> 3);
        const int ja = __shfl_sync(0xffffffffu, j0, r), jb = __shfl_sync(0xffffffffu, j1, r);
        v0[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ja >= 0) v0[q] = __ldg(reinterpret_cast<const float4*>(base + (size_t)ja * ld) + c);
        if (jb >= 0) v1[q] = __ldg(reinterpret_cast<const float4*>(base + (size_t)jb * ld) + c);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = 4 * q + (lane >> 3);
        *reinterpret_cast<float4*>(tc_tile_chunk(tile, r, c)) = v0[q];
        *reinterpret_cast<float4*>(tc_tile_chunk(tile + TC_ROWTILE, r, c)) = v1[q];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(tc_tile_chunk(tile, lane, k));
        const float4 b = *reinterpret_cast<const float4*>(tc_tile_chunk(tile + TC_ROWTILE, lane, k));
        x0[4 * k] = a.x; x0[4 * k + 1] = a.y; x0[4 * k + 2] = a.z; x0[4 * k + 3] = a.w;
        x1[4 * k] = b.x; x1[4 * k + 1] = b.y; x1[4 * k + 2] = b.z; x1[4 * k + 3] = b.w;
    }
    __syncwarp();
}

// the same row index in two different arrays (e.g. gradient row and input row of one node); needs two tiles
__device__ __forceinline__ void warp_load_rows32_ab(float* tile, const float* __restrict__ baseA, int ldA,
                                                    const float* __restrict__ baseB, int ldB, int j, float (&xA)[32],
                                                    float (&xB)[32]) {
    const int lane = threadIdx.x & 31, c = lane & 7;
    float4 v0[8], v1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = 4 * q + (lane >> 3);
        const int jr = __shfl_sync(0xffffffffu, j, r);
        v0[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jr >= 0) {
            v0[q] = __ldg(reinterpret_cast<const float4*>(baseA + (size_t)jr * ldA) + c);
            v1[q] = __ldg(reinterpret_cast<const float4*>(baseB + (size_t)jr * ldB) + c);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = 4 * q + (lane >> 3);
        *reinterpret_cast<float4*>(tc_tile_chunk(tile, r, c)) = v0[q];
        *reinterpret_cast<float4*>(tc_tile_chunk(tile + TC_ROWTILE, r, c)) = v1[q];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(tc_tile_chunk(tile, lane, k));
        const float4 b = *reinterpret_cast<const float4*>(tc_tile_chunk(tile + TC_ROWTILE, lane, k));
        xA[4 * k] = a.x; xA[4 * k + 1] = a.y; xA[4 * k + 2] = a.z; xA[4 * k + 3] = a.w;
        xB[4 * k] = b.x; xB[4 * k + 1] = b.y; xB[4 * k + 2] = b.z; xB[4 * k + 3] = b.w;
    }
    __syncwarp();
}

// the warp's 32 consecutive rows row0 .. row0+31 (those < n_rows are written): lane's row x -> base + (row0 + lane) * ld
__device__ __forceinline__ void warp_store_rows32(float* tile, float* __restrict__ base, int ld, int row0, int n_rows,
                                                  const float (&x)[32]) {
    const int lane = threadIdx.x & 31, c = lane & 7;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(tc_tile_chunk(tile, lane, k)) = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = 4 * q + (lane >> 3);
        if (row0 + r < n_rows)
            *(reinterpret_cast<float4*>(base + (size_t)(row0 + r) * ld) + c) = *reinterpret_cast<const float4*>(tc_tile_chunk(tile, r, c));
    }
    __syncwarp();
}

// the first four in-edges of this thread's node (all of them on a pixel-wise mesh): loaded once per tile
struct TcEdges {
    int k0, k1;
    int j[4];
    float e0[4], e1[4];
};
__device__ __forceinline__ void tc_load_edges(TcEdges& te, const int* __restrict__ ptr, const int* __restrict__ nbr,
                                              const float* __restrict__ ea, int i, bool valid) {
    te.k0 = valid ? __ldg(ptr + i) : 0;
    te.k1 = valid ? __ldg(ptr + i + 1) : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool on = te.k0 + e < te.k1;
        te.j[e] = on ? __ldg(nbr + te.k0 + e) : 0;
        te.e0[e] = (on && ea) ? __ldg(ea + (size_t)(te.k0 + e) * 2) : 0.f;
        te.e1[e] = (on && ea) ? __ldg(ea + (size_t)(te.k0 + e) * 2 + 1) : 0.f;
    }
}

// out-edges of a node for the source-side kernel: j = target node, e0 = in-CSR slot of the edge (int bits)
__device__ __forceinline__ void tc_load_out_edges(TcEdges& te, const int* __restrict__ ptr, const int* __restrict__ dst,
                                                  const int* __restrict__ kin, int i, bool valid) {
    te.k0 = valid ? __ldg(ptr + i) : 0;
    te.k1 = valid ? __ldg(ptr + i + 1) : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool on = te.k0 + e < te.k1;
        te.j[e] = on ? __ldg(dst + te.k0 + e) : 0;
        te.e0[e] = __int_as_float(on ? __ldg(kin + te.k0 + e) : 0);
        te.e1[e] = 0.f;
    }
}

struct TcCtx {                // scalars only (no indexed members): stays in registers
    uint32_t u_col, ah_col, al_col;   // TMEM addresses (allocation base + column) of U and the A operand halves
    uint32_t p_base;                   // TMEM address of the P block (may be a separate allocation)
    uint8_t* wbase;          // double-buffered weight image slot (slot b at wbase + b * wslot), filled one conv ahead
    uint32_t wslot;
    uint64_t* wfull;         // [2] "image landed" barriers (byte-counted)
    uint32_t wpar;           // bit b = phase parity of wfull[b]
    int toggle;              // slot of the current conv
    uint64_t* bar;           // all MMA groups of this CTA commit here, every commit is waited exactly once
    uint32_t parity;
    bool pending;            // a committed group has not been waited yet
    uint32_t tmem, lane_base, lane_off;
    uint32_t stash_col;      // 96 spare TMEM columns (forward: I, F, C' between the gate slots)
    float* rtile_g;          // row tile for gradient rows (may alias rtile)
    float* rtile;            // this warp's two row tiles (shared memory) for the coalesced row I/O
};

// one thread: start the bulk copy of a conv's image into slot `buf`
__device__ __forceinline__ void tc_prefetch_image(TcCtx& cx, int buf, const uint8_t* img, uint32_t bytes) {
    tc::mbar_expect_tx(cx.wfull + buf, bytes);
    tc::bulk_g2s(cx.wbase + (size_t)buf * cx.wslot, img, bytes, cx.wfull + buf);
}

__device__ __forceinline__ void tc_wait(TcCtx& cx) {
    tc::mbar_wait(cx.bar, cx.parity);
    cx.parity ^= 1;
    cx.pending = false;
    tc::fence_after_sync();
}

}  // namespace qmp
