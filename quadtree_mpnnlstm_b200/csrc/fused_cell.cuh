// Weight image of the decoder-cell kernel (fused_cell_fwd.cu): the eight convs of one GConvLSTM step -- four on the
// 4-wide input X, four on the 32-wide hidden state H, one pair per gate (model/model.py:394-463) -- packed so that the
// dense contractions of ALL gates run as a few wide tcgen05 instructions instead of one narrow chain per conv:
//
//   W1   [160 x 32]  logit projections of the four H convs side by side: block g = rows 40g .. 40g+39 =
//                     u rows (32) | edge-attribute rows w (2) | zero (6)                       -> U  = h W1^T    (N = 160)
//   W3   [128 x 40]  skip projections of all gates, columns h (32) | x (4) | zero (4)          -> P  = [h|x] W3^T (N = 128)
//   W2_g [ 32 x 48]  value projections of gate g, columns z_h (32) | ze0 ze1 zs (3) | zero (5) | z_x (4) | ze0 ze1 zs 0
//                                                                                             -> P_g += [z_h|..|z_x|..] W2_g^T
// every matrix as a K-major B operand (no swizzle, fused_tc.cuh img_off), split into TF32 hi / lo parts, followed by the
// fp32 vectors the SIMT phases read.  Built by qmp_fused_pack_cell from the two padded packs (fused.cuh layout).
#pragma once
#include "fused_tc.cuh"

namespace qmp {

struct CellLayout {
    static constexpr int UB = 40;                         // columns per H conv in U
    static constexpr int KU = 32, NU = 4 * UB;            // U  contraction
    static constexpr int KS = 40, NS = 4 * FC;            // skip contraction
    static constexpr int KZ = 48, NZ = FC;                // value contraction of one gate
    static constexpr int W1H = 0, W1L = W1H + NU * KU * 4;
    static constexpr int W3H = W1L + NU * KU * 4, W3L = W3H + NS * KS * 4;
    static constexpr int W2H = W3L + NS * KS * 4;         // gate g: hi at W2H + g * W2G, lo NZ * KZ * 4 bytes further
    static constexpr int W2G = 2 * NZ * KZ * 4;
    static constexpr int MMA_BYTES = W2H + 4 * W2G;
    static constexpr int B1H = MMA_BYTES;                 // [4][UB]  logit biases of the H convs (u | w | 0)
    static constexpr int W1X = B1H + 4 * UB * 4;          // [4][6][4] logit weights of the X convs (u rows 0..3, w rows 4, 5)
    static constexpr int B1X = W1X + 4 * 24 * 4;          // [4][8]
    static constexpr int B3S = B1X + 4 * 8 * 4;           // [4][32]  skip biases, X conv + H conv of each gate
    static constexpr int BYTES = B3S + 4 * FC * 4;
};
static_assert(CellLayout::BYTES % 16 == 0, "bulk copies move 16-byte units");

// ---- shared by fused_cell_fwd.cu and fused_cell_bwd.cu -------------------------------------------------------------
constexpr int CELL_WORKERS = 512, CELL_THREADS = CELL_WORKERS + 32;   // 16 worker warps + the warp that issues the MMAs
constexpr int XS = 44;                        // exchange row stride in floats (conflict-free for thread-per-row 16-byte accesses)
constexpr int XPLANE = 128 * XS;              // one plane = 128 node rows; four planes (conv / gate)

#ifdef QMP_CELL_TRACE
// timeline of CTA 0 (threads 0 and 160): (tag, clock) pairs, read back by qmpx_cell_trace_dump (scripts/cell_trace.py)
static __device__ float g_cell_trace[2][2048];
static __device__ unsigned long long g_cell_cta[256][4];      // per CTA: globaltimer at entry, after the prologue, at the last tile's end, at exit
__device__ __forceinline__ unsigned long long cell_gtime() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
    return v;
}
#define CELL_CTA(k) do { if (threadIdx.x == 0 && blockIdx.x < 256) g_cell_cta[blockIdx.x][k] = cell_gtime(); } while (0)
#define CELL_MARK(tag)                                                                       \
    do {                                                                                     \
        if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 512)) {                   \
            float* tr__ = g_cell_trace[threadIdx.x ? 1 : 0];                                 \
            const int n__ = (int)tr__[0];                                                    \
            if (n__ < 1000) {                                                                \
                tr__[1 + 2 * n__] = (float)(tag);                                            \
                tr__[2 + 2 * n__] = (float)((unsigned)clock64() & 0xFFFFFFu);                \
                tr__[0] = (float)(n__ + 1);                                                  \
            }                                                                                \
        }                                                                                    \
    } while (0)
#else
#define CELL_MARK(tag) do { } while (0)
#define CELL_CTA(k) do { } while (0)
#endif

__device__ __forceinline__ void cell_stage8(uint32_t hi_addr, uint32_t lo_addr, const float (&v)[8]) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float hi, lo;
        tc::split_tf32(v[i], hi, lo);
        h[i] = __float_as_uint(hi);
        l[i] = __float_as_uint(lo);
    }
    tc::tmem_st8(hi_addr, h);
    tc::tmem_st8(lo_addr, l);
}

__device__ __forceinline__ void ld8(float (&v)[8], const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// sum over the 8 lanes of an octet of eight values per lane; lane l of the octet ends with the total of v[l]
__device__ __forceinline__ float octet_reduce8(const float (&v)[8], int l8) {
    const bool b2 = l8 & 4, b1 = l8 & 2, b0 = l8 & 1;
    float r4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b2 ? v[i + 4] : v[i], send = b2 ? v[i] : v[i + 4];
        r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    float r2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b1 ? r4[i + 2] : r4[i], send = b1 ? r4[i] : r4[i + 2];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float keep = b0 ? r2[1] : r2[0], send = b0 ? r2[0] : r2[1];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ float quad_sum(float v) {          // over the 4 lanes that differ in bits 0, 1
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float octet_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

// LayerNorm over the 32 features of a node held as 4 values in each of the 8 lanes of an octet (biased variance)
__device__ __forceinline__ void octet_layer_norm(float (&x)[4], float eps, const float4 g, const float4 b) {
    const float mean = octet_sum((x[0] + x[1]) + (x[2] + x[3])) * (1.f / FC);
    const float d0 = x[0] - mean, d1 = x[1] - mean, d2 = x[2] - mean, d3 = x[3] - mean;
    const float var = octet_sum(fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, d0 * d0)))) * (1.f / FC);
    const float rstd = rsqrtf(var + eps);
    x[0] = fmaf(d0 * rstd, g.x, b.x);
    x[1] = fmaf(d1 * rstd, g.y, b.y);
    x[2] = fmaf(d2 * rstd, g.z, b.z);
    x[3] = fmaf(d3 * rstd, g.w, b.w);
}

__device__ __forceinline__ void cell_sync() { asm volatile("bar.sync 0, %0;" ::"n"(CELL_THREADS) : "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }


}  // namespace qmp
