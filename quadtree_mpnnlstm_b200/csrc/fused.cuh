// Building blocks of the fused cell kernels (fused_fwd.cu / fused_bwd.cu): one thread owns one mesh node
// and runs, for every TransformerConv that reads that node, the dense per-node contractions in registers
// against weights broadcast from shared memory, and the edge gather / segment softmax in between.
// Nothing but the node rows, the CSR, the edge logits and the final activations touches global memory:
// the per-conv intermediates of the modular path (U, Z, dZ, dU: ~0.4 GB per decoder step) never exist.
//
// Packed weights (built by pack_tconv_fused() on the host side, zero padded to the compile-time cap DC):
//   W1 [DC+2][DC]   rows 0..D-1: logit weights u = W1 x + b1; rows DC, DC+1: edge-attr weights w
//   b1 [DC+2]
//   W2 [32][DC+4]   cols 0..D-1: lin_value; DC, DC+1: lin_edge; DC+2: lin_value.bias; DC+3: 0
//   W3 [32][DC], b3 [32]   lin_skip
// (see attn.cu for the folding of PyG's q/k/v/edge/skip parameters into W1/W2/W3).
#pragma once
#include "common.cuh"

namespace qmp {

constexpr int FC = 32;                       // hidden size of the fused path

template <int DC> struct ConvSizes {
    static constexpr int W1 = (DC + 2) * DC, B1 = DC + 4, W2 = FC * (DC + 4), W3 = FC * DC, B3 = FC;
    static constexpr int TOTAL = W1 + B1 + W2 + W3 + B3;      // floats per conv in shared memory (multiple of 4)
};

// The effective dropout seed of a launch (by-value seed mixed with the device-resident salt, common.cuh) is formed ONCE per
// CTA by thread 0 and read from shared memory at every use: per edge a broadcast LDS instead of a dependent global load +
// 64-bit multiply-add (measured: the per-edge form cost 0.7 ms of the 47.9 ms step).  Kernels call qmp_seed_init() before
// their first block-wide barrier.
static __shared__ unsigned long long qmp_seed_sm;
__device__ __forceinline__ void qmp_seed_init(unsigned long long seed, const unsigned long long* salt) {
    if (threadIdx.x == 0) qmp_seed_sm = salted_seed(seed, salt);
}
#define QMP_SEED_SM (qmp::qmp_seed_sm)

// counter-based keep mask for attention dropout: same (seed, edge slot, conv) -> same decision in every kernel of the
// fused family (forward, backward target / source, FFMA or tcgen05).  32-bit mix (murmur3 finaliser) of the slot index
// and both halves of the seed: a handful of integer instructions per edge.
__device__ __forceinline__ float fdropout_scale(unsigned long long seed, long long idx, float p) {
    if (p <= 0.f) return 1.f;
    unsigned int h = (unsigned int)idx * 0x9E3779B1u + (unsigned int)seed;
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= (unsigned int)(seed >> 32);
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    return (u >= p) ? __fdividef(1.f, 1.f - p) : 0.f;
}

// zero-padded row load; vec = rows are 16-byte aligned and D % 4 == 0
template <int DC>
__device__ __forceinline__ void load_row(float (&x)[DC], const float* __restrict__ p, int D, bool vec) {
    if (vec) {
#pragma unroll
        for (int k = 0; k < DC; k += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < D) v = __ldg(reinterpret_cast<const float4*>(p + k));
            x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < DC; ++k) x[k] = (k < D) ? __ldg(p + k) : 0.f;
    }
}

// y[r] = b[r] + sum_k W[r][k] x[k], W row-major [R][DC] in shared memory (broadcast 16-byte reads)
template <int R, int DC>
__device__ __forceinline__ void matvec_rows(float (&y)[R], const float* __restrict__ W, const float* __restrict__ b,
                                            const float (&x)[DC]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float acc = b ? b[r] : 0.f;
#pragma unroll
        for (int k = 0; k < DC; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + r * DC + k);
            acc = fmaf(w.x, x[k], acc);
            acc = fmaf(w.y, x[k + 1], acc);
            acc = fmaf(w.z, x[k + 2], acc);
            acc = fmaf(w.w, x[k + 3], acc);
        }
        y[r] = acc;
    }
}

// y[r] += sum_k W[r][k] x[k] for a row stride LD (>= KC) and KC columns used
template <int R, int KC, int LD>
__device__ __forceinline__ void matvec_acc(float (&y)[R], const float* __restrict__ W, const float (&x)[KC]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float acc = y[r];
#pragma unroll
        for (int k = 0; k < KC; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + r * LD + k);
            acc = fmaf(w.x, x[k], acc);
            acc = fmaf(w.y, x[k + 1], acc);
            acc = fmaf(w.z, x[k + 2], acc);
            acc = fmaf(w.w, x[k + 3], acc);
        }
        y[r] = acc;
    }
}

// gate nonlinearities straight on the MUFU units (ex2.approx / rcp.approx, relative error ~2^-22, far inside the 1e-4
// parity bar): 4 / 5 instructions instead of the ~13 / ~16 of __expf + __fdividef with their range handling.
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_exp(float x) { return fast_ex2(x * 1.4426950408889634f); }      // exp(-inf) = 0
__device__ __forceinline__ float sigm(float x) { return fast_rcp(1.f + fast_ex2(x * -1.4426950408889634f)); }
__device__ __forceinline__ float ftanh(float x) {               // 1 - 2 / (1 + e^2x): saturates cleanly at +-1
    return fmaf(-2.f, fast_rcp(1.f + fast_ex2(x * 2.8853900817779268f)), 1.f);
}

// LayerNorm over FC register values (biased variance, like torch.nn.LayerNorm)
__device__ __forceinline__ void ln_stats(const float (&x)[FC], float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < FC; ++c) s += x[c];
    mean = s * (1.f / FC);
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < FC; ++c) {
        const float d = x[c] - mean;
        v = fmaf(d, d, v);
    }
    rstd = rsqrtf(v * (1.f / FC) + eps);
}

template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&x)[N], bool vec) {
    if (vec) {
#pragma unroll
        for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(p + k) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) p[k] = x[k];
    }
}

}  // namespace qmp
