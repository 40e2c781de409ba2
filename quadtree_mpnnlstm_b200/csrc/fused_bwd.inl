// Fused backward of a conv layer group (counterpart of fused_fwd.inl), thread = node, two kernels:
//
//   target side (per node i, per conv): dz = W2^T dP_i;  for the edges entering i:  alpha from the saved
//     logits, d alpha = dz . x_j + ...,  ds = alpha (d alpha - sum alpha d alpha)  -> ds[E, NC];
//     z = sum alpha x_j and du = sum ds x_j (operands of the weight-gradient reductions, written once);
//     dx_i (self part) = W3^T dP_i + W1^T du.
//   source side (per node j, per conv): over the edges LEAVING j,  a = sum alpha_ij dP_i,  b = sum ds_ij x_i
//     (raw 32-wide rows, gathered through the out-CSR),  dx_j += W2[:, :D]^T a + W1 b + b1 sum ds.
//
// The source-side identity (u_i = W1 x_i + b1 and dz_i = W2^T dP_i are linear) is what removes the per-node
// u / dz intermediates of the modular path from memory.  dP comes from qmp_lstm_gates_bwd (gate mode) or is
// the upstream gradient itself (plain mode).  Weight gradients are reductions over nodes of
// dP (x) [z | x | 1] and [du] (x) [x | 1], done by qmp_gemm_tn_acc on the buffers written here.
//
// Backward weight pack per conv (cap DC): W1 [(DC+2)][DC] | b1 [DC+4] | W1T [DC][DC+4] | W2T [DC+4][32] | W3T [DC][32].
#pragma once
#include "fused.cuh"

namespace qmp {

template <int DC> struct BwdSizes {
    static constexpr int W1 = (DC + 2) * DC, B1 = DC + 4, W1T = DC * (DC + 4), W2T = (DC + 4) * FC, W3T = DC * FC;
    static constexpr int TOTAL = W1 + B1 + W1T + W2T + W3T;
};

struct FusedBwdArgs {
    int N;
    const int* ptr; const int* nbr; const int* kin; const float* ea;     // in-CSR (target) or out-CSR + kin (source)
    const float* xa; int lda; int DA; int GA; const float* wa;
    const float* xb; int ldb; int DB; int GB; int sharedB; const float* wb;
    int NC, mode, C;
    const float* dP; int lddp;
    const float* logit; const float* mstat; const float* linv;
    float* ds;
    float* ZsA; float* dUsA; float* ZsB; float* dUsB;                    // [N, G, cap+4]
    float* dxa; float* dxb; int need_dxa, need_dxb;
    int onepass;                                                         // tcgen05 target kernel only: source side by reductions (dx zero on entry)
    float drop_p; unsigned long long seed; const unsigned long long* salt;
};

// y[k] = sum_o WT[k][o] g[o] for k < R (WT row-major [R][FC])
template <int R>
__device__ __forceinline__ void matvec_T(float (&y)[R], const float* __restrict__ WT, const float (&g)[FC]) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int o = 0; o < FC; o += 4) {
            const float4 w = *reinterpret_cast<const float4*>(WT + k * FC + o);
            acc = fmaf(w.x, g[o], acc);
            acc = fmaf(w.y, g[o + 1], acc);
            acc = fmaf(w.z, g[o + 2], acc);
            acc = fmaf(w.w, g[o + 3], acc);
        }
        y[k] = acc;
    }
}

__device__ __forceinline__ void load_dP(float (&g)[FC], const FusedBwdArgs& a, int i, int c) {
    if (a.mode == 1) {
        const int slot = (c < a.GA) ? c : ((c - a.GA) & 3);      // gate fed by this conv
        load_row<FC>(g, a.dP + (size_t)i * a.lddp + (size_t)slot * FC, FC, true);
    } else {
        const float* p = a.dP + (size_t)i * a.lddp + (size_t)c * a.C;
#pragma unroll
        for (int o = 0; o < FC; ++o) g[o] = (o < a.C) ? p[o] : 0.f;
    }
}

template <int DC>
__device__ __forceinline__ void conv_bwd_target(const FusedBwdArgs& a, int i, int c, int gseg, const float* __restrict__ xin,
                                                int ld, int D, const float* __restrict__ ws, float* __restrict__ Zs,
                                                float* __restrict__ dUs, int G, float* __restrict__ dxrow, bool dx_first) {
    using S = BwdSizes<DC>;
    const float* W1T = ws + S::W1 + S::B1;
    const float* W2T = W1T + S::W1T;
    const float* W3T = W2T + S::W2T;
    const bool vec = (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(xin) & 15) == 0);
    float dz[DC + 4];
    {
        float g[FC];
        load_dP(g, a, i, c);
        matvec_T<DC + 4>(dz, W2T, g);
    }
    const float m = a.mstat[(size_t)i * a.NC + c], li = a.linv[(size_t)i * a.NC + c];
    const int k0 = a.ptr[i], k1 = a.ptr[i + 1];
    float tsum = 0.f;
    for (int kk = k0; kk < k1; ++kk) {          // pass 1: d alpha, t = sum alpha d alpha
        const int j = a.nbr[kk];
        float xj[DC];
        load_row<DC>(xj, xin + (size_t)j * ld, D, vec);
        const float a0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, a1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        float dal = fmaf(dz[DC], a0, fmaf(dz[DC + 1], a1, dz[DC + 2]));
#pragma unroll
        for (int k = 0; k < DC; ++k) dal = fmaf(dz[k], xj[k], dal);
        dal *= fdropout_scale(QMP_SEED_SM, (long long)kk * a.NC + c, a.drop_p);
        const float al = __expf(a.logit[(size_t)kk * a.NC + c] - m) * li;
        tsum = fmaf(al, dal, tsum);
        a.ds[(size_t)kk * a.NC + c] = dal;       // stash, finalised in pass 2
    }
    float du[DC + 4], z[DC + 4];
#pragma unroll
    for (int k = 0; k < DC + 4; ++k) { du[k] = 0.f; z[k] = 0.f; }
    for (int kk = k0; kk < k1; ++kk) {          // pass 2: ds, du, z
        const int j = a.nbr[kk];
        float xj[DC];
        load_row<DC>(xj, xin + (size_t)j * ld, D, vec);
        const float a0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, a1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        const float al = __expf(a.logit[(size_t)kk * a.NC + c] - m) * li;
        const float dsv = al * (a.ds[(size_t)kk * a.NC + c] - tsum);
        a.ds[(size_t)kk * a.NC + c] = dsv;
        const float alk = al * fdropout_scale(QMP_SEED_SM, (long long)kk * a.NC + c, a.drop_p);
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
    }
    store_row<DC + 4>(Zs + ((size_t)i * G + gseg) * (DC + 4), z, true);
    store_row<DC + 4>(dUs + ((size_t)i * G + gseg) * (DC + 4), du, true);
    if (dxrow) {
        float dx[DC];
        {
            float g[FC];
            load_dP(g, a, i, c);
            matvec_T<DC>(dx, W3T, g);
        }
#pragma unroll
        for (int k = 0; k < DC; ++k) {           // += W1T[k][:] . [du, dw0, dw1]
            float acc = dx[k];
            const float* w = W1T + k * (DC + 4);
#pragma unroll
            for (int r = 0; r < DC + 4; r += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(w + r);
                acc = fmaf(wv.x, du[r], acc);
                acc = fmaf(wv.y, du[r + 1], acc);
                acc = fmaf(wv.z, du[r + 2], acc);
                acc = fmaf(wv.w, du[r + 3], acc);
            }
            dx[k] = acc;
        }
#pragma unroll
        for (int k = 0; k < DC; ++k)
            if (k < D) dxrow[k] = dx_first ? dx[k] : dxrow[k] + dx[k];
    }
}

template <int DAC, int DBC>
__global__ void __launch_bounds__(128) fused_bwd_target_kernel(FusedBwdArgs a) {
    qmp_seed_init(a.seed, a.salt);
    extern __shared__ __align__(16) float sw[];
    constexpr int TA = (DAC > 0) ? BwdSizes<(DAC > 0 ? DAC : 4)>::TOTAL : 0;
    constexpr int TB = BwdSizes<DBC>::TOTAL;
    const int na = a.GA * TA, nb = a.GB * TB;
    float* swA = sw;
    float* swB = sw + na;
    for (int idx = threadIdx.x * 4; idx < na; idx += 128 * 4)
        *reinterpret_cast<float4*>(swA + idx) = *reinterpret_cast<const float4*>(a.wa + idx);
    for (int idx = threadIdx.x * 4; idx < nb; idx += 128 * 4)
        *reinterpret_cast<float4*>(swB + idx) = *reinterpret_cast<const float4*>(a.wb + idx);
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= a.N) return;
    if constexpr (DAC > 0) {
        for (int g = 0; g < a.GA; ++g)
            conv_bwd_target<DAC>(a, i, g, g, a.xa, a.lda, a.DA, swA + g * TA, a.ZsA, a.dUsA, a.GA,
                                 a.need_dxa ? a.dxa + (size_t)i * a.lda : nullptr, g == 0);
    }
    for (int g = 0; g < a.GB; ++g) {
        const int off = a.sharedB ? 0 : g * a.DB;
        conv_bwd_target<DBC>(a, i, a.GA + g, g, a.xb + off, a.ldb, a.DB, swB + g * TB, a.ZsB, a.dUsB, a.GB,
                             a.need_dxb ? a.dxb + (size_t)i * a.ldb + off : nullptr, a.sharedB ? (g == 0) : true);
    }
}

// one conv's contribution to dx_j from the edges leaving j; acc[k] += ...
template <int DC>
__device__ __forceinline__ void conv_bwd_source(const FusedBwdArgs& a, int j, int c, const float* __restrict__ xin, int ld,
                                                int D, const float* __restrict__ ws, float (&acc)[DC]) {
    using S = BwdSizes<DC>;
    const float* W1 = ws;
    const float* b1 = ws + S::W1;
    const float* W2T = ws + S::W1 + S::B1 + S::W1T;
    const bool vec = (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(xin) & 15) == 0);
    float av[FC], bv[DC];
#pragma unroll
    for (int o = 0; o < FC; ++o) av[o] = 0.f;
#pragma unroll
    for (int k = 0; k < DC; ++k) bv[k] = 0.f;
    float sds = 0.f;
    const int k1 = a.ptr[j + 1];
    for (int kk = a.ptr[j]; kk < k1; ++kk) {
        const int i = a.nbr[kk], kin = a.kin[kk];
        const float al = __expf(a.logit[(size_t)kin * a.NC + c] - a.mstat[(size_t)i * a.NC + c]) * a.linv[(size_t)i * a.NC + c] *
                         fdropout_scale(QMP_SEED_SM, (long long)kin * a.NC + c, a.drop_p);
        const float dsv = a.ds[(size_t)kin * a.NC + c];
        float g[FC];
        load_dP(g, a, i, c);
#pragma unroll
        for (int o = 0; o < FC; ++o) av[o] = fmaf(al, g[o], av[o]);
        float xi[DC];
        load_row<DC>(xi, xin + (size_t)i * ld, D, vec);
#pragma unroll
        for (int k = 0; k < DC; ++k) bv[k] = fmaf(dsv, xi[k], bv[k]);
        sds += dsv;
    }
#pragma unroll
    for (int k = 0; k < DC; ++k) {
        float s = fmaf(b1[k], sds, acc[k]);
        const float* w2 = W2T + k * FC;
#pragma unroll
        for (int o = 0; o < FC; o += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w2 + o);
            s = fmaf(w.x, av[o], s);
            s = fmaf(w.y, av[o + 1], s);
            s = fmaf(w.z, av[o + 2], s);
            s = fmaf(w.w, av[o + 3], s);
        }
        const float* w1 = W1 + k * DC;
#pragma unroll
        for (int d = 0; d < DC; d += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w1 + d);
            s = fmaf(w.x, bv[d], s);
            s = fmaf(w.y, bv[d + 1], s);
            s = fmaf(w.z, bv[d + 2], s);
            s = fmaf(w.w, bv[d + 3], s);
        }
        acc[k] = s;
    }
}

template <int DAC, int DBC>
__global__ void __launch_bounds__(128) fused_bwd_source_kernel(FusedBwdArgs a) {
    qmp_seed_init(a.seed, a.salt);
    extern __shared__ __align__(16) float sw[];
    constexpr int TA = (DAC > 0) ? BwdSizes<(DAC > 0 ? DAC : 4)>::TOTAL : 0;
    constexpr int TB = BwdSizes<DBC>::TOTAL;
    const int na = a.GA * TA, nb = a.GB * TB;
    float* swA = sw;
    float* swB = sw + na;
    for (int idx = threadIdx.x * 4; idx < na; idx += 128 * 4)
        *reinterpret_cast<float4*>(swA + idx) = *reinterpret_cast<const float4*>(a.wa + idx);
    for (int idx = threadIdx.x * 4; idx < nb; idx += 128 * 4)
        *reinterpret_cast<float4*>(swB + idx) = *reinterpret_cast<const float4*>(a.wb + idx);
    __syncthreads();
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= a.N) return;
    if constexpr (DAC > 0) {
        if (a.GA && a.need_dxa) {
            float acc[DAC];
#pragma unroll
            for (int k = 0; k < DAC; ++k) acc[k] = 0.f;
            for (int g = 0; g < a.GA; ++g) conv_bwd_source<DAC>(a, j, g, a.xa, a.lda, a.DA, swA + g * TA, acc);
            float* row = a.dxa + (size_t)j * a.lda;
#pragma unroll
            for (int k = 0; k < DAC; ++k)
                if (k < a.DA) row[k] += acc[k];
        }
    }
    if (a.need_dxb) {
        float acc[DBC];
#pragma unroll
        for (int k = 0; k < DBC; ++k) acc[k] = 0.f;
        for (int g = 0; g < a.GB; ++g) {
            const int off = a.sharedB ? 0 : g * a.DB;
            conv_bwd_source<DBC>(a, j, a.GA + g, a.xb + off, a.ldb, a.DB, swB + g * TB, acc);
            if (!a.sharedB) {
                float* row = a.dxb + (size_t)j * a.ldb + off;
#pragma unroll
                for (int k = 0; k < DBC; ++k) {
                    if (k < a.DB) row[k] += acc[k];
                    acc[k] = 0.f;
                }
            }
        }
        if (a.sharedB) {
            float* row = a.dxb + (size_t)j * a.ldb;
#pragma unroll
            for (int k = 0; k < DBC; ++k)
                if (k < a.DB) row[k] += acc[k];
        }
    }
}

template <int DAC, int DBC>
int launch_bwd(const FusedBwdArgs& a, int which, cudaStream_t st) {
    constexpr int TA = (DAC > 0) ? BwdSizes<(DAC > 0 ? DAC : 4)>::TOTAL : 0;
    const size_t smem = sizeof(float) * ((size_t)a.GA * TA + (size_t)a.GB * BwdSizes<DBC>::TOTAL);
    QMP_REQUIRE(smem <= 220 * 1024, "fused backward: weights do not fit in shared memory");
    if (which == 0) {
        auto kern = fused_bwd_target_kernel<DAC, DBC>;
        QMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<cdiv(a.N, 128), 128, smem, st>>>(a);
    } else {
        auto kern = fused_bwd_source_kernel<DAC, DBC>;
        QMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<cdiv(a.N, 128), 128, smem, st>>>(a);
    }
    QMP_LAUNCH_CHECK("fused_bwd kernel");
    return 0;
}

}  // namespace qmp
