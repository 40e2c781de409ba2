// Fused forward of one conv "layer group" of the graph-conv LSTM cell (TransformerConv, hidden 32):
// per node, every conv of the group (logit projection -> edge gather + segment softmax + aggregation ->
// output projection + skip) and, in gate mode, the LSTM gate math with its LayerNorms and the decoder
// head input -- one launch instead of ~9, no intermediate tensors.
//
// Reference: GConvLSTM.forward (model/model.py:394-463) around PyG TransformerConv (model/model.py:51),
// Encoder/Decoder norms (model/seq2seq.py:59-66, 138-165).  Math identical to the modular kernels
// (attn.cu, gemm.cu, lstm.cu), which remain the general path and the cross-check in the tests.
//
// Thread = node.  Segment A: GA convs sharing a narrow input (the cell's X, D <= 8).  Segment B: GB convs on
// 32..36-wide inputs, shared (the cell's H) or one input block per conv (deeper layers of a stack).
// mode 1 (gates): conv_A[g] + conv_B[g] (+ conv_B[4+g]) accumulate into gate g = i, f, c, o.
// mode 0 (plain): conv c writes out[:, c*C .. c*C+C) (optionally through relu).
#pragma once
#include "fused.cuh"

namespace qmp {

struct FusedFwdArgs {
    int N;
    const int* ptr; const int* nbr; const float* ea;
    const float* xa; int lda; int DA; int GA; const float* wa;
    const float* xb; int ldb; int DB; int GB; int sharedB; const float* wb;
    int NC, mode, relu_out, C;
    float* out; int ldo;
    const float* Cprev; const float* params; int norm_h, norm_c, norm_o; float eps;
    float* gates; float* Craw; float* Oout; float* Hout; float* Cout; float* head_in; int ldh; const float* concat;
    float* logit; float* mstat; float* linv;
    float drop_p; unsigned long long seed; const unsigned long long* salt;
};

template <int DC>
__device__ __forceinline__ void conv_accumulate(const FusedFwdArgs& a, int i, int c, const float* __restrict__ xin, int ld,
                                                int D, const float* __restrict__ ws, float (&P)[FC]) {
    using S = ConvSizes<DC>;
    const float* W1 = ws;
    const float* b1 = W1 + S::W1;
    const float* W2 = b1 + S::B1;
    const float* W3 = W2 + S::W2;
    const float* b3 = W3 + S::W3;
    const bool vec = (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(xin) & 15) == 0);

    float u[DC], w01[2];
    {
        float x[DC];
        load_row<DC>(x, xin + (size_t)i * ld, D, vec);
        matvec_rows<DC, DC>(u, W1, b1, x);
        matvec_rows<2, DC>(w01, W1 + DC * DC, b1 + DC, x);
    }
    float z[DC];
#pragma unroll
    for (int k = 0; k < DC; ++k) z[k] = 0.f;
    float m = -INFINITY, l = 0.f, zs = 0.f, ze0 = 0.f, ze1 = 0.f;
    const int k1 = a.ptr[i + 1];
    for (int kk = a.ptr[i]; kk < k1; ++kk) {
        const int j = a.nbr[kk];
        float xj[DC];
        load_row<DC>(xj, xin + (size_t)j * ld, D, vec);
        const float a0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, a1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        float s = fmaf(w01[0], a0, w01[1] * a1);
#pragma unroll
        for (int k = 0; k < DC; ++k) s = fmaf(u[k], xj[k], s);
        a.logit[(size_t)kk * a.NC + c] = s;
        const float mn = fmaxf(m, s);
        const float sc = __expf(m - mn), p = __expf(s - mn);
        const float pk = p * fdropout_scale(QMP_SEED_SM, (long long)kk * a.NC + c, a.drop_p);
        l = fmaf(l, sc, p);
        zs = fmaf(zs, sc, pk);
        ze0 = fmaf(ze0, sc, pk * a0);
        ze1 = fmaf(ze1, sc, pk * a1);
#pragma unroll
        for (int k = 0; k < DC; ++k) z[k] = fmaf(z[k], sc, pk * xj[k]);
        m = mn;
    }
    const float li = (l > 0.f) ? 1.f / l : 0.f;
    a.mstat[(size_t)i * a.NC + c] = m;
    a.linv[(size_t)i * a.NC + c] = li;
#pragma unroll
    for (int k = 0; k < DC; ++k) z[k] *= li;
    ze0 *= li; ze1 *= li; zs *= li;

    float x[DC];
    load_row<DC>(x, xin + (size_t)i * ld, D, vec);
#pragma unroll
    for (int o = 0; o < FC; ++o) {
        float acc = P[o] + b3[o];
        const float* w2 = W2 + o * (DC + 4);
#pragma unroll
        for (int k = 0; k < DC; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w2 + k);
            acc = fmaf(w.x, z[k], acc);
            acc = fmaf(w.y, z[k + 1], acc);
            acc = fmaf(w.z, z[k + 2], acc);
            acc = fmaf(w.w, z[k + 3], acc);
        }
        const float4 wt = *reinterpret_cast<const float4*>(w2 + DC);
        acc = fmaf(wt.x, ze0, acc);
        acc = fmaf(wt.y, ze1, acc);
        acc = fmaf(wt.z, zs, acc);
        const float* w3 = W3 + o * DC;
#pragma unroll
        for (int k = 0; k < DC; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w3 + k);
            acc = fmaf(w.x, x[k], acc);
            acc = fmaf(w.y, x[k + 1], acc);
            acc = fmaf(w.z, x[k + 2], acc);
            acc = fmaf(w.w, x[k + 3], acc);
        }
        P[o] = acc;
    }
}

// params rows (lstm.cu): 0 wci 1 wcf 2 wco 3 bi 4 bf 5 bc 6 bo 7 gh 8 bh 9 gc 10 bc 11 go 12 bo
__device__ __forceinline__ void gate_epilogue(const FusedFwdArgs& a, int i, int s, const float* __restrict__ prm, float (&P)[FC]) {
    float* gs = a.gates + (size_t)i * 4 * FC;
    float cp[FC];
    if (a.Cprev) load_row<FC>(cp, a.Cprev + (size_t)i * FC, FC, true);
    else {
#pragma unroll
        for (int o = 0; o < FC; ++o) cp[o] = 0.f;
    }
    if (s == 0 || s == 1) {          // I, F
        const float* wc = prm + (s == 0 ? 0 : 1) * FC;
        const float* bb = prm + (s == 0 ? 3 : 4) * FC;
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = sigm(P[o] + wc[o] * cp[o] + bb[o]);
        store_row<FC>(gs + s * FC, P, true);
        return;
    }
    if (s == 2) {                    // T, then C' = F C + I T
        float I[FC], Fg[FC];
        load_row<FC>(I, gs, FC, true);            // written by this thread in slots 0 / 1
        load_row<FC>(Fg, gs + FC, FC, true);
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = ftanh(P[o] + prm[5 * FC + o]);
        store_row<FC>(gs + 2 * FC, P, true);
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = fmaf(Fg[o], cp[o], I[o] * P[o]);
        store_row<FC>(a.Craw + (size_t)i * FC, P, true);
        return;
    }
    // s == 3: O, H', norms, head
    float Cn[FC];
    load_row<FC>(Cn, a.Craw + (size_t)i * FC, FC, true);
#pragma unroll
    for (int o = 0; o < FC; ++o) P[o] = sigm(P[o] + prm[2 * FC + o] * Cn[o] + prm[6 * FC + o]);   // O
    store_row<FC>(gs + 3 * FC, P, true);
    if (a.Oout) store_row<FC>(a.Oout + (size_t)i * FC, P, true);
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
        store_row<FC>(a.Hout + (size_t)i * FC, Hh, true);
    }
    if (a.norm_c) {
        ln_stats(Cn, a.eps, mean, rstd);
#pragma unroll
        for (int o = 0; o < FC; ++o) Cn[o] = (Cn[o] - mean) * rstd * prm[9 * FC + o] + prm[10 * FC + o];
    }
    store_row<FC>(a.Cout + (size_t)i * FC, Cn, true);
    if (a.head_in) {
        if (a.norm_o) {
            ln_stats(P, a.eps, mean, rstd);
#pragma unroll
            for (int o = 0; o < FC; ++o) P[o] = (P[o] - mean) * rstd * prm[11 * FC + o] + prm[12 * FC + o];
        }
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = fmaxf(P[o], 0.f);
        float* hr = a.head_in + (size_t)i * a.ldh;
        store_row<FC>(hr, P, (a.ldh % 4 == 0));
        if (a.concat) hr[FC] = a.concat[i];
        for (int k = FC + 1; k < a.ldh; ++k) hr[k] = 0.f;      // pad columns of the 16-byte aligned head rows
    }
}

template <int DAC, int DBC>
__global__ void __launch_bounds__(128) fused_fwd_kernel(FusedFwdArgs a) {
    qmp_seed_init(a.seed, a.salt);
    extern __shared__ __align__(16) float sw[];
    constexpr int TA = (DAC > 0) ? ConvSizes<(DAC > 0 ? DAC : 4)>::TOTAL : 0;
    constexpr int TB = ConvSizes<DBC>::TOTAL;
    const int na = a.GA * TA, nb = a.GB * TB;
    float* swA = sw;
    float* swB = sw + na;
    float* prm = swB + nb;
    for (int idx = threadIdx.x * 4; idx < na; idx += 128 * 4)
        *reinterpret_cast<float4*>(swA + idx) = *reinterpret_cast<const float4*>(a.wa + idx);
    for (int idx = threadIdx.x * 4; idx < nb; idx += 128 * 4)
        *reinterpret_cast<float4*>(swB + idx) = *reinterpret_cast<const float4*>(a.wb + idx);
    if (a.mode == 1)
        for (int idx = threadIdx.x; idx < 13 * FC; idx += 128) prm[idx] = a.params[idx];
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= a.N) return;

    const int nslots = (a.mode == 1) ? 4 : a.NC;
    for (int s = 0; s < nslots; ++s) {
        float P[FC];
#pragma unroll
        for (int o = 0; o < FC; ++o) P[o] = 0.f;
        if (a.mode == 1) {
            if constexpr (DAC > 0) {
                if (a.GA) conv_accumulate<DAC>(a, i, s, a.xa, a.lda, a.DA, swA + s * TA, P);
            }
            conv_accumulate<DBC>(a, i, a.GA + s, a.xb + (a.sharedB ? 0 : s * a.DB), a.ldb, a.DB, swB + s * TB, P);
            if (a.GB == 8)
                conv_accumulate<DBC>(a, i, a.GA + 4 + s, a.xb + (4 + s) * a.DB, a.ldb, a.DB, swB + (4 + s) * TB, P);
            gate_epilogue(a, i, s, prm, P);
        } else {
            bool done = false;
            if constexpr (DAC > 0) {
                if (s < a.GA) {
                    conv_accumulate<DAC>(a, i, s, a.xa, a.lda, a.DA, swA + s * TA, P);
                    done = true;
                }
            }
            if (!done) {
                const int g = s - a.GA;
                conv_accumulate<DBC>(a, i, s, a.xb + (a.sharedB ? 0 : g * a.DB), a.ldb, a.DB, swB + g * TB, P);
            }
            float* orow = a.out + (size_t)i * a.ldo + (size_t)s * a.C;
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

template <int DAC, int DBC>
int launch_fwd(const FusedFwdArgs& a, cudaStream_t st) {
    constexpr int TA = (DAC > 0) ? ConvSizes<(DAC > 0 ? DAC : 4)>::TOTAL : 0;
    const size_t smem = sizeof(float) * ((size_t)a.GA * TA + (size_t)a.GB * ConvSizes<DBC>::TOTAL + 13 * FC);
    QMP_REQUIRE(smem <= 220 * 1024, "fused forward: weights do not fit in shared memory");
    auto kern = fused_fwd_kernel<DAC, DBC>;
    QMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(a.N, 128), 128, smem, st>>>(a);
    QMP_LAUNCH_CHECK("fused_fwd_kernel");
    return 0;
}

}  // namespace qmp
