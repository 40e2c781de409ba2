// Gate epilogue of the graph-conv LSTM cell, forward and backward, fused with the LayerNorms the
// encoder / decoder apply right after the cell and with the decoder's head input.
//
// Reference: GConvLSTM.forward, model/model.py:394-463 (peephole LSTM); Encoder/Decoder norm_h, norm_c,
// norm_o + relu + concat, model/seq2seq.py:59-66, 138-165.  With P = conv_x_*(X) + conv_h_*(H) [N, 4C]
// (gate order i, f, c, o):
//     I = sig(P_i + w_ci*C + b_i)   F = sig(P_f + w_cf*C + b_f)   T = tanh(P_c + b_c)
//     C' = F*C + I*T                O = sig(P_o + w_co*C' + b_o)  H' = O*tanh(C')
//     H_out = LN_h(H'), C_out = LN_c(C') (optional), head_in = [relu(LN_o(O)), concat] (optional)
// One warp per node, lane = channel (C <= 128: up to 4 channels per lane), so every LayerNorm
// statistic is a warp reduction and no intermediate leaves registers.
#include "common.cuh"
#include "lstm_oct.cuh"

namespace qmp {


struct LstmArgs {
    int N, C;
    const float* P; int ldp;          // [N, 4C]
    const float* Cprev;               // [N, C] or NULL (zeros)
    const float* params;              // [P_COUNT, C]
    int norm_h, norm_c, norm_o;       // which LayerNorms are applied
    float eps;
    float* gates;                     // [N, 4C] saved I, F, T, O
    float* Craw;                      // [N, C]  C' before the norm
    float* Oout;                      // [N, C] raw O (optional)
    float* Hout;                      // [N, C]
    float* Cout;                      // [N, C]
    float* head_in; int ldh;          // [N, C+1] optional
    const float* concat;              // [N] optional (column C of head_in)
    // backward
    const float* dHout; const float* dCout; const float* dOdirect;
    const float* dHead; int lddh;     // [N, >= C] optional: gradient of head_in[:, :C]
    float* dP; int lddp;              // [N, 4C]
    float* dCprev;                    // [N, C] optional
    float* dparams;                   // [P_COUNT, C] accumulated with atomics
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int QC>
__device__ __forceinline__ void layer_norm_fwd(const float (&x)[QC], int C, int lane, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < QC; ++q) s += (lane + 32 * q < C) ? x[q] : 0.f;
    mean = warp_sum(s) / (float)C;
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < QC; ++q) {
        const float d = (lane + 32 * q < C) ? x[q] - mean : 0.f;
        v = fmaf(d, d, v);
    }
    rstd = rsqrtf(warp_sum(v) / (float)C + eps);
}

// dx for y = LN(x) * gamma + beta given dy; xhat = (x - mean) * rstd.  Also accumulates dgamma / dbeta.
template <int QC>
__device__ __forceinline__ void layer_norm_bwd(const float (&xhat)[QC], const float (&dy)[QC], const float (&gamma)[QC],
                                               int C, int lane, float rstd, float (&dx)[QC]) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < QC; ++q) {
        const bool ok = lane + 32 * q < C;
        const float g = ok ? dy[q] * gamma[q] : 0.f;
        s1 += g;
        s2 = fmaf(g, ok ? xhat[q] : 0.f, s2);
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int q = 0; q < QC; ++q) dx[q] = rstd * (dy[q] * gamma[q] - s1 - xhat[q] * s2);
}

template <int QC>
__global__ void __launch_bounds__(256) lstm_fwd_kernel(LstmArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int C = a.C;
    float prm[P_COUNT][QC];
#pragma unroll
    for (int p = 0; p < P_COUNT; ++p)
#pragma unroll
        for (int q = 0; q < QC; ++q) prm[p][q] = (lane + 32 * q < C) ? a.params[p * C + lane + 32 * q] : 0.f;
    for (int i = warp; i < a.N; i += nwarps) {
        float I[QC], F[QC], T[QC], O[QC], Cn[QC], H[QC];
        const float* pr = a.P + (size_t)i * a.ldp;
#pragma unroll
        for (int q = 0; q < QC; ++q) {
            const int c = lane + 32 * q;
            const bool ok = c < C;
            const float cp = (ok && a.Cprev) ? a.Cprev[(size_t)i * C + c] : 0.f;
            const float pi = ok ? pr[c] : 0.f, pf = ok ? pr[C + c] : 0.f, pc = ok ? pr[2 * C + c] : 0.f,
                        po = ok ? pr[3 * C + c] : 0.f;
            I[q] = sigmoidf_(pi + prm[P_WCI][q] * cp + prm[P_BI][q]);
            F[q] = sigmoidf_(pf + prm[P_WCF][q] * cp + prm[P_BF][q]);
            T[q] = tanhf(pc + prm[P_BC][q]);
            Cn[q] = F[q] * cp + I[q] * T[q];
            O[q] = sigmoidf_(po + prm[P_WCO][q] * Cn[q] + prm[P_BO][q]);
            H[q] = O[q] * tanhf(Cn[q]);
            if (ok) {
                float* gs = a.gates + (size_t)i * 4 * C;
                gs[c] = I[q]; gs[C + c] = F[q]; gs[2 * C + c] = T[q]; gs[3 * C + c] = O[q];
                a.Craw[(size_t)i * C + c] = Cn[q];
                if (a.Oout) a.Oout[(size_t)i * C + c] = O[q];
            }
        }
        float mean, rstd;
        if (a.norm_h) {
            layer_norm_fwd<QC>(H, C, lane, a.eps, mean, rstd);
#pragma unroll
            for (int q = 0; q < QC; ++q) H[q] = (H[q] - mean) * rstd * prm[P_GH][q] + prm[P_BH][q];
        }
        if (a.norm_c) {
            layer_norm_fwd<QC>(Cn, C, lane, a.eps, mean, rstd);
#pragma unroll
            for (int q = 0; q < QC; ++q) Cn[q] = (Cn[q] - mean) * rstd * prm[P_GC][q] + prm[P_BCN][q];
        }
#pragma unroll
        for (int q = 0; q < QC; ++q) {
            const int c = lane + 32 * q;
            if (c < C) {
                a.Hout[(size_t)i * C + c] = H[q];
                a.Cout[(size_t)i * C + c] = Cn[q];
            }
        }
        if (a.head_in) {
            if (a.norm_o) {
                layer_norm_fwd<QC>(O, C, lane, a.eps, mean, rstd);
#pragma unroll
                for (int q = 0; q < QC; ++q) O[q] = (O[q] - mean) * rstd * prm[P_GO][q] + prm[P_BON][q];
            }
            float* hr = a.head_in + (size_t)i * a.ldh;
#pragma unroll
            for (int q = 0; q < QC; ++q) {
                const int c = lane + 32 * q;
                if (c < C) hr[c] = fmaxf(O[q], 0.f);
            }
            if (lane == 0 && a.concat) hr[C] = a.concat[i];
        }
    }
}

template <int QC>
__global__ void __launch_bounds__(256) lstm_bwd_kernel(LstmArgs a) {
    __shared__ float s_dp[P_COUNT * 128];
    for (int t = threadIdx.x; t < P_COUNT * 128; t += blockDim.x) s_dp[t] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int C = a.C;
    float prm[P_COUNT][QC], dprm[P_COUNT][QC];
#pragma unroll
    for (int p = 0; p < P_COUNT; ++p)
#pragma unroll
        for (int q = 0; q < QC; ++q) {
            prm[p][q] = (lane + 32 * q < C) ? a.params[p * C + lane + 32 * q] : 0.f;
            dprm[p][q] = 0.f;
        }
    for (int i = warp; i < a.N; i += nwarps) {
        float I[QC], F[QC], T[QC], O[QC], Cn[QC], cp[QC], tc[QC], H[QC];
        float dH[QC], dC[QC], dO[QC];
        const float* gs = a.gates + (size_t)i * 4 * C;
#pragma unroll
        for (int q = 0; q < QC; ++q) {
            const int c = lane + 32 * q;
            const bool ok = c < C;
            I[q] = ok ? gs[c] : 0.f; F[q] = ok ? gs[C + c] : 0.f; T[q] = ok ? gs[2 * C + c] : 0.f;
            O[q] = ok ? gs[3 * C + c] : 0.f;
            Cn[q] = ok ? a.Craw[(size_t)i * C + c] : 0.f;
            cp[q] = (ok && a.Cprev) ? a.Cprev[(size_t)i * C + c] : 0.f;
            tc[q] = tanhf(Cn[q]);
            H[q] = O[q] * tc[q];
            dH[q] = (ok && a.dHout) ? a.dHout[(size_t)i * C + c] : 0.f;
            dC[q] = (ok && a.dCout) ? a.dCout[(size_t)i * C + c] : 0.f;
            dO[q] = (ok && a.dOdirect) ? a.dOdirect[(size_t)i * C + c] : 0.f;
        }
        float mean, rstd;
        if (a.norm_h) {  // dH currently w.r.t. LN_h(H')
            float xh[QC], dx[QC];
            layer_norm_fwd<QC>(H, C, lane, a.eps, mean, rstd);
#pragma unroll
            for (int q = 0; q < QC; ++q) {
                xh[q] = (H[q] - mean) * rstd;
                dprm[P_GH][q] = fmaf(dH[q], xh[q], dprm[P_GH][q]);
                dprm[P_BH][q] += dH[q];
            }
            layer_norm_bwd<QC>(xh, dH, prm[P_GH], C, lane, rstd, dx);
#pragma unroll
            for (int q = 0; q < QC; ++q) dH[q] = dx[q];
        }
        if (a.norm_c) {
            float xh[QC], dx[QC];
            layer_norm_fwd<QC>(Cn, C, lane, a.eps, mean, rstd);
#pragma unroll
            for (int q = 0; q < QC; ++q) {
                xh[q] = (Cn[q] - mean) * rstd;
                dprm[P_GC][q] = fmaf(dC[q], xh[q], dprm[P_GC][q]);
                dprm[P_BCN][q] += dC[q];
            }
            layer_norm_bwd<QC>(xh, dC, prm[P_GC], C, lane, rstd, dx);
#pragma unroll
            for (int q = 0; q < QC; ++q) dC[q] = dx[q];
        }
        if (a.dHead) {  // head_in[:, :C] = relu(LN_o(O))
            float dy[QC];
            const float* dh = a.dHead + (size_t)i * a.lddh;
            if (a.norm_o) {
                float xh[QC], dx[QC];
                layer_norm_fwd<QC>(O, C, lane, a.eps, mean, rstd);
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    const int c = lane + 32 * q;
                    xh[q] = (O[q] - mean) * rstd;
                    const float y = xh[q] * prm[P_GO][q] + prm[P_BON][q];
                    dy[q] = (c < C && y > 0.f) ? dh[c] : 0.f;
                    dprm[P_GO][q] = fmaf(dy[q], xh[q], dprm[P_GO][q]);
                    dprm[P_BON][q] += dy[q];
                }
                layer_norm_bwd<QC>(xh, dy, prm[P_GO], C, lane, rstd, dx);
#pragma unroll
                for (int q = 0; q < QC; ++q) dO[q] += dx[q];
            } else {
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    const int c = lane + 32 * q;
                    dO[q] += (c < C && O[q] > 0.f) ? dh[c] : 0.f;
                }
            }
        }
        float* dpr = a.dP + (size_t)i * a.lddp;
#pragma unroll
        for (int q = 0; q < QC; ++q) {
            const int c = lane + 32 * q;
            const float dOt = dH[q] * tc[q] + dO[q];
            const float dOp = dOt * O[q] * (1.f - O[q]);
            const float dCn = dC[q] + dH[q] * O[q] * (1.f - tc[q] * tc[q]) + dOp * prm[P_WCO][q];
            const float dIp = dCn * T[q] * I[q] * (1.f - I[q]);
            const float dFp = dCn * cp[q] * F[q] * (1.f - F[q]);
            const float dTp = dCn * I[q] * (1.f - T[q] * T[q]);
            if (c < C) {
                dpr[c] = dIp; dpr[C + c] = dFp; dpr[2 * C + c] = dTp; dpr[3 * C + c] = dOp;
                if (a.dCprev) a.dCprev[(size_t)i * C + c] = dCn * F[q] + dIp * prm[P_WCI][q] + dFp * prm[P_WCF][q];
            }
            dprm[P_WCI][q] = fmaf(dIp, cp[q], dprm[P_WCI][q]);
            dprm[P_WCF][q] = fmaf(dFp, cp[q], dprm[P_WCF][q]);
            dprm[P_WCO][q] = fmaf(dOp, Cn[q], dprm[P_WCO][q]);
            dprm[P_BI][q] += dIp; dprm[P_BF][q] += dFp; dprm[P_BC][q] += dTp; dprm[P_BO][q] += dOp;
        }
    }
#pragma unroll
    for (int p = 0; p < P_COUNT; ++p)
#pragma unroll
        for (int q = 0; q < QC; ++q)
            if (lane + 32 * q < C) atomicAdd(&s_dp[p * 128 + lane + 32 * q], dprm[p][q]);
    __syncthreads();
    if (a.dparams)
        for (int t = threadIdx.x; t < P_COUNT * C; t += blockDim.x) {
            const float v = s_dp[(t / C) * 128 + t % C];
            if (v != 0.f) atomicAdd(&a.dparams[t], v);
        }
}

__global__ void __launch_bounds__(256, 2) lstm_bwd_oct_kernel(LstmArgs a) {
    __shared__ float s_dp[P_COUNT * 32];
    __shared__ __align__(16) float s_prm[P_COUNT * 32];
    for (int t = threadIdx.x; t < P_COUNT * 32; t += blockDim.x) {
        s_dp[t] = 0.f;
        s_prm[t] = a.params[t];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, o8 = lane >> 3, l8 = lane & 7;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    float dprm[P_COUNT][4];
#pragma unroll
    for (int p = 0; p < P_COUNT; ++p)
#pragma unroll
        for (int k = 0; k < 4; ++k) dprm[p][k] = 0.f;
    auto PRM = [&](int p, float (&v)[4]) { f4(v, *reinterpret_cast<const float4*>(s_prm + p * 32 + 4 * l8)); };
    const int npass = (a.N + 3) / 4;
    for (int ps = warp; ps < npass; ps += nwarps) {
        const int i = 4 * ps + o8;
        const bool valid = i < a.N;
        const size_t r32 = (size_t)i * 32 + 4 * l8;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        float I[4], F[4], T[4], O[4], Cn[4], cp[4], dH[4], dC[4], dO[4], dhd[4];
        {
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
        }
        float dI[4], dF[4], dT[4], dOp[4], dCp[4];
        oct_gate_bwd(I, F, T, O, Cn, cp, dH, dC, dO, dhd, a.dHead != nullptr, a.norm_h, a.norm_c, a.norm_o, a.eps, PRM, dprm, dI, dF, dT,
                     dOp, dCp);
        if (valid) {
            float* dpr = a.dP + (size_t)i * a.lddp + 4 * l8;
            *reinterpret_cast<float4*>(dpr) = make_float4(dI[0], dI[1], dI[2], dI[3]);
            *reinterpret_cast<float4*>(dpr + 32) = make_float4(dF[0], dF[1], dF[2], dF[3]);
            *reinterpret_cast<float4*>(dpr + 64) = make_float4(dT[0], dT[1], dT[2], dT[3]);
            *reinterpret_cast<float4*>(dpr + 96) = make_float4(dOp[0], dOp[1], dOp[2], dOp[3]);
            if (a.dCprev) *reinterpret_cast<float4*>(a.dCprev + r32) = make_float4(dCp[0], dCp[1], dCp[2], dCp[3]);
        }
    }
#pragma unroll
    for (int p = 0; p < P_COUNT; ++p)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = dprm[p][k];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (o8 == 0) atomicAdd(&s_dp[p * 32 + 4 * l8 + k], v);
        }
    __syncthreads();
    if (a.dparams)
        for (int t = threadIdx.x; t < P_COUNT * 32; t += blockDim.x) {
            const float v = s_dp[t];
            if (v != 0.f) atomicAdd(&a.dparams[t], v);
        }
}

// decoder head tail (model/seq2seq.py:167-178, 427-428): out = tanh(drop(y)) + x0 [-> sigmoid];
// x_next = [out, x[:, 1:]]
__global__ void head_finish_fwd_kernel(const float* __restrict__ y, const float* __restrict__ x, int N, int F, int binary,
                                       float drop_p, unsigned long long seed, const unsigned long long* __restrict__ salt,
                                       float* __restrict__ out, float* __restrict__ x_next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float keep = 1.f;
    if (drop_p > 0.f) {
        seed = salted_seed(seed, salt);
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        keep = ((float)(z >> 40) * (1.0f / 16777216.0f) >= drop_p) ? 1.f / (1.f - drop_p) : 0.f;
    }
    float o = tanhf(y[i] * keep) + x[(size_t)i * F];
    if (binary) o = 1.f / (1.f + expf(-o));
    out[i] = o;
    if (x_next) {
        x_next[(size_t)i * F] = o;
        for (int c = 1; c < F; ++c) x_next[(size_t)i * F + c] = x[(size_t)i * F + c];
    }
}

// dy and the whole gradient row dx [N, F] of the step's input x: column 0 from d_out (+ d_xnext[:, 0]), columns 1.. = d_xnext
// (zeros without it); out is the forward result
__global__ void head_finish_bwd_kernel(const float* __restrict__ y, const float* __restrict__ out, const float* __restrict__ x,
                                       const float* __restrict__ d_out, const float* __restrict__ d_xnext, int N, int F,
                                       int binary, float drop_p, unsigned long long seed,
                                       const unsigned long long* __restrict__ salt, float* __restrict__ dy,
                                       float* __restrict__ dx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float keep = 1.f;
    if (drop_p > 0.f) {
        seed = salted_seed(seed, salt);
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        keep = ((float)(z >> 40) * (1.0f / 16777216.0f) >= drop_p) ? 1.f / (1.f - drop_p) : 0.f;
    }
    float g = (d_out ? d_out[i] : 0.f) + (d_xnext ? d_xnext[(size_t)i * F] : 0.f);
    if (binary) g *= out[i] * (1.f - out[i]);
    const float th = tanhf(y[i] * keep);
    dy[i] = g * (1.f - th * th) * keep;
    dx[(size_t)i * F] = g;
    for (int f = 1; f < F; ++f) dx[(size_t)i * F + f] = d_xnext ? d_xnext[(size_t)i * F + f] : 0.f;
}

__global__ void relu_mask_kernel(const float* __restrict__ y, float* __restrict__ dy, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(y[i] > 0.f)) dy[i] = 0.f;
}

// out-of-place form: out = g where y > 0, else 0 (16-byte vectors when n and the pointers allow it)
__global__ void relu_mask_to_kernel(const float* __restrict__ y, const float* __restrict__ g, float* __restrict__ out, long long n,
                                    int vec) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        if (i * 4 >= n) return;
        const float4 a = __ldg(reinterpret_cast<const float4*>(y) + i), b = __ldg(reinterpret_cast<const float4*>(g) + i);
        reinterpret_cast<float4*>(out)[i] = make_float4(a.x > 0.f ? b.x : 0.f, a.y > 0.f ? b.y : 0.f, a.z > 0.f ? b.z : 0.f,
                                                        a.w > 0.f ? b.w : 0.f);
    } else if (i < n) {
        out[i] = y[i] > 0.f ? g[i] : 0.f;
    }
}

}  // namespace qmp
using namespace qmp;

static int lstm_grid(int N) {
    const int want = (N + 7) / 8;  // 8 warps per block
    return want < 148 * 4 ? (want > 0 ? want : 1) : 148 * 4;
}

#define QMP_DISPATCH_QC(C, CALL)                                             \
    do {                                                                     \
        if ((C) <= 32) { CALL(1); }                                          \
        else if ((C) <= 64) { CALL(2); }                                     \
        else if ((C) <= 128) { CALL(4); }                                    \
        else { qmp::set_error("lstm gates: hidden size %d > 128 unsupported", (C)); return -1; } \
    } while (0)

// params: [13, C] = w_c_i, w_c_f, w_c_o, b_i, b_f, b_c, b_o, norm_h.{weight,bias}, norm_c.{weight,bias},
// norm_o.{weight,bias}.  P [N, 4C] (ld = ldp).  Outputs: gates [N,4C], Craw/Hout/Cout [N,C];
// Oout [N,C], head_in [N, ldh] (needs concat [N] for column C) optional.
QMP_API int qmp_lstm_gates_fwd(int N, int C, const float* P, int ldp, const float* Cprev, const float* params, int norm_h,
                               int norm_c, int norm_o, float eps, float* gates, float* Craw, float* Oout, float* Hout,
                               float* Cout, float* head_in, int ldh, const float* concat, void* stream) {
    if (N <= 0) return 0;
    LstmArgs a{};
    a.N = N; a.C = C; a.P = P; a.ldp = ldp; a.Cprev = Cprev; a.params = params; a.norm_h = norm_h; a.norm_c = norm_c;
    a.norm_o = norm_o; a.eps = eps; a.gates = gates; a.Craw = Craw; a.Oout = Oout; a.Hout = Hout; a.Cout = Cout;
    a.head_in = head_in; a.ldh = ldh; a.concat = concat;
#define CALL(QQ) lstm_fwd_kernel<QQ><<<lstm_grid(N), 256, 0, (cudaStream_t)stream>>>(a)
    QMP_DISPATCH_QC(C, CALL);
#undef CALL
    QMP_LAUNCH_CHECK("qmp_lstm_gates_fwd");
    return 0;
}

// dHout / dCout / dOdirect [N,C] and dHead [N, lddh] may each be NULL.  Writes dP [N, 4C] (ld = lddp) and
// dCprev [N,C] (optional); ACCUMULATES parameter gradients into dparams [13, C].
QMP_API int qmp_lstm_gates_bwd(int N, int C, const float* gates, const float* Craw, const float* Cprev,
                               const float* params, int norm_h, int norm_c, int norm_o, float eps, const float* dHout,
                               const float* dCout, const float* dOdirect, const float* dHead, int lddh, float* dP,
                               int lddp, float* dCprev, float* dparams, void* stream) {
    if (N <= 0) return 0;
    LstmArgs a{};
    a.N = N; a.C = C; a.gates = const_cast<float*>(gates); a.Craw = const_cast<float*>(Craw); a.Cprev = Cprev;
    a.params = params; a.norm_h = norm_h; a.norm_c = norm_c; a.norm_o = norm_o; a.eps = eps; a.dHout = dHout;
    a.dCout = dCout; a.dOdirect = dOdirect; a.dHead = dHead; a.lddh = lddh; a.dP = dP; a.lddp = lddp;
    a.dCprev = dCprev; a.dparams = dparams;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (C == 32 && lddp % 4 == 0 && (!dHead || lddh % 4 == 0) && al16(gates) && al16(Craw) && al16(Cprev) && al16(params) &&
        al16(dHout) && al16(dCout) && al16(dOdirect) && al16(dHead) && al16(dP) && al16(dCprev)) {
        const int want = (N + 31) / 32;          // 8 warps x 4 nodes per pass
        lstm_bwd_oct_kernel<<<want < 148 * 2 ? want : 148 * 2, 256, 0, (cudaStream_t)stream>>>(a);
        QMP_LAUNCH_CHECK("qmp_lstm_gates_bwd");
        return 0;
    }
#define CALL(QQ) lstm_bwd_kernel<QQ><<<lstm_grid(N), 256, 0, (cudaStream_t)stream>>>(a)
    QMP_DISPATCH_QC(C, CALL);
#undef CALL
    QMP_LAUNCH_CHECK("qmp_lstm_gates_bwd");
    return 0;
}

QMP_API int qmp_head_finish_fwd(const float* y, const float* x, int N, int F, int binary, float drop_p,
                                unsigned long long seed, float* out, float* x_next, void* stream) {
    if (N <= 0) return 0;
    head_finish_fwd_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(y, x, N, F, binary, drop_p, seed, qmp::dropout_salt(), out, x_next);
    QMP_LAUNCH_CHECK("qmp_head_finish_fwd");
    return 0;
}

QMP_API int qmp_head_finish_bwd(const float* y, const float* out, const float* x, const float* d_out,
                                const float* d_xnext, int N, int F, int binary, float drop_p, unsigned long long seed,
                                float* dy, float* dx, void* stream) {
    if (N <= 0) return 0;
    head_finish_bwd_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(y, out, x, d_out, d_xnext, N, F, binary,
                                                                           drop_p, seed, qmp::dropout_salt(), dy, dx);
    QMP_LAUNCH_CHECK("qmp_head_finish_bwd");
    return 0;
}

// dy[i] = 0 where y[i] <= 0 (backward of relu applied to y)
QMP_API int qmp_relu_mask(const float* y, float* dy, long long n, void* stream) {
    if (n <= 0) return 0;
    relu_mask_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(y, dy, n);
    QMP_LAUNCH_CHECK("qmp_relu_mask");
    return 0;
}

// out[i] = g[i] where y[i] > 0, else 0: the backward of relu without the clone an in-place mask needs (autograd owns g)
QMP_API int qmp_relu_mask_to(const float* y, const float* g, float* out, long long n, void* stream) {
    if (n <= 0) return 0;
    const int vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const long long work = vec ? n / 4 : n;
    relu_mask_to_kernel<<<(unsigned)cdiv(work, 256), 256, 0, (cudaStream_t)stream>>>(y, g, out, n, vec);
    QMP_LAUNCH_CHECK("qmp_relu_mask_to");
    return 0;
}


// ---- GConvGRU gate arithmetic (model/model.py:236-259) as two element-wise stages around conv_h_h, which reads H * R ----------
namespace qmp {
__device__ __forceinline__ float gru_sig(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void gru_gates1_fwd_kernel(long long n, const float* __restrict__ az, const float* __restrict__ bz, const float* __restrict__ ar,
                                      const float* __restrict__ br, const float* __restrict__ H, float* __restrict__ Z,
                                      float* __restrict__ R, float* __restrict__ HR) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float z = gru_sig(az[i] + bz[i]), r = gru_sig(ar[i] + br[i]);
    Z[i] = z;
    R[i] = r;
    HR[i] = H[i] * r;
}
__global__ void gru_gates1_bwd_kernel(long long n, const float* __restrict__ Z, const float* __restrict__ R, const float* __restrict__ H,
                                      const float* __restrict__ dZ, const float* __restrict__ dR, const float* __restrict__ dHR,
                                      float* __restrict__ dpz, float* __restrict__ dpr, float* __restrict__ dH) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float z = Z[i], r = R[i], ghr = dHR ? dHR[i] : 0.f;
    dpz[i] = (dZ ? dZ[i] : 0.f) * z * (1.f - z);
    dpr[i] = ((dR ? dR[i] : 0.f) + ghr * H[i]) * r * (1.f - r);
    dH[i] = ghr * r;
}
__global__ void gru_gates2_fwd_kernel(long long n, const float* __restrict__ ah, const float* __restrict__ bh, const float* __restrict__ Z,
                                      const float* __restrict__ H, float* __restrict__ Ht, float* __restrict__ Hn) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t = tanhf(ah[i] + bh[i]), z = Z[i];
    Ht[i] = t;
    Hn[i] = z * H[i] + (1.f - z) * t;
}
__global__ void gru_gates2_bwd_kernel(long long n, const float* __restrict__ Z, const float* __restrict__ H, const float* __restrict__ Ht,
                                      const float* __restrict__ dHn, float* __restrict__ dph, float* __restrict__ dZ, float* __restrict__ dH) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = dHn[i], z = Z[i], t = Ht[i];
    dZ[i] = g * (H[i] - t);
    dH[i] = g * z;
    dph[i] = g * (1.f - z) * (1.f - t * t);
}
}  // namespace qmp

// GConvGRU, first gate stage on n = N * C elements: Z = sigmoid(az + bz), R = sigmoid(ar + br), HR = H * R (model/model.py:240-250;
// az / bz / ar / br are the outputs of conv_x_z / conv_h_z / conv_x_r / conv_h_r).
QMP_API int qmp_gru_gates1_fwd(long long n, const float* az, const float* bz, const float* ar, const float* br, const float* H, float* Z,
                               float* R, float* HR, void* stream) {
    if (n <= 0) return 0;
    gru_gates1_fwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, az, bz, ar, br, H, Z, R, HR);
    QMP_LAUNCH_CHECK("qmp_gru_gates1_fwd");
    return 0;
}

// Backward of qmp_gru_gates1_fwd: dZ / dR / dHR may be NULL (zero).  dpz = d az = d bz, dpr = d ar = d br, dH = dHR * R.
QMP_API int qmp_gru_gates1_bwd(long long n, const float* Z, const float* R, const float* H, const float* dZ, const float* dR,
                               const float* dHR, float* dpz, float* dpr, float* dH, void* stream) {
    if (n <= 0) return 0;
    gru_gates1_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, Z, R, H, dZ, dR, dHR, dpz, dpr, dH);
    QMP_LAUNCH_CHECK("qmp_gru_gates1_bwd");
    return 0;
}

// GConvGRU, second gate stage: Ht = tanh(ah + bh), Hn = Z * H + (1 - Z) * Ht (model/model.py:251-258); Ht is saved.
QMP_API int qmp_gru_gates2_fwd(long long n, const float* ah, const float* bh, const float* Z, const float* H, float* Ht, float* Hn,
                               void* stream) {
    if (n <= 0) return 0;
    gru_gates2_fwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, ah, bh, Z, H, Ht, Hn);
    QMP_LAUNCH_CHECK("qmp_gru_gates2_fwd");
    return 0;
}

// Backward of qmp_gru_gates2_fwd: dph = d ah = d bh, dZ, dH (the Z * H term).
QMP_API int qmp_gru_gates2_bwd(long long n, const float* Z, const float* H, const float* Ht, const float* dHn, float* dph, float* dZ,
                               float* dH, void* stream) {
    if (n <= 0) return 0;
    gru_gates2_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, Z, H, Ht, dHn, dph, dZ, dH);
    QMP_LAUNCH_CHECK("qmp_gru_gates2_bwd");
    return 0;
}
