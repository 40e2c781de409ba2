// Gate backward of the graph-conv LSTM cell for hidden size 32 in OCTET layout (8 lanes per node, one float4 of the 32 channels
// per lane, 4 nodes per warp instruction) -- shared by lstm_bwd_oct_kernel (lstm.cu) and by the prologue of the decoder-cell
// backward kernel (fused_cell_bwd.cu), where the gate epilogue's backward runs inside the message-passing kernel.
// Reference: autograd of GConvLSTM.forward (model/model.py:394-463) and of the norms / head input (model/seq2seq.py:59-66,
// 138-165).
#pragma once
#include "common.cuh"

namespace qmp {

enum { P_WCI = 0, P_WCF, P_WCO, P_BI, P_BF, P_BC, P_BO, P_GH, P_BH, P_GC, P_BCN, P_GO, P_BON, P_COUNT };

__device__ __forceinline__ float oct_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}
__device__ __forceinline__ void f4(float (&v)[4], const float4 q) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

// y = LN(x) * gamma + beta over the 32 channels of an octet: xhat of this lane's 4 channels and rstd
__device__ __forceinline__ float oct_ln_fwd(const float (&x)[4], float eps, float (&xh)[4]) {
    const float mean = oct_sum((x[0] + x[1]) + (x[2] + x[3])) * (1.f / 32.f);
    float d[4], v = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        d[k] = x[k] - mean;
        v = fmaf(d[k], d[k], v);
    }
    const float rstd = rsqrtf(oct_sum(v) * (1.f / 32.f) + eps);
#pragma unroll
    for (int k = 0; k < 4; ++k) xh[k] = d[k] * rstd;
    return rstd;
}
// dx given dy (in place), accumulating dgamma / dbeta
__device__ __forceinline__ void oct_ln_bwd(const float (&xh)[4], float (&dy)[4], const float (&gamma)[4], float rstd,
                                           float (&dgamma)[4], float (&dbeta)[4]) {
    float g[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dgamma[k] = fmaf(dy[k], xh[k], dgamma[k]);
        dbeta[k] += dy[k];
        g[k] = dy[k] * gamma[k];
        s1 += g[k];
        s2 = fmaf(g[k], xh[k], s2);
    }
    s1 = oct_sum(s1) * (1.f / 32.f);
    s2 = oct_sum(s2) * (1.f / 32.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) dy[k] = rstd * (g[k] - s1 - xh[k] * s2);
}


// One node's gate backward on this lane's 4 channels.  Inputs: saved gate activations I, F, T, O, raw C' (Cn), C (cp), the
// incoming gradients dH (w.r.t. LN_h(H') if norm_h), dC (w.r.t. LN_c(C') if norm_c), dO (direct), dhd (w.r.t. head_in[:, :32],
// used when has_head).  PRM(p, v) loads this lane's 4 channels of parameter row p.  Outputs: the four gate pre-activation
// gradients and dC_prev; dprm accumulates the 13 parameter rows.  dH / dC / dO / dhd are clobbered.
template <class PRM>
__device__ __forceinline__ void oct_gate_bwd(const float (&I)[4], const float (&F)[4], const float (&T)[4], const float (&O)[4],
                                             const float (&Cn)[4], const float (&cp)[4], float (&dH)[4], float (&dC)[4], float (&dO)[4],
                                             float (&dhd)[4], bool has_head, int norm_h, int norm_c, int norm_o, float eps, PRM PRM_,
                                             float (&dprm)[P_COUNT][4], float (&dI)[4], float (&dF)[4], float (&dT)[4], float (&dOp)[4],
                                             float (&dCp)[4]) {
#define PRM PRM_
    float tc[4], H[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        tc[k] = fast_tanh(Cn[k]);
        H[k] = O[k] * tc[k];
    }
    float xh[4];
    if (norm_h) {                          // dH arrives w.r.t. LN_h(H')
        float g[4];
        PRM(P_GH, g);
        const float rstd = oct_ln_fwd(H, eps, xh);
        oct_ln_bwd(xh, dH, g, rstd, dprm[P_GH], dprm[P_BH]);
    }
    if (norm_c) {
        float g[4];
        PRM(P_GC, g);
        const float rstd = oct_ln_fwd(Cn, eps, xh);
        oct_ln_bwd(xh, dC, g, rstd, dprm[P_GC], dprm[P_BCN]);
    }
    if (has_head) {                           // head_in[:, :C] = relu(LN_o(O))
        if (norm_o) {
            float g[4], b[4];
            PRM(P_GO, g);
            PRM(P_BON, b);
            const float rstd = oct_ln_fwd(O, eps, xh);
#pragma unroll
            for (int k = 0; k < 4; ++k) dhd[k] = (fmaf(xh[k], g[k], b[k]) > 0.f) ? dhd[k] : 0.f;
            oct_ln_bwd(xh, dhd, g, rstd, dprm[P_GO], dprm[P_BON]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) dhd[k] = (O[k] > 0.f) ? dhd[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dO[k] += dhd[k];
    }
    float wci[4], wcf[4], wco[4];
    PRM(P_WCI, wci);
    PRM(P_WCF, wcf);
    PRM(P_WCO, wco);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float dOt = fmaf(dH[k], tc[k], dO[k]);
        dOp[k] = dOt * O[k] * (1.f - O[k]);
        const float dCn = dC[k] + dH[k] * O[k] * (1.f - tc[k] * tc[k]) + dOp[k] * wco[k];
        dI[k] = dCn * T[k] * I[k] * (1.f - I[k]);
        dF[k] = dCn * cp[k] * F[k] * (1.f - F[k]);
        dT[k] = dCn * I[k] * (1.f - T[k] * T[k]);
        dCp[k] = dCn * F[k] + dI[k] * wci[k] + dF[k] * wcf[k];
        dprm[P_WCI][k] = fmaf(dI[k], cp[k], dprm[P_WCI][k]);
        dprm[P_WCF][k] = fmaf(dF[k], cp[k], dprm[P_WCF][k]);
        dprm[P_WCO][k] = fmaf(dOp[k], Cn[k], dprm[P_WCO][k]);
        dprm[P_BI][k] += dI[k];
        dprm[P_BF][k] += dF[k];
        dprm[P_BC][k] += dT[k];
        dprm[P_BO][k] += dOp[k];
    }
#undef PRM
}

}  // namespace qmp
