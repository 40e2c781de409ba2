// Edge phase of PyG 2.2.0 GATConv / GATv2Conv with one head (the reference's CONVOLUTION_KWARGS, model/model.py:55-56:
// heads = 1, edge_dim = 2): segment softmax of additive attention logits over the in-edges and the weighted aggregate of the
// source rows.  The linear maps around it (lin_src / lin_l / lin_r, the att dot products, bias) are node GEMMs (gemm.cu).
//     mode 1 (GATConv):    logit_e = lrelu(as[j] + ad[i] + we . ea_e)                          as / ad: per-node scalars
//     mode 2 (GATv2Conv):  logit_e = sum_c att_c lrelu(XL[j,c] + XR[i,c] + We[c,:] . ea_e)
//     alpha = softmax over the in-edges of i (PyG utils.softmax: max-shifted, sum + 1e-16);  out_i = sum_e alpha_e XL[j_e]
// One thread per target node over the in-CSR; no config of the hot path selects these convs, so the kernels
// are written for exactness, not speed: fp32 SIMT, source-side gradients by atomics.
#include "common.cuh"

namespace qmp {

struct GatArgs {
    int N, C, mode;
    const int* ptr; const int* nbr; const float* ea;          // in-CSR, edge attributes [E, 2] in CSR order
    const float* XL; int ldl;                                 // [N, C] source rows
    const float* as; const float* ad; const float* we;        // mode 1: [N], [N], [2]
    const float* XR; int ldr; const float* We; const float* att;      // mode 2: [N, C], [C, 2], [C]
    float slope;
    float* out; int ldo;                                      // [N, C]
    float* alpha;                                             // [E] saved attention coefficients
    // backward
    const float* dOut; int lddo;
    float* dlog;                                              // [E] scratch: d logit_e
    float* dXL; float* das; float* dad; float* dwe;           // dXL [N, C] and das [N] zero on entry (atomics); dad [N]; dwe [2] accumulated
    float* dXR; float* dWe; float* datt;                      // mode 2: dXR [N, C] written; dWe [C, 2], datt [C] accumulated
};

__device__ __forceinline__ float gat_lrelu(float v, float s) { return v > 0.f ? v : s * v; }

__device__ __forceinline__ float gat_logit(const GatArgs& a, int i, int j, int kk) {
    const float e0 = a.ea ? __ldg(a.ea + 2 * (size_t)kk) : 0.f, e1 = a.ea ? __ldg(a.ea + 2 * (size_t)kk + 1) : 0.f;
    if (a.mode == 1) return gat_lrelu(__ldg(a.as + j) + __ldg(a.ad + i) + a.we[0] * e0 + a.we[1] * e1, a.slope);
    float s = 0.f;
    for (int c = 0; c < a.C; ++c) {
        const float m = __ldg(a.XL + (size_t)j * a.ldl + c) + __ldg(a.XR + (size_t)i * a.ldr + c) + __ldg(a.We + 2 * c) * e0 + __ldg(a.We + 2 * c + 1) * e1;
        s = fmaf(__ldg(a.att + c), gat_lrelu(m, a.slope), s);
    }
    return s;
}

__global__ void __launch_bounds__(128) gat_fwd_kernel(const GatArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const int k0 = __ldg(a.ptr + i), k1 = __ldg(a.ptr + i + 1);
    float mx = -INFINITY;
    for (int kk = k0; kk < k1; ++kk) {
        const float l = gat_logit(a, i, __ldg(a.nbr + kk), kk);
        a.alpha[kk] = l;
        mx = fmaxf(mx, l);
    }
    float den = 0.f;
    for (int kk = k0; kk < k1; ++kk) den += expf(a.alpha[kk] - mx);
    den += 1e-16f;
    for (int c = 0; c < a.C; ++c) a.out[(size_t)i * a.ldo + c] = 0.f;
    for (int kk = k0; kk < k1; ++kk) {
        const float al = expf(a.alpha[kk] - mx) / den;
        a.alpha[kk] = al;
        const float* xr = a.XL + (size_t)__ldg(a.nbr + kk) * a.ldl;
        for (int c = 0; c < a.C; ++c) a.out[(size_t)i * a.ldo + c] = fmaf(al, __ldg(xr + c), a.out[(size_t)i * a.ldo + c]);
    }
}

__device__ __forceinline__ float gat_warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(128) gat_bwd_kernel(const GatArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.N;
    const int k0 = valid ? __ldg(a.ptr + i) : 0, k1 = valid ? __ldg(a.ptr + i + 1) : 0;
    const float* g = a.dOut + (size_t)(valid ? i : 0) * a.lddo;
    // d alpha_e = dOut_i . XL[j];  t = sum alpha d alpha;  d logit_e = alpha (d alpha - t);  dXL[j] += alpha dOut_i
    float t = 0.f;
    for (int kk = k0; kk < k1; ++kk) {
        const float* xr = a.XL + (size_t)__ldg(a.nbr + kk) * a.ldl;
        float d = 0.f;
        for (int c = 0; c < a.C; ++c) d = fmaf(__ldg(g + c), __ldg(xr + c), d);
        a.dlog[kk] = d;
        t = fmaf(a.alpha[kk], d, t);
    }
    float dad = 0.f, dw0 = 0.f, dw1 = 0.f;
    for (int kk = k0; kk < k1; ++kk) {
        const int j = __ldg(a.nbr + kk);
        const float al = a.alpha[kk];
        const float dl = al * (a.dlog[kk] - t);
        a.dlog[kk] = dl;
        for (int c = 0; c < a.C; ++c) atomicAdd(a.dXL + (size_t)j * a.ldl + c, al * __ldg(g + c));
        if (a.mode == 1) {
            const float e0 = a.ea ? __ldg(a.ea + 2 * (size_t)kk) : 0.f, e1 = a.ea ? __ldg(a.ea + 2 * (size_t)kk + 1) : 0.f;
            const float pre = __ldg(a.as + j) + __ldg(a.ad + i) + a.we[0] * e0 + a.we[1] * e1;
            const float gg = dl * (pre > 0.f ? 1.f : a.slope);
            atomicAdd(a.das + j, gg);
            dad += gg;
            dw0 = fmaf(gg, e0, dw0);
            dw1 = fmaf(gg, e1, dw1);
        }
    }
    if (a.mode == 1) {
        if (valid) a.dad[i] = dad;
        dw0 = gat_warp_sum(dw0);
        dw1 = gat_warp_sum(dw1);
        if ((threadIdx.x & 31) == 0 && a.dwe) {
            atomicAdd(a.dwe, dw0);
            atomicAdd(a.dwe + 1, dw1);
        }
        return;
    }
    // mode 2: per channel -- d m_c = d logit att_c lrelu'(m_c) feeds XL[j,c], XR[i,c], We[c,:]; d att_c = d logit lrelu(m_c)
    for (int c = 0; c < a.C; ++c) {
        float dxr = 0.f, da = 0.f, d0 = 0.f, d1 = 0.f;
        const float attc = __ldg(a.att + c), w0 = __ldg(a.We + 2 * c), w1 = __ldg(a.We + 2 * c + 1);
        const float xrc = valid ? __ldg(a.XR + (size_t)i * a.ldr + c) : 0.f;
        for (int kk = k0; kk < k1; ++kk) {
            const int j = __ldg(a.nbr + kk);
            const float e0 = a.ea ? __ldg(a.ea + 2 * (size_t)kk) : 0.f, e1 = a.ea ? __ldg(a.ea + 2 * (size_t)kk + 1) : 0.f;
            const float m = __ldg(a.XL + (size_t)j * a.ldl + c) + xrc + w0 * e0 + w1 * e1;
            const float dl = a.dlog[kk];
            const float dm = dl * attc * (m > 0.f ? 1.f : a.slope);
            atomicAdd(a.dXL + (size_t)j * a.ldl + c, dm);
            dxr += dm;
            da = fmaf(dl, gat_lrelu(m, a.slope), da);
            d0 = fmaf(dm, e0, d0);
            d1 = fmaf(dm, e1, d1);
        }
        if (valid) a.dXR[(size_t)i * a.ldr + c] = dxr;
        da = gat_warp_sum(da);
        d0 = gat_warp_sum(d0);
        d1 = gat_warp_sum(d1);
        if ((threadIdx.x & 31) == 0) {
            if (a.datt) atomicAdd(a.datt + c, da);
            if (a.dWe) {
                atomicAdd(a.dWe + 2 * c, d0);
                atomicAdd(a.dWe + 2 * c + 1, d1);
            }
        }
    }
}

}  // namespace qmp
using namespace qmp;

// Edge phase of GATConv (mode 1) / GATv2Conv (mode 2), one head: see the top of this file.  XL [N, ldl] source rows (C columns);
// mode 1: as_ / ad [N] attention scalars, we [2]; mode 2: XR [N, ldr], We [C, 2], att [C].  Writes out [N, ldo] (no bias) and the
// attention coefficients alpha [E] (in-CSR order, saved for the backward pass).
QMP_API int qmp_gat_fwd(int N, int C, int mode, const int* in_ptr, const int* in_src, const float* ea, const float* XL, int ldl,
                        const float* as_, const float* ad, const float* we, const float* XR, int ldr, const float* We, const float* att,
                        float slope, float* out, int ldo, float* alpha, void* stream) {
    if (N <= 0) return 0;
    QMP_REQUIRE((mode == 1 && as_ && ad && we) || (mode == 2 && XR && We && att), "qmp_gat_fwd: mode 1 needs as / ad / we, mode 2 XR / We / att");
    GatArgs a{};
    a.N = N; a.C = C; a.mode = mode; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.XL = XL; a.ldl = ldl; a.as = as_; a.ad = ad; a.we = we;
    a.XR = XR; a.ldr = ldr; a.We = We; a.att = att; a.slope = slope; a.out = out; a.ldo = ldo; a.alpha = alpha;
    gat_fwd_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(a);
    QMP_LAUNCH_CHECK("gat_fwd_kernel");
    return 0;
}

// Backward of qmp_gat_fwd given dOut [N, lddo]: dXL [N, ldl] and (mode 1) das [N] are zeroed here and accumulated with atomics;
// dad [N] / dXR [N, ldr] are written; dwe [2] / dWe [C, 2] / datt [C] are ACCUMULATED (may be NULL).  dlog [E] is scratch.
QMP_API int qmp_gat_bwd(int N, int C, int mode, const int* in_ptr, const int* in_src, const float* ea, const float* XL, int ldl,
                        const float* as_, const float* ad, const float* we, const float* XR, int ldr, const float* We, const float* att,
                        float slope, const float* alpha, const float* dOut, int lddo, float* dlog, float* dXL, float* das, float* dad,
                        float* dwe, float* dXR, float* dWe, float* datt, void* stream) {
    if (N <= 0) return 0;
    QMP_REQUIRE((mode == 1 && as_ && ad && we && das && dad) || (mode == 2 && XR && We && att && dXR), "qmp_gat_bwd: missing arguments for the mode");
    QMP_REQUIRE(ldl == C, "qmp_gat_bwd: dXL is zeroed as one block (ldl must equal C)");
    GatArgs a{};
    a.N = N; a.C = C; a.mode = mode; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.XL = XL; a.ldl = ldl; a.as = as_; a.ad = ad; a.we = we;
    a.XR = XR; a.ldr = ldr; a.We = We; a.att = att; a.slope = slope; a.alpha = const_cast<float*>(alpha); a.dOut = dOut; a.lddo = lddo;
    a.dlog = dlog; a.dXL = dXL; a.das = das; a.dad = dad; a.dwe = dwe; a.dXR = dXR; a.dWe = dWe; a.datt = datt;
    cudaStream_t st = (cudaStream_t)stream;
    QMP_CUDA(cudaMemsetAsync(dXL, 0, (size_t)N * C * sizeof(float), st));
    if (mode == 1) QMP_CUDA(cudaMemsetAsync(das, 0, (size_t)N * sizeof(float), st));
    gat_bwd_kernel<<<cdiv(N, 128), 128, 0, st>>>(a);
    QMP_LAUNCH_CHECK("gat_bwd_kernel");
    return 0;
}
