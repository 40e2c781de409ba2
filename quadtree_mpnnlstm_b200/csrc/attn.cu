// TransformerConv message passing over the in-CSR: edge logits, segment softmax, weighted aggregation,
// forward and backward, for G convolutions that share one graph in a single launch.
//
// PyG 2.2.0 TransformerConv (heads=1, concat=False, edge_dim=2, root_weight=True; reference
// model/model.py:51, restated in oracle/convs_ref.py) computes, for target i and sources j:
//     q_i = Wq x_i + bq,  k_ij = Wk x_j + bk + We e_ij,  v_ij = Wv x_j + bv + We e_ij
//     alpha_ij = softmax_j(q_i . k_ij / sqrt(C)),   out_i = sum_j alpha_ij v_ij + Ws x_i + bs
// Softmax is invariant to the j-independent term q_i . bk, and the remaining logit is linear in x_j:
//     s_ij = u_i . x_j + w_i . e_ij,   u_i = Wk^T q_i / sqrt(C),   w_i = We^T q_i / sqrt(C)
// and the aggregate is linear too:
//     out_i = Wv (sum_j alpha_ij x_j) + bv (sum_j alpha_ij) + We (sum_j alpha_ij e_ij) + Ws x_i + bs
// So the edges only ever touch the RAW D-wide node rows x_j (gathered once, coalesced), not the
// C-wide keys / values; everything dense is a per-node GEMM before (u, w) and after (out) this kernel.
//
// Thread layout: 8 lanes per (node, conv) pair; lane l holds features l, l+8, ... (Q per lane), so a
// 32-float row is one 128-byte request per 8 lanes.  Dot products reduce with 3 xor-shuffles inside
// the 8-lane group.  Softmax is the online form (running max / sum), one pass over the neighbours.
#include "common.cuh"

namespace qmp {

struct AttnArgs {
    int N, G, D;
    const int* ptr;       // in-CSR row pointer [N+1]           (fwd, bwd_target)
    const int* nbr;       // in_src [E]                          (fwd, bwd_target)
    const float* ea;      // edge attrs in in-CSR order [E, 2]
    const float* x;       // node rows; conv g of node i reads x + i*ldx + g*xoff, D floats
    int ldx, xoff;
    const float* U;       // [N, G*(D+2)]: u (D), w (2) per conv
    float* Z;             // [N, G*(D+3)]: xbar (D), eabar (2), alpha-sum (1)
    float* logit;         // [E, G]
    float* mstat;         // [N, G] running max
    float* linv;          // [N, G] 1 / sum exp
    float drop_p;         // attention dropout probability (0 = off)
    unsigned long long seed; const unsigned long long* salt;
    // backward
    const float* dZ;      // [N, G*(D+3)]
    float* dU;            // [N, G*(D+2)]
    float* ds;            // [E, G]
    // backward, source side
    const int* optr;      // out-CSR row pointer
    const int* odst;      // out_dst [E]
    const int* okin;      // out_kin [E]
    float* dx;            // conv g of node j accumulates into dx + j*lddx + g*dxoff
    int lddx, dxoff;
    int dx_accumulate;    // 0: overwrite, 1: add to what is there
};

__device__ __forceinline__ float group8_sum(float v, unsigned mask) {
    v += __shfl_xor_sync(mask, v, 1);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 4);
    return v;
}

// counter-based keep mask for attention dropout: same (seed, edge slot, conv) -> same decision in fwd and bwd
__device__ __forceinline__ float dropout_scale(unsigned long long seed, long long idx, float p) {
    if (p <= 0.f) return 1.f;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    return (u >= p) ? 1.f / (1.f - p) : 0.f;
}

template <int Q>
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pair = t >> 3;
    const int sub = (int)(t & 7);
    if (pair >= (long long)a.N * a.G) return;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    const int i = (int)(pair / a.G), g = (int)(pair % a.G);
    const int D = a.D;
    const float* urow = a.U + (size_t)i * a.G * (D + 2) + (size_t)g * (D + 2);
    float u[Q], xb[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = sub + 8 * q;
        u[q] = (k < D) ? urow[k] : 0.f;
        xb[q] = 0.f;
    }
    const float w0 = urow[D], w1 = urow[D + 1];
    float m = -INFINITY, l = 0.f, e0 = 0.f, e1 = 0.f, asum = 0.f;
    const int k1 = a.ptr[i + 1];
    for (int kk = a.ptr[i]; kk < k1; ++kk) {
        const int j = a.nbr[kk];
        const float* xr = a.x + (size_t)j * a.ldx + (size_t)g * a.xoff;
        float xj[Q];
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = sub + 8 * q;
            xj[q] = (k < D) ? xr[k] : 0.f;
            part = fmaf(u[q], xj[q], part);
        }
        const float a0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, a1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        const float s = group8_sum(part, gmask) + w0 * a0 + w1 * a1;
        if (sub == 0) a.logit[(size_t)kk * a.G + g] = s;
        const float mn = fmaxf(m, s);
        const float sc = expf(m - mn);   // first edge: exp(-inf) = 0
        const float p = expf(s - mn);
        const float pk = p * dropout_scale(QMP_SEED(a), (long long)kk * a.G + g, a.drop_p);
        l = l * sc + p;
        asum = asum * sc + pk;
        e0 = e0 * sc + pk * a0;
        e1 = e1 * sc + pk * a1;
#pragma unroll
        for (int q = 0; q < Q; ++q) xb[q] = xb[q] * sc + pk * xj[q];
        m = mn;
    }
    const float li = (l > 0.f) ? 1.f / l : 0.f;
    float* zrow = a.Z + (size_t)i * a.G * (D + 3) + (size_t)g * (D + 3);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = sub + 8 * q;
        if (k < D) zrow[k] = xb[q] * li;
    }
    if (sub == 0) {
        zrow[D] = e0 * li;
        zrow[D + 1] = e1 * li;
        zrow[D + 2] = asum * li;
        a.mstat[pair] = m;
        a.linv[pair] = li;
    }
}

// target side: d alpha, d logit (ds), dU.  Two passes over the neighbours.
template <int Q>
__global__ void __launch_bounds__(256) attn_bwd_target_kernel(AttnArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pair = t >> 3;
    const int sub = (int)(t & 7);
    if (pair >= (long long)a.N * a.G) return;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    const int i = (int)(pair / a.G), g = (int)(pair % a.G);
    const int D = a.D;
    const float* dz = a.dZ + (size_t)i * a.G * (D + 3) + (size_t)g * (D + 3);
    float dxb[Q], du[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = sub + 8 * q;
        dxb[q] = (k < D) ? dz[k] : 0.f;
        du[q] = 0.f;
    }
    const float de0 = dz[D], de1 = dz[D + 1], dsig = dz[D + 2];
    const float m = a.mstat[pair], li = a.linv[pair];
    const int k0 = a.ptr[i], k1 = a.ptr[i + 1];
    float tsum = 0.f;
    for (int kk = k0; kk < k1; ++kk) {
        const int j = a.nbr[kk];
        const float* xr = a.x + (size_t)j * a.ldx + (size_t)g * a.xoff;
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = sub + 8 * q;
            part = fmaf(dxb[q], (k < D) ? xr[k] : 0.f, part);
        }
        const float a0 = a.ea ? a.ea[(size_t)kk * 2] : 0.f, a1 = a.ea ? a.ea[(size_t)kk * 2 + 1] : 0.f;
        // gradient w.r.t. the (dropped-out) weight, chained through the keep mask
        const float keep = dropout_scale(QMP_SEED(a), (long long)kk * a.G + g, a.drop_p);
        const float dal = keep * (group8_sum(part, gmask) + de0 * a0 + de1 * a1 + dsig);
        const float al = expf(a.logit[(size_t)kk * a.G + g] - m) * li;
        tsum = fmaf(al, dal, tsum);
        if (sub == 0) a.ds[(size_t)kk * a.G + g] = dal;  // stash; finalised below
    }
    float dw0 = 0.f, dw1 = 0.f;
    for (int kk = k0; kk < k1; ++kk) {
        const int j = a.nbr[kk];
        const float* xr = a.x + (size_t)j * a.ldx + (size_t)g * a.xoff;
        const float al = expf(a.logit[(size_t)kk * a.G + g] - m) * li;
        float dal = 0.f;
        if (sub == 0) dal = a.ds[(size_t)kk * a.G + g];
        dal = __shfl_sync(gmask, dal, (threadIdx.x & 31) & ~7);
        const float dsv = al * (dal - tsum);
        if (sub == 0) a.ds[(size_t)kk * a.G + g] = dsv;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = sub + 8 * q;
            du[q] = fmaf(dsv, (k < D) ? xr[k] : 0.f, du[q]);
        }
        if (a.ea) {
            dw0 = fmaf(dsv, a.ea[(size_t)kk * 2], dw0);
            dw1 = fmaf(dsv, a.ea[(size_t)kk * 2 + 1], dw1);
        }
    }
    float* dur = a.dU + (size_t)i * a.G * (D + 2) + (size_t)g * (D + 2);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = sub + 8 * q;
        if (k < D) dur[k] = du[q];
    }
    if (sub == 0) {
        dur[D] = dw0;
        dur[D + 1] = dw1;
    }
}

// source side: dx_j = sum over edges leaving j of (alpha' * dxbar_i + ds * u_i).
// shared != 0: the G convs read the same row, so one 8-lane group sums all of them into one dx row.
template <int Q>
__global__ void __launch_bounds__(256) attn_bwd_source_kernel(AttnArgs a, int shared) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pair = t >> 3;
    const int sub = (int)(t & 7);
    const int groups = shared ? 1 : a.G;
    if (pair >= (long long)a.N * groups) return;
    const int j = (int)(pair / groups);
    const int g_lo = shared ? 0 : (int)(pair % groups), g_hi = shared ? a.G : g_lo + 1;
    const int D = a.D;
    float acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = 0.f;
    const int k0 = a.optr[j], k1 = a.optr[j + 1];
    for (int g = g_lo; g < g_hi; ++g) {
        for (int kk = k0; kk < k1; ++kk) {
            const int i = a.odst[kk], kin = a.okin[kk];
            const long long pi = (long long)i * a.G + g;
            const float al = expf(a.logit[(size_t)kin * a.G + g] - a.mstat[pi]) * a.linv[pi] *
                             dropout_scale(QMP_SEED(a), (long long)kin * a.G + g, a.drop_p);
            const float dsv = a.ds[(size_t)kin * a.G + g];
            const float* dz = a.dZ + (size_t)i * a.G * (D + 3) + (size_t)g * (D + 3);
            const float* ur = a.U + (size_t)i * a.G * (D + 2) + (size_t)g * (D + 2);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int k = sub + 8 * q;
                if (k < D) acc[q] = fmaf(al, dz[k], fmaf(dsv, ur[k], acc[q]));
            }
        }
    }
    float* dst = a.dx + (size_t)j * a.lddx + (size_t)g_lo * a.dxoff;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = sub + 8 * q;
        if (k < D) dst[k] = a.dx_accumulate ? dst[k] + acc[q] : acc[q];
    }
}

#define QMP_DISPATCH_Q(D, CALL)                                             \
    do {                                                                    \
        const int q__ = ((D) + 7) / 8;                                      \
        if (q__ <= 1) { CALL(1); }                                          \
        else if (q__ <= 2) { CALL(2); }                                     \
        else if (q__ <= 4) { CALL(4); }                                     \
        else if (q__ <= 5) { CALL(5); }                                     \
        else if (q__ <= 8) { CALL(8); }                                     \
        else if (q__ <= 9) { CALL(9); }                                     \
        else if (q__ <= 16) { CALL(16); }                                   \
        else if (q__ <= 17) { CALL(17); }                                   \
        else { qmp::set_error("attention: D=%d > 136 unsupported", (D)); return -1; } \
    } while (0)

}  // namespace qmp
using namespace qmp;

// Forward.  x rows: conv g of node j reads x + j*ldx + g*xoff (xoff = 0 when the G convs share the input).
// U [N, G*(D+2)] from the first node GEMM; ea [E,2] in in-CSR order (or NULL).  Writes Z [N, G*(D+3)],
// logit [E,G], mstat/linv [N,G].
QMP_API int qmp_attn_fwd(int N, int G, int D, const int* in_ptr, const int* in_src, const float* ea, const float* x,
                         int ldx, int xoff, const float* U, float* Z, float* logit, float* mstat, float* linv,
                         float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0 || G <= 0) return 0;
    AttnArgs a{};
    a.N = N; a.G = G; a.D = D; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.x = x; a.ldx = ldx; a.xoff = xoff;
    a.U = U; a.Z = Z; a.logit = logit; a.mstat = mstat; a.linv = linv; a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    const int blocks = cdiv((long long)N * G * 8, 256);
#define CALL(QQ) attn_fwd_kernel<QQ><<<blocks, 256, 0, (cudaStream_t)stream>>>(a)
    QMP_DISPATCH_Q(D, CALL);
#undef CALL
    QMP_LAUNCH_CHECK("qmp_attn_fwd");
    return 0;
}

// Backward, target side: reads dZ, writes ds [E,G] (gradient of the logits) and dU [N, G*(D+2)].
QMP_API int qmp_attn_bwd_target(int N, int G, int D, const int* in_ptr, const int* in_src, const float* ea,
                                const float* x, int ldx, int xoff, const float* logit, const float* mstat,
                                const float* linv, const float* dZ, float* ds, float* dU, float drop_p,
                                unsigned long long seed, void* stream) {
    if (N <= 0 || G <= 0) return 0;
    AttnArgs a{};
    a.N = N; a.G = G; a.D = D; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.x = x; a.ldx = ldx; a.xoff = xoff;
    a.logit = const_cast<float*>(logit); a.mstat = const_cast<float*>(mstat); a.linv = const_cast<float*>(linv);
    a.dZ = dZ; a.ds = ds; a.dU = dU; a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    const int blocks = cdiv((long long)N * G * 8, 256);
#define CALL(QQ) attn_bwd_target_kernel<QQ><<<blocks, 256, 0, (cudaStream_t)stream>>>(a)
    QMP_DISPATCH_Q(D, CALL);
#undef CALL
    QMP_LAUNCH_CHECK("qmp_attn_bwd_target");
    return 0;
}

// Backward, source side: dx rows (conv g of node j at dx + j*lddx + g*dxoff; shared=1 sums the G convs
// into one row at dx + j*lddx).  accumulate=1 adds to the existing contents.
QMP_API int qmp_attn_bwd_source(int N, int G, int D, const int* out_ptr, const int* out_dst, const int* out_kin,
                                const float* logit, const float* mstat, const float* linv, const float* ds,
                                const float* dZ, const float* U, float* dx, int lddx, int dxoff, int shared,
                                int accumulate, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0 || G <= 0) return 0;
    AttnArgs a{};
    a.N = N; a.G = G; a.D = D; a.optr = out_ptr; a.odst = out_dst; a.okin = out_kin;
    a.logit = const_cast<float*>(logit); a.mstat = const_cast<float*>(mstat); a.linv = const_cast<float*>(linv);
    a.ds = const_cast<float*>(ds); a.dZ = dZ; a.U = U; a.dx = dx; a.lddx = lddx; a.dxoff = dxoff;
    a.dx_accumulate = accumulate; a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    const int blocks = cdiv((long long)N * (shared ? 1 : G) * 8, 256);
#define CALL(QQ) attn_bwd_source_kernel<QQ><<<blocks, 256, 0, (cudaStream_t)stream>>>(a, shared)
    QMP_DISPATCH_Q(D, CALL);
#undef CALL
    QMP_LAUNCH_CHECK("qmp_attn_bwd_source");
    return 0;
}
