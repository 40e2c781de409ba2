// Weight pack of a group of G TransformerConvs (layout: fused.cuh) straight from the PyG-named parameter tensors, and its
// backward, as ONE launch each.  Built from tensor ops (slice / pad / cat / bmm under autograd, convs.pack_tconv +
// fused.pack_fused) the eight packs of a training step were ~400 tiny launches, 1.9 ms of the 42 ms step; the arithmetic is a
// (D+2) x (D+1) bilinear form per conv:
//     [W1 | b1] = [Wk | We]^T [Wq | bq] / sqrt(C)     (logit_ij * sqrt(C) = q_i . (Wk x_j + bk + We e_ij); the bk term cancels in
//     W2 = [Wv | We | bv],  W3 = Ws,  b3 = bs           the softmax -- convs.pack_tconv has the derivation)
// zero padded to the caps (DC columns, 32 rows).  The parameters are reached through a device table of pointers
// tab [G][9] = lin_query.{weight,bias}, lin_key.{weight,bias}, lin_value.{weight,bias}, lin_edge.weight, lin_skip.{weight,bias}
// (weights row-major [C, D] / [C, 2], as torch.nn.Linear stores them).  Reference: PyG TransformerConv's parameters as the
// reference instantiates them (model/model.py:51; seq2seq.py:117-121).
#include "common.cuh"
#include "fused.cuh"

namespace qmp {

struct PackConv {
    const float *Wq, *bq, *Wk, *bk, *Wv, *bv, *We, *Ws, *bs;
};
__device__ __forceinline__ PackConv pack_conv(const long long* __restrict__ tab, int g) {
    const long long* t = tab + (size_t)g * 9;
    PackConv p;
    p.Wq = reinterpret_cast<const float*>(t[0]); p.bq = reinterpret_cast<const float*>(t[1]);
    p.Wk = reinterpret_cast<const float*>(t[2]); p.bk = reinterpret_cast<const float*>(t[3]);
    p.Wv = reinterpret_cast<const float*>(t[4]); p.bv = reinterpret_cast<const float*>(t[5]);
    p.We = reinterpret_cast<const float*>(t[6]); p.Ws = reinterpret_cast<const float*>(t[7]);
    p.bs = reinterpret_cast<const float*>(t[8]);
    return p;
}
// [Wk | We][k][r], r in 0..D+1, and [Wq | bq][k][c], c in 0..D
__device__ __forceinline__ float pk_ke(const PackConv& p, int k, int r, int D) { return r < D ? p.Wk[k * D + r] : p.We[k * 2 + (r - D)]; }
__device__ __forceinline__ float pk_qb(const PackConv& p, int k, int c, int D) { return c < D ? p.Wq[k * D + c] : p.bq[k]; }
// row of the unpadded [D+2] index space <-> padded position (rows / entries D..DC-1 are padding; the two edge rows sit at DC, DC+1)
__device__ __forceinline__ int pk_unpad(int j, int D, int DC) { return j < DC ? (j < D ? j : -1) : (j < DC + 2 ? D + (j - DC) : -1); }
__device__ __forceinline__ int pk_pad(int r, int D, int DC) { return r < D ? r : DC + (r - D); }

__global__ void __launch_bounds__(256) pack_tconv_fwd_kernel(const long long* __restrict__ tab, int D, int DC, int C, float* __restrict__ out) {
    const PackConv p = pack_conv(tab, blockIdx.x);
    const int o1 = (DC + 2) * DC, o2 = o1 + DC + 4, o3 = o2 + FC * (DC + 4), o4 = o3 + FC * DC, total = o4 + FC;
    float* o = out + (size_t)blockIdx.x * total;
    const float s = rsqrtf((float)C);
    for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += gridDim.y * blockDim.x) {
        float v = 0.f;
        if (idx < o2) {                                       // W1 | b1: the bilinear form
            int rr, c;
            if (idx < o1) { rr = pk_unpad(idx / DC, D, DC); c = idx % DC; if (c >= D) rr = -1; }
            else { rr = pk_unpad(idx - o1, D, DC); c = D; }
            if (rr >= 0) {
                float acc = 0.f;
                for (int k = 0; k < C; ++k) acc = fmaf(pk_ke(p, k, rr, D), pk_qb(p, k, c, D), acc);
                v = acc * s;
            }
        } else if (idx < o3) {                                // W2 = [Wv | 0 | We | bv | 0]
            const int r = (idx - o2) / (DC + 4), j = (idx - o2) % (DC + 4);
            if (r < C) {
                if (j < D) v = p.Wv[r * D + j];
                else if (j == DC || j == DC + 1) v = p.We[r * 2 + (j - DC)];
                else if (j == DC + 2) v = p.bv[r];
            }
        } else if (idx < o4) {
            const int r = (idx - o3) / DC, c = (idx - o3) % DC;
            if (r < C && c < D) v = p.Ws[r * D + c];
        } else {
            const int r = idx - o4;
            if (r < C) v = p.bs[r];
        }
        o[idx] = v;
    }
}

// g [G, TOTAL(DC)] -> grads [G, P], P = C (4 D + 6): gWq [C,D] | gbq [C] | gWk [C,D] | gbk [C] (= 0) | gWv [C,D] | gbv [C] | gWe [C,2] |
// gWs [C,D] | gbs [C]
__global__ void __launch_bounds__(256) pack_tconv_bwd_kernel(const long long* __restrict__ tab, int D, int DC, int C, const float* __restrict__ g,
                                                             float* __restrict__ grads) {
    const PackConv p = pack_conv(tab, blockIdx.x);
    const int o1 = (DC + 2) * DC, o2 = o1 + DC + 4, o3 = o2 + FC * (DC + 4), o4 = o3 + FC * DC, total = o4 + FC;
    const float* gp = g + (size_t)blockIdx.x * total;
    const int P = C * (4 * D + 6);
    float* out = grads + (size_t)blockIdx.x * P;
    const float s = rsqrtf((float)C);
    // gradient of the bilinear form's output [D+2][D+1] read through the padding
    auto gW1b = [&](int rr, int c) { return c < D ? gp[pk_pad(rr, D, DC) * DC + c] : gp[o1 + pk_pad(rr, D, DC)]; };
    const int nq = C * D, q0 = 0, q1 = nq, q2 = q1 + C, q3 = q2 + nq, q4 = q3 + C, q5 = q4 + nq, q6 = q5 + C, q7 = q6 + 2 * C, q8 = q7 + nq;
    for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < P; idx += gridDim.y * blockDim.x) {
        float v = 0.f;
        if (idx < q2) {                                       // gWq | gbq = s [Wk | We] gW1b
            const int k = idx < q1 ? (idx - q0) / D : idx - q1, c = idx < q1 ? (idx - q0) % D : D;
            for (int rr = 0; rr < D + 2; ++rr) v = fmaf(pk_ke(p, k, rr, D), gW1b(rr, c), v);
            v *= s;
        } else if (idx < q3) {                                // gWk = s [Wq | bq] gW1b^T (rows 0..D-1)
            const int k = (idx - q2) / D, rr = (idx - q2) % D;
            for (int c = 0; c <= D; ++c) v = fmaf(pk_qb(p, k, c, D), gW1b(rr, c), v);
            v *= s;
        } else if (idx < q4) {
            v = 0.f;                                          // lin_key.bias: no influence on the output
        } else if (idx < q5) {
            const int k = (idx - q4) / D, j = (idx - q4) % D;
            v = gp[o2 + k * (DC + 4) + j];
        } else if (idx < q6) {
            v = gp[o2 + (idx - q5) * (DC + 4) + DC + 2];
        } else if (idx < q7) {                                // gWe: through the logits (rows D, D+1 of the form) and through the values
            const int k = (idx - q6) / 2, j = (idx - q6) % 2;
            for (int c = 0; c <= D; ++c) v = fmaf(pk_qb(p, k, c, D), gW1b(D + j, c), v);
            v = v * s + gp[o2 + k * (DC + 4) + DC + j];
        } else if (idx < q8) {
            const int k = (idx - q7) / D, c = (idx - q7) % D;
            v = gp[o3 + k * DC + c];
        } else {
            v = gp[o4 + (idx - q8)];
        }
        out[idx] = v;
    }
}

}  // namespace qmp
using namespace qmp;

// out [G, TOTAL(DC)] = padded weight pack of G TransformerConvs (in D, out C <= 32) from the parameter pointer table tab [G, 9]
// (device int64; order at the top of this file).
QMP_API int qmp_pack_tconv_fwd(const long long* tab, int G, int D, int DC, int C, float* out, void* stream) {
    if (G <= 0) return 0;
    QMP_REQUIRE(D >= 1 && D <= DC && C >= 1 && C <= FC, "qmp_pack_tconv_fwd: need 1 <= D <= DC, 1 <= C <= 32");
    pack_tconv_fwd_kernel<<<dim3(G, 16), 256, 0, (cudaStream_t)stream>>>(tab, D, DC, C, out);      // 16 slices of a conv's pack per CTA row
    QMP_LAUNCH_CHECK("pack_tconv_fwd_kernel");
    qmp::after_producer();
    return 0;
}

// Backward of qmp_pack_tconv_fwd: g [G, TOTAL(DC)] -> grads [G, C (4 D + 6)], per conv gWq | gbq | gWk | gbk | gWv | gbv | gWe | gWs | gbs.
QMP_API int qmp_pack_tconv_bwd(const long long* tab, int G, int D, int DC, int C, const float* g, float* grads, void* stream) {
    if (G <= 0) return 0;
    QMP_REQUIRE(D >= 1 && D <= DC && C >= 1 && C <= FC, "qmp_pack_tconv_bwd: need 1 <= D <= DC, 1 <= C <= 32");
    pack_tconv_bwd_kernel<<<dim3(G, 16), 256, 0, (cudaStream_t)stream>>>(tab, D, DC, C, g, grads);
    QMP_LAUNCH_CHECK("pack_tconv_bwd_kernel");
    return 0;
}
