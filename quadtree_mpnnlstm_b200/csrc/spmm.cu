// Weighted neighbourhood sums for GCNConv / ChebConv (PyG 2.2.0 semantics restated in
// oracle/convs_ref.py; selected by the reference at model/model.py:39-57).
//
//   GCN  (add_self_loops=False): deg_i = sum of w over edges ENTERING i; val = deg^-1/2[src] w deg^-1/2[dst]
//   Cheb (sym, lambda_max = 2):  self-loops dropped; deg_j = sum of w over edges LEAVING j;
//                                val = -deg^-1/2[src] w deg^-1/2[dst]; the +1 / -1 diagonals cancel
// Values are stored per in-CSR slot.  y = alpha * (S x) + beta * z, forward through the in-CSR
// (gather from sources) and transposed through the out-CSR (gather from targets, value looked up
// through out_kin) -- no scatter, no atomics, fixed summation order.
#include "common.cuh"

namespace qmp {

__global__ void degree_kernel(const int* __restrict__ ptr, const int* __restrict__ nbr, const int* __restrict__ vidx,
                              const float* __restrict__ w, int N, int drop_self, float* __restrict__ dis) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float d = 0.f;
    for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
        if (drop_self && nbr[k] == i) continue;
        d += w ? w[vidx ? vidx[k] : k] : 1.f;
    }
    const float r = 1.0f / sqrtf(d);          // torch pow(-0.5); inf (deg 0) -> 0
    dis[i] = isinf(r) ? 0.f : r;
}

__global__ void edge_norm_kernel(const int* __restrict__ in_ptr, const int* __restrict__ in_src, const float* __restrict__ w,
                                 const float* __restrict__ dis, int N, int cheb, float* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int k = in_ptr[i]; k < in_ptr[i + 1]; ++k) {
        const int j = in_src[k];
        const float wk = w ? w[k] : 1.f;
        float v = dis[j] * wk * dis[i];
        if (cheb) v = (j == i) ? 0.f : -v;
        val[k] = v;
    }
}

// y[i, c] = alpha * sum_k val[vidx ? vidx[k] : k] * x[nbr[k], c] + beta * z[i, c]
__global__ void spmm_kernel(const int* __restrict__ ptr, const int* __restrict__ nbr, const int* __restrict__ vidx,
                            const float* __restrict__ val, int N, int width, const float* __restrict__ x, int ldx,
                            float alpha, float beta, const float* __restrict__ z, int ldz, float* __restrict__ y, int ldy) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * width) return;
    const int i = (int)(t / width), c = (int)(t % width);
    float s = 0.f;
    const int k1 = ptr[i + 1];
    for (int k = ptr[i]; k < k1; ++k) s = fmaf(val[vidx ? vidx[k] : k], x[(size_t)nbr[k] * ldx + c], s);
    float out = alpha * s;
    if (z) out += beta * z[(size_t)i * ldz + c];
    y[(size_t)i * ldy + c] = out;
}

}  // namespace qmp
using namespace qmp;

// mode 0 = GCN norm, 1 = Cheb norm.  w [E] in in-CSR order or NULL (all ones).  val [E] (in-CSR order),
// dis [N] scratch.
QMP_API int qmp_edge_norm(int mode, int N, const int* in_ptr, const int* in_src, const int* out_ptr, const int* out_dst,
                          const int* out_kin, const float* w, float* dis, float* val, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0) return 0;
    if (mode == 0)
        degree_kernel<<<cdiv(N, 256), 256, 0, st>>>(in_ptr, in_src, nullptr, w, N, 0, dis);
    else
        degree_kernel<<<cdiv(N, 256), 256, 0, st>>>(out_ptr, out_dst, out_kin, w, N, 1, dis);
    edge_norm_kernel<<<cdiv(N, 256), 256, 0, st>>>(in_ptr, in_src, w, dis, N, mode, val);
    QMP_LAUNCH_CHECK("qmp_edge_norm");
    return 0;
}

// Forward: ptr=in_ptr, nbr=in_src, vidx=NULL.  Transposed: ptr=out_ptr, nbr=out_dst, vidx=out_kin.
QMP_API int qmp_spmm(int N, int width, const int* ptr, const int* nbr, const int* vidx, const float* val, const float* x,
                     int ldx, float alpha, float beta, const float* z, int ldz, float* y, int ldy, void* stream) {
    if ((long long)N * width == 0) return 0;
    spmm_kernel<<<cdiv((long long)N * width, 256), 256, 0, (cudaStream_t)stream>>>(ptr, nbr, vidx, val, N, width, x, ldx,
                                                                                 alpha, beta, z, ldz, y, ldy);
    QMP_LAUNCH_CHECK("qmp_spmm");
    return 0;
}
