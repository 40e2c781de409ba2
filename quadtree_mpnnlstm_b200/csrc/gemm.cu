// Dense per-node contractions of the cell (the "gate GEMM"), fp32 CUDA-core path.
//
//   gemm_nt : C[b][i, j] (+)= sum_k A[b][i, k] * B[b][j, k] + bias[b][j]      i < n nodes, j < m outputs
//   gemm_nn : C[b][i, j] (+)= sum_k A[b][i, k] * B[b][k, j]                   (data gradients: dX = dY * W)
//   gemm_tn : C[b][i, j]  += sum_r A[b][r, i] * B[b][r, j]                    r < n nodes (weight gradients)
// All row-major with leading dimensions and per-batch element strides, so a batch can address the
// gate blocks of one wide activation row (stride = block width) or share one operand (stride = 0).
// K is tiny here (4..200): these contractions are HBM/L2-bound, fp32 accumulation throughout.
// 64 x 64 output tile, 16-deep k panel, 256 threads x (4 x 4) registers.
#include "common.cuh"

namespace qmp {

constexpr int BM = 64, BN = 64, BK = 16;

struct GemmArgs {
    const float* A; const float* B; const float* bias; float* C;
    int n, m, k;                   // C is n x m
    int lda, ldb, ldc;
    long long sA, sB, sC, sBias;   // batch strides (elements)
    int accumulate;                // C += instead of C =
    int relu;                      // apply max(.,0) after bias / accumulate
    int b_ones;                    // gemm_tn: B has a virtual trailing column of ones (bias gradients)
    int rows_per_chunk;            // gemm_tn: rows of A/B reduced by one CTA
    int chunks;
};

template <bool B_IS_KxM>
__global__ void __launch_bounds__(256) gemm_rowA_kernel(GemmArgs g) {
    // A tile: rows = nodes (i), read A[i, k].  B tile: NT reads B[j, k]; NN reads B[k, j].
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int b = blockIdx.z;
    const float* A = g.A + b * g.sA;
    const float* B = g.B + b * g.sB;
    float* C = g.C + b * g.sC;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < g.k; k0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = threadIdx.x + 256 * q;
            {   // A: 64 rows x 16 k, consecutive threads along k
                const int r = idx >> 4, kk = idx & 15;
                const int gi = i0 + r, gk = k0 + kk;
                As[kk][r] = (gi < g.n && gk < g.k) ? A[(size_t)gi * g.lda + gk] : 0.f;
            }
            if (B_IS_KxM) {  // B[k, j]: consecutive threads along j
                const int kk = idx >> 6, c = idx & 63;
                const int gj = j0 + c, gk = k0 + kk;
                Bs[kk][c] = (gj < g.m && gk < g.k) ? B[(size_t)gk * g.ldb + gj] : 0.f;
            } else {         // B[j, k]: consecutive threads along k
                const int c = idx >> 4, kk = idx & 15;
                const int gj = j0 + c, gk = k0 + kk;
                Bs[kk][c] = (gj < g.m && gk < g.k) ? B[(size_t)gj * g.ldb + gk] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = As[kk][ty * 4 + q];
                bb[q] = Bs[kk][tx * 4 + q];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], bb[q], acc[p][q]);
        }
        __syncthreads();
    }
    const float* bias = g.bias ? g.bias + b * g.sBias : nullptr;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int gi = i0 + ty * 4 + p;
        if (gi >= g.n) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gj = j0 + tx * 4 + q;
            if (gj >= g.m) continue;
            float v = acc[p][q];
            if (bias) v += bias[gj];
            float* dst = C + (size_t)gi * g.ldc + gj;
            if (g.accumulate) v += *dst;
            if (g.relu) v = fmaxf(v, 0.f);
            *dst = v;
        }
    }
}

// C[i, j] += sum_r A[r, i] * B[r, j]; grid.z = batch * chunks, each CTA reduces rows_per_chunk rows
__global__ void __launch_bounds__(256) gemm_tn_kernel(GemmArgs g) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int b = blockIdx.z / g.chunks, chunk = blockIdx.z % g.chunks;
    const float* A = g.A + b * g.sA;
    const float* B = g.B + b * g.sB;
    float* C = g.C + b * g.sC;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r_begin = chunk * g.rows_per_chunk;
    const int r_end = min(g.n, r_begin + g.rows_per_chunk);
    const int mb = g.b_ones ? g.k - 1 : g.k;  // here g.m = cols of A (i), g.k = logical cols of B (j)
    float acc[4][4] = {};
    for (int r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = threadIdx.x + 256 * q;
            const int rr = idx >> 6, c = idx & 63;
            const int gr = r0 + rr;
            const bool rok = gr < r_end;
            As[rr][c] = (rok && i0 + c < g.m) ? A[(size_t)gr * g.lda + i0 + c] : 0.f;
            float bv = 0.f;
            if (rok) {
                const int gj = j0 + c;
                if (gj < mb) bv = B[(size_t)gr * g.ldb + gj];
                else if (gj < g.k) bv = 1.f;
            }
            Bs[rr][c] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = As[kk][ty * 4 + q];
                bb[q] = Bs[kk][tx * 4 + q];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], bb[q], acc[p][q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int gi = i0 + ty * 4 + p;
        if (gi >= g.m) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gj = j0 + tx * 4 + q;
            if (gj >= g.k) continue;
            atomicAdd(C + (size_t)gi * g.ldc + gj, acc[p][q]);
        }
    }
}

// tcgen05 paths (tc_gemm.cu); return 1 when they took the call
int tc_gemm_rows_try(const float* A, const float* B, const float* bias, float* C, int n, int m, int k, int lda, int ldb,
                     int ldc, long long sA, long long sB, long long sC, long long sBias, int batch, int b_is_kxm,
                     int accumulate, int relu, cudaStream_t st, int* rc);
int tc_gemm_tn_try(const float* A, const float* B, float* C, int n, int ma, int mb, int lda, int ldb, int ldc,
                   long long sA, long long sB, long long sC, int batch, int b_ones, cudaStream_t st, int* rc);
static int g_use_tc = 0;   // measured on B200 (profiles/r01_c): for K <= 40 the smem staging + TMEM round trip outweighs the math

}  // namespace qmp
using namespace qmp;

// 1: dense contractions on tcgen05 (3xTF32) when the shape fits; 0: FFMA kernels (default).  Returns the old value.
QMP_API int qmp_set_tensor_cores(int enable) {
    const int old = g_use_tc;
    g_use_tc = enable ? 1 : 0;
    return old;
}

// C[b] (n x m) (+)= A[b] (n x k) * op(B[b]) + bias[b];  b_is_kxm = 0: B is m x k (C = A B^T); 1: B is k x m.
QMP_API int qmp_gemm(const float* A, const float* B, const float* bias, float* C, int n, int m, int k, int lda, int ldb,
                     int ldc, long long sA, long long sB, long long sC, long long sBias, int batch, int b_is_kxm,
                     int accumulate, int relu, void* stream) {
    if (n <= 0 || m <= 0 || batch <= 0) return 0;
    QMP_REQUIRE(k >= 0 && batch <= 65535, "qmp_gemm: bad k/batch");
    if (g_use_tc) {
        int rc = 0;
        if (tc_gemm_rows_try(A, B, bias, C, n, m, k, lda, ldb, ldc, sA, sB, sC, sBias, batch, b_is_kxm, accumulate, relu,
                             (cudaStream_t)stream, &rc))
            return rc;
    }
    GemmArgs g{A, B, bias, C, n, m, k, lda, ldb, ldc, sA, sB, sC, sBias, accumulate, relu, 0, 0, 0};
    dim3 grid(cdiv(m, BN), cdiv(n, BM), batch);
    if (b_is_kxm)
        gemm_rowA_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    else
        gemm_rowA_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    QMP_LAUNCH_CHECK("qmp_gemm");
    return 0;
}

// C[b] (ma x mb) += A[b]^T (n x ma)^T * B[b] (n x mb); with b_ones the last of the mb columns of B is an
// implicit column of ones (it is not read from memory).  C must be initialised by the caller.
QMP_API int qmp_gemm_tn_acc(const float* A, const float* B, float* C, int n, int ma, int mb, int lda, int ldb, int ldc,
                            long long sA, long long sB, long long sC, int batch, int b_ones, void* stream) {
    if (n <= 0 || ma <= 0 || mb <= 0 || batch <= 0) return 0;
    if (g_use_tc) {
        int rc = 0;
        if (tc_gemm_tn_try(A, B, C, n, ma, mb, lda, ldb, ldc, sA, sB, sC, batch, b_ones, (cudaStream_t)stream, &rc)) return rc;
    }
    int rows = 256;
    int chunks = cdiv(n, rows);
    while ((long long)chunks * batch > 60000) {
        rows *= 2;
        chunks = cdiv(n, rows);
    }
    GemmArgs g{A, B, nullptr, C, n, ma, mb, lda, ldb, ldc, sA, sB, sC, 0, 1, 0, b_ones, rows, chunks};
    dim3 grid(cdiv(mb, BN), cdiv(ma, BM), batch * chunks);
    gemm_tn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    QMP_LAUNCH_CHECK("qmp_gemm_tn_acc");
    return 0;
}
