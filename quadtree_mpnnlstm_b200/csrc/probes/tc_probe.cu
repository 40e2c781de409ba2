// Stand-alone check of the tcgen05 path: C[M x N] = A[M x K] * B[N x K]^T for one or more 128-row tiles,
// 3xTF32 (or plain TF32), operands staged through shared memory by ordinary stores, accumulator in TMEM.
// Exists so the descriptor / layout / fence conventions of tc.cuh are validated in isolation on the GPU
// (tests/test_parity_convs.py::test_tcgen05_gemm) before the fused cell kernels rely on them.
#include "../common.cuh"
#include "../tc.cuh"

namespace qmp {

// one CTA = one 128-row tile; 128 threads (thread t stages and later reads back row t)
template <int SPLIT>
__global__ void __launch_bounds__(128) tc_gemm_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ C, int M, int N, int K, uint32_t tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int KC = K / 4;
    const uint32_t a_bytes = 128u * K * 4, b_bytes = (uint32_t)N * K * 4;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    const int t = threadIdx.x, warp = t >> 5;
    const int row0 = blockIdx.x * 128;

    if (t == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, tmem_cols);

    // stage A (this thread's row) and B (rows strided over the block), split into hi / lo
    {
        const int gr = row0 + t;
        for (int k = 0; k < K; ++k) {
            const float x = (gr < M) ? A[(size_t)gr * K + k] : 0.f;
            float hi, lo;
            tc::split_tf32(x, hi, lo);
            *(float*)(a_hi + tc::tile_off(t, k, KC)) = SPLIT ? hi : x;
            *(float*)(a_lo + tc::tile_off(t, k, KC)) = lo;
        }
        for (int n = t; n < N; n += 128)
            for (int k = 0; k < K; ++k) {
                float hi, lo;
                tc::split_tf32(B[(size_t)n * K + k], hi, lo);
                *(float*)(b_hi + tc::tile_off(n, k, KC)) = SPLIT ? hi : B[(size_t)n * K + k];
                *(float*)(b_lo + tc::tile_off(n, k, KC)) = lo;
            }
    }
    tc::fence_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (t == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        const uint32_t lbo = 128, sbo = 128u * KC;
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 8; ++ks) {                       // one MMA covers K = 8 (32 bytes = 2 chunks)
            const uint32_t koff = (uint32_t)ks * 2 * lbo;
            const uint64_t dah = tc::make_desc(tc::smem_u32(a_hi) + koff, lbo, sbo);
            const uint64_t dal = tc::make_desc(tc::smem_u32(a_lo) + koff, lbo, sbo);
            const uint64_t dbh = tc::make_desc(tc::smem_u32(b_hi) + koff, lbo, sbo);
            const uint64_t dbl = tc::make_desc(tc::smem_u32(b_lo) + koff, lbo, sbo);
            tc::mma_tf32(tmem, dah, dbh, idesc, acc);
            acc = 1;
            if (SPLIT) {
                tc::mma_tf32(tmem, dal, dbh, idesc, 1);
                tc::mma_tf32(tmem, dah, dbl, idesc, 1);
            }
        }
        tc::commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();

    // epilogue: thread t owns accumulator row t (TMEM lane t); warp w may only touch lanes 32*(w%4)..+31
    const int gr = row0 + t;
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (gr < M) {
#pragma unroll
            for (int i = 0; i < 8; ++i) C[(size_t)gr * N + c0 + i] = v[i];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

}  // namespace qmp
using namespace qmp;

// C [M, N] = A [M, K] * B [N, K]^T on the tensor cores (test hook).  K % 8 == 0, N % 8 == 0, 8 <= N <= 256.
// split = 1: 3xTF32 (fp32-level accuracy); 0: plain TF32.
QMP_API int qmp_tc_gemm_probe(const float* A, const float* B, float* C, int M, int N, int K, int split, void* stream) {
    QMP_REQUIRE(K % 8 == 0 && K >= 8 && N % 8 == 0 && N >= 8 && N <= 256, "qmp_tc_gemm_probe: need K % 8 == 0, N % 8 == 0, N <= 256");
    uint32_t cols = 32;
    while ((int)cols < N) cols <<= 1;
    const size_t smem = 2 * (size_t)(128 + N) * K * 4;
    QMP_REQUIRE(smem <= 200 * 1024, "qmp_tc_gemm_probe: tile does not fit in shared memory");
    auto kern = split ? tc_gemm_probe_kernel<1> : tc_gemm_probe_kernel<0>;
    QMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(M, 128), 128, smem, (cudaStream_t)stream>>>(A, B, C, M, N, K, cols);
    QMP_LAUNCH_CHECK("qmp_tc_gemm_probe");
    return 0;
}
