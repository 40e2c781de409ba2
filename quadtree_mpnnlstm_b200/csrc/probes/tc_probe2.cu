// Second tcgen05 probe (test hook): the two operand conventions the fused kernels use beyond tc_probe.cu.
//   variant 0..3: C [128, N] = A^T B with A [K, 128] and B [K, N] node-major in memory, i.e. MN-major operands in the
//                 canonical no-swizzle layout of fused_wgrad.cu.  bit 0: set the a_major / b_major descriptor bits;
//                 bit 1: swap the LBO / SBO descriptor fields (diagnostic).
//   variant 4:    C [128, N] = A B^T with A [128, K] taken from TENSOR MEMORY (written by tcgen05.st, lane = row,
//                 column = k) and B [N, K] K-major in shared memory.
// Plain TF32 (no split): layout questions show up as O(1) errors, TF32 rounding as ~1e-3.
#include "../common.cuh"
#include "../tc.cuh"

namespace qmp {

__global__ void __launch_bounds__(128) tc_probe2_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ C, int N, int K, int variant) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t cols = 512;
    if (t == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (variant < 4) {
        const int SBO = K * 16 + 16;
        uint8_t* a_s = smem;
        uint8_t* b_s = smem + 32 * SBO;
        for (int idx = t; idx < (32 + N / 4) * SBO / 4; idx += 128) reinterpret_cast<float*>(smem)[idx] = 0.f;
        __syncthreads();
        for (int idx = t; idx < K * 32; idx += 128) {          // A: node k, chunk q of 4 M-values
            const int k = idx / 32, q = idx % 32;
            *reinterpret_cast<float4*>(a_s + q * SBO + (k >> 3) * 128 + (k & 7) * 16) =
                *reinterpret_cast<const float4*>(A + (size_t)k * 128 + q * 4);
        }
        for (int idx = t; idx < K * (N / 4); idx += 128) {
            const int k = idx / (N / 4), q = idx % (N / 4);
            *reinterpret_cast<float4*>(b_s + q * SBO + (k >> 3) * 128 + (k & 7) * 16) =
                *reinterpret_cast<const float4*>(B + (size_t)k * N + q * 4);
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (t == 0) {
            uint32_t idesc = tc::make_idesc_tf32(128, N);
            if (variant & 1) idesc |= (1u << 15) | (1u << 16);
            const uint32_t lbo = (variant & 2) ? (uint32_t)SBO : 128u, sbo = (variant & 2) ? 128u : (uint32_t)SBO;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t da = tc::make_desc(tc::smem_u32(a_s) + ks * 128, lbo, sbo);
                const uint64_t db = tc::make_desc(tc::smem_u32(b_s) + ks * 128, lbo, sbo);
                tc::mma_tf32(tmem, da, db, idesc, ks ? 1u : 0u);
            }
            tc::commit(&bar);
        }
    } else {
        // A row t -> TMEM lane t, columns 256 .. 256+K ; B K-major canonical layout (tc.cuh) in shared memory
        const int KC = K / 4;
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(A[(size_t)t * K + k0 + i]);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                             tmem + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)k0),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int n = t; n < N; n += 128)
            for (int k = 0; k < K; ++k) *(float*)(smem + tc::tile_off(n, k, KC)) = B[(size_t)n * K + k];
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (t == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(128, N);
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t db = tc::make_desc(tc::smem_u32(smem) + ks * 256, 128, 128u * KC);
                const uint32_t ta = tmem + 256u + (uint32_t)ks * 8;
                const uint32_t accum = ks ? 1u : 0u;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                    "}\n" ::"r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
            tc::commit(&bar);
        }
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) C[(size_t)t * N + c0 + i] = v[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, cols);
}

}  // namespace qmp
using namespace qmp;

// Test hook, see the top of this file.  N % 16 == 0, 16 <= N <= 128; K % 8 == 0, 8 <= K <= 64.
QMP_API int qmp_tc_probe2(const float* A, const float* B, float* C, int N, int K, int variant, void* stream) {
    QMP_REQUIRE(N % 16 == 0 && N >= 16 && N <= 128 && K % 8 == 0 && K >= 8 && K <= 64, "qmp_tc_probe2: bad shape");
    const size_t smem = variant < 4 ? (size_t)(32 + N / 4) * (K * 16 + 16) : (size_t)N * K * 4;
    QMP_CUDA(cudaFuncSetAttribute(tc_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    tc_probe2_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, N, K, variant);
    QMP_LAUNCH_CHECK("qmp_tc_probe2");
    return 0;
}
