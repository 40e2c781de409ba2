// Third tcgen05 probe (test hook): issue cost of small kind::tf32 MMAs.  One thread issues `reps` MMAs of shape
// 128 x N x 8 (A from tensor memory or from shared memory; all into one accumulator or rotating over four), commits,
// and the CTA measures the clock until the commit lands.  Results (cycles per MMA) drive the tile shapes of the
// fused kernels; the numbers are recorded in DESIGN.md.
#include "../common.cuh"
#include "../tc.cuh"

namespace qmp {

__global__ void __launch_bounds__(128) tc_probe3_kernel(long long* out, int N, int reps, int a_tmem, int rotate) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    for (int idx = t; idx < 48 * 1024 / 4; idx += 128) reinterpret_cast<float*>(smem)[idx] = 0.f;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0;
    if (t == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        const uint64_t da = tc::make_desc(tc::smem_u32(smem), 128, 256);
        const uint64_t db = tc::make_desc(tc::smem_u32(smem) + 8192, 128, 256);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tmem + (rotate ? (uint32_t)(r & 3) * 64u : 0u);
            if (a_tmem) tc::mma_tf32_ts(d, tmem + 480, db, idesc, 1);
            else tc::mma_tf32(d, da, db, idesc, 1);
        }
        tc::commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    if (t == 0) {
        out[0] = clock64() - t0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace qmp
using namespace qmp;

// out[0] = cycles for `reps` MMAs (128 x N x 8, tf32) + commit.  a_tmem: A operand from tensor memory; rotate: four accumulators.
QMP_API int qmp_tc_probe3(long long* out, int N, int reps, int a_tmem, int rotate, void* stream) {
    QMP_REQUIRE(N % 16 == 0 && N >= 16 && N <= 64 + 0 * rotate || (!rotate && N <= 256), "qmp_tc_probe3: bad N");
    QMP_CUDA(cudaFuncSetAttribute(tc_probe3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    tc_probe3_kernel<<<1, 128, 48 * 1024, (cudaStream_t)stream>>>(out, N, reps, a_tmem, rotate);
    QMP_LAUNCH_CHECK("qmp_tc_probe3");
    return 0;
}
