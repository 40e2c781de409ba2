// Probe (test hook, libqmp_probe.so only): node-major panels brought in by 2-D TMA with the 128-byte swizzle, used as
// MN-major A / B operands of tcgen05.mma kind::tf32 -- the operand form of cell_wgrad.cu -- with every descriptor field a
// run-time argument, plus a dump of the shared-memory stage, so that one GPU run can sweep the conventions.
//   A [K, 128], B [K, NB] fp32 node-major (K = 16 nodes);  D [128, NB] = A^T B;  smem_dump [(4 + NB / 32) * 2048 bytes]
#include "../common.cuh"
#include "../tc.cuh"
#include <cuda.h>

namespace qmp {

__device__ __forceinline__ void mnp_tma(void* dst, const CUtensorMap* tm, int col, int row, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     tc::smem_u32(dst)),
                 "l"(tm), "r"(col), "r"(row), "r"(tc::smem_u32(bar))
                 : "memory");
}

struct MnpArgs {
    float* D; uint8_t* dump;
    int NB, idesc_extra, lbo, sbo, layout_type, kstep_bytes, ksteps;
};

__global__ void __launch_bounds__(128) mn_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                       const MnpArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int t = threadIdx.x, warp = t >> 5;
    const int npanel = 4 + a.NB / 32;
    if (t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        tc::mbar_expect_tx(&bars[0], (uint32_t)npanel * 2048u);
        for (int p = 0; p < 4; ++p) mnp_tma(smem + p * 2048, &tmA, 32 * p, 0, &bars[0]);
        for (int p = 0; p < a.NB / 32; ++p) mnp_tma(smem + (4 + p) * 2048, &tmB, 32 * p, 0, &bars[0]);
    }
    tc::mbar_wait(&bars[0], 0);
    for (int i = t; i < npanel * 2048 / 16; i += 128) reinterpret_cast<float4*>(a.dump)[i] = reinterpret_cast<const float4*>(smem)[i];
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (t == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, a.NB) | (uint32_t)a.idesc_extra;
        for (int ks = 0; ks < a.ksteps; ++ks) {
            auto desc = [&](uint32_t addr) {
                uint64_t d = 0;
                d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
                d |= (uint64_t)(((uint32_t)a.lbo >> 4) & 0x3FFFu) << 16;
                d |= (uint64_t)(((uint32_t)a.sbo >> 4) & 0x3FFFu) << 32;
                d |= (uint64_t)1 << 46;
                d |= (uint64_t)a.layout_type << 61;
                return d;
            };
            const uint32_t base = tc::smem_u32(smem) + (uint32_t)(ks * a.kstep_bytes);
            tc::mma_tf32(tmem, desc(base), desc(base + 4 * 2048), idesc, ks ? 1u : 0u);
        }
        tc::commit(&bars[1]);
    }
    tc::mbar_wait(&bars[1], 0);
    tc::fence_after_sync();
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < a.NB; c0 += 8) {
        float v[8];
        tc::tmem_ld8(lane_base + (uint32_t)c0, v);
        for (int i = 0; i < 8; ++i) a.D[(size_t)t * a.NB + c0 + i] = v[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

typedef CUresult (*mnp_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace qmp
using namespace qmp;

QMP_API int qmp_mn_probe(const float* A, const float* B, float* D, void* smem_dump, int K, int NB, int idesc_extra, int lbo, int sbo,
                         int layout_type, int kstep_bytes, int tma_swizzle, void* stream) {
    QMP_REQUIRE(K == 16 && NB % 32 == 0 && NB >= 32 && NB <= 256, "qmp_mn_probe: K = 16, NB a multiple of 32");
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    QMP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    QMP_REQUIRE(fn != nullptr, "qmp_mn_probe: no cuTensorMapEncodeTiled");
    auto enc = reinterpret_cast<mnp_encode_fn>(fn);
    CUtensorMap tmA, tmB;
    auto mk = [&](CUtensorMap* m, const float* base, int cols) {
        const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)K};
        const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
        const cuuint32_t box[2] = {32u, 16u}, es[2] = {1u, 1u};
        return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        (CUtensorMapSwizzle)tma_swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    QMP_REQUIRE(mk(&tmA, A, 128) == 0 && mk(&tmB, B, NB) == 0, "qmp_mn_probe: tensor map encode failed");
    MnpArgs a{};
    a.D = D; a.dump = (uint8_t*)smem_dump; a.NB = NB; a.idesc_extra = idesc_extra; a.lbo = lbo; a.sbo = sbo; a.layout_type = layout_type;
    a.kstep_bytes = kstep_bytes; a.ksteps = K / 8;
    const int smem = (4 + NB / 32) * 2048 + 1024;
    QMP_CUDA(cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    mn_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, a);
    QMP_LAUNCH_CHECK("qmp_mn_probe");
    return 0;
}
