// tcgen05 / TMEM building blocks (sm_100a) for the per-node dense contractions.
//
// The gate "GEMMs" of the cell are [128-node tile] x [tiny K = 8..72] x [N <= 256] products.  They run as
// tcgen05.mma kind::tf32 with the split-precision scheme 3xTF32 (x = hi + lo, hi = x with the 13 low
// mantissa bits cleared, lo = x - hi exactly):  A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32
// accumulation in TMEM; the dropped A_lo B_lo term and the truncation of the lo parts are O(2^-22), i.e.
// fp32-level, which is what the 1e-4-per-forecast-step parity bar needs over 100 recurrent steps (plain
// TF32 is O(2^-11) per product and does not hold it).
//
// Shared-memory operand layout: K-major, SWIZZLE_NONE ("interleave") canonical layout -- 8-row x 16-byte core
// matrices; element (r, k) of a tile with KC = K/4 16-byte chunks per row lives at byte
//     (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4,      LBO = 128, SBO = 128 * KC
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::K>: ((8,n),2):((1,SBO),LBO) in 16-byte units).
// Operands are written by ordinary st.shared from registers (rows arrive through gathers, not TMA: the
// 361-wide grid rows and the CSR gathers are not expressible as tensor-map boxes), then published to the
// async proxy with fence.proxy.async before the MMA is issued.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;     // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version 1 (Blackwell), bits [46,48)
    return d;                                              // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// A operand taken from tensor memory (lane = row of A, 32-bit column = k): D[tmem_d] (+)= A[tmem_a] * B[desc_b]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- bulk asynchronous copy global -> shared (TMA engine, 1-D), completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// L2 prefetch of a contiguous global range (TMA engine, no registers or shared memory involved): the streaming inputs of the
// NEXT tile are requested while the current tile computes, so that its first loads find them in L2
__device__ __forceinline__ void l2_prefetch(const void* p, long long bytes) {
    if (p == nullptr || bytes <= 0) return;
    const unsigned long long a = reinterpret_cast<unsigned long long>(p), a0 = a & ~15ull;
    const unsigned int n = (unsigned int)((bytes + (long long)(a - a0) + 15) & ~15ll);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole-warp: allocate `cols` (power of two >= 32) TMEM columns, base address written to *slot (shared)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
// several allocations by one CTA: allocate each, relinquish the permit once after the last
__device__ __forceinline__ void tmem_alloc_only(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// warp w reads TMEM lanes 32*(w%4)..+31 (one accumulator row per thread): 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// warp w writes TMEM lanes 32*(w%4)..+31: 8 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 8 columns without the wait (issue several, then tmem_ld_wait() once)
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// split for 3xTF32: x = hi + lo with hi ROUNDED to the nearest tf32.  The tensor core reads the upper 19 bits of an fp32
// operand, i.e. it TRUNCATES: with hi = x & 0xFFFFE000 the remainder lo is always of x's sign and keeps up to 13 significant
// bits, the hardware chops it to 11, and every operand carries a ONE-SIDED error of up to 2^-20 |x| -- measured at the full
// configs[1] size as forecasts drifting from 1.5e-6 (step 0) to 2.2e-4 (step 84) against the oracle.  With hi rounded
// (add half an ulp to the bit pattern, then truncate: two full-rate integer instructions; cvt.rna.tf32.f32 does the same on
// the slower conversion pipe) |lo| <= 2^-11 |x| has either sign, lo = x - hi is exact in fp32, and the hardware's truncation
// of lo is an error of at most 2^-21 |x| whose sign follows lo's, i.e. unbiased.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u);
    lo = x - hi;
}

// byte offset of element (r, k) in a K-major no-swizzle tile with KC 16-byte chunks per row
__device__ __forceinline__ uint32_t tile_off(int r, int k, int KC) {
    return (uint32_t)((r >> 3) * (128 * KC) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

}  // namespace tc
}  // namespace qmp
