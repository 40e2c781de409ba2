// Weight gradients of the decoder cell (4 TransformerConvs on the 4-wide X + 4 on the 32-wide H, one pair per gate) in ONE
// streaming launch: TMA tiles -> tcgen05, no SIMT staging of operands.  Counterpart of fused_cell_bwd.cu, which writes the
// per-node rows this kernel reduces over the mesh nodes (autograd of model/model.py:394-463 around PyG TransformerConv):
//
//     gW3 = sum_i g_i (x) x_i,  gb3 = sum_i g_i,  gW2 = sum_i g_i (x) Z_i,  gW1 = sum_i dU_i (x) x_i,  gb1 = sum_i dU_i
//
// Every one of these is a "TN" product with the NODE index as the reduction dimension.  The rows are node-major in memory
// ([node][component]), i.e. exactly the MN-major operand form of tcgen05.mma: one node = one 128-byte row of 32 components,
// 8 nodes = K of one kind::tf32 instruction.  For 32-bit MN-major operands the hardware has ONE shared-memory layout: the
// 128-byte swizzle with 32-byte atoms (UMMA LayoutType SWIZZLE_128B_BASE32B: 32-byte chunk index XOR (row & 3), 4-row atoms;
// the plain 128-byte swizzle returns zeros -- scripts/mn_probe.py).  A 2-D tensor map with a {32 components, 16 nodes} box and
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B drops a panel into shared memory in exactly that layout, and the whole cell needs TWO
// wide instructions per 8 nodes (x3 for the TF32 split):
//
//     D1 [128 x 224] += [g_i g_f g_c g_o] (x) [Zh_0 Zh_1 Zh_2 Zh_3 | sd1 | h | sd0]      (gW2, gW3, gb3 of all eight convs)
//     D2 [ 64 x 160] += [h | sd0]         (x) [du_0 du_1 du_2 du_3 | sg]                 (gW1, gb1 of all eight convs)
//
// with the small panels  sd0 = x(4) | 1 0 0 0 | (ze0 ze1 zs 0) of the four H convs | 0 (8),   sd1 = (z(4) | ze0 ze1 zs 0) of
// the four X convs,   sg = (dw0 dw1) of the four H convs | du(4) of the four X convs | (dw0 dw1) of the four X convs.
// Cross-gate blocks of D1 are computed and ignored (the issue cost of an MMA does not depend on N).  The old kernel
// (fused_wgrad.cu, still used for the other conv groups) issued 4 x 24 narrow MMAs per 64 nodes and moved every operand
// through registers: 59-77 us per decoder frame against ~20 us of tensor-pipe time here.
//
// Roles (192 threads, one persistent CTA per SM, 16-node stages round-robin over the CTAs, 3-deep ring):
//   warp 4: TMA producer (16 panel loads per stage, completion counted in bytes on an mbarrier);
//   warps 0-3: 3xTF32 split in place -- hi = tf32-rounded value (overwrites the panel), lo = x - hi into the stage's second
//              half -- an elementwise pass, no transposition; afterwards the flush of the accumulators;
//   warp 5: MMA issuer; tcgen05.commit releases the stage back to the producer.
// Accumulators stay in tensor memory for the CTA's whole node range and are flushed once, through shared memory, with
// 16-byte vector reductions (red.global.add.v4.f32) into the padded pack layout (fused.cuh).
#include "common.cuh"
#include "fused.cuh"
#include "tc.cuh"
#include <cuda.h>

namespace qmp {

constexpr int CW_NODES = 16;                                  // nodes per stage (two K = 8 steps)
constexpr int CW_PANEL = CW_NODES * 128;                      // bytes of one 32-component panel
constexpr int CW_NPANEL = 16;
constexpr int CW_HALF = CW_NPANEL * CW_PANEL;                 // hi half of a stage; the lo half follows
constexpr int CW_STAGE = 2 * CW_HALF;
constexpr int CW_NSTAGE = 3;
constexpr int CW_SMEM = CW_NSTAGE * CW_STAGE + 1024;          // + slack for the 1024-byte alignment of the swizzle atoms
// panel order inside a stage: A of D1 | B of D1 (its last two panels are also the A of D2) | B of D2
constexpr int CW_P_DP = 0, CW_P_ZH = 4, CW_P_SD1 = 8, CW_P_H = 9, CW_P_SD0 = 10, CW_P_DU = 11, CW_P_SG = 15;
constexpr int CW_N1 = 224, CW_N2 = 160;
constexpr uint32_t CW_D1 = 0, CW_D2 = 224;                    // tensor-memory columns of the two accumulators
constexpr int CW_T1LD = 228, CW_T2LD = 164;                   // flush tiles in shared memory (floats per row)
constexpr int CW_THREADS = 192;
static_assert(128 * CW_T1LD * 4 + 64 * CW_T2LD * 4 <= CW_NSTAGE * CW_STAGE, "flush tiles reuse the ring");

struct CwgArgs {
    int N;
    float* gwa;                                               // [4, TOTAL(4)]  X convs
    float* gwb;                                               // [4, TOTAL(32)] H convs
};

// shared-memory matrix descriptor, MN-major, SWIZZLE_128B_BASE32B (cute::UMMA::SmemDescriptor, layout_type 1):
// ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units -- LBO = distance between 32-component panels, SBO = between 4-node atoms
__device__ __forceinline__ uint64_t cw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

__device__ __forceinline__ void cw_tma(void* dst, const CUtensorMap* tm, int col, int row, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     tc::smem_u32(dst)),
                 "l"(tm), "r"(col), "r"(row), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cw_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cw_sync_flush() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void cw_red4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(CW_THREADS, 1)
    cell_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dp, const __grid_constant__ CUtensorMap tm_z,
                      const __grid_constant__ CUtensorMap tm_sd, const __grid_constant__ CUtensorMap tm_h,
                      const __grid_constant__ CUtensorMap tm_du, const __grid_constant__ CUtensorMap tm_sg, const CwgArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_raw[CW_NSTAGE], full_lo[CW_NSTAGE], empty[CW_NSTAGE], done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int nstages = (a.N + CW_NODES - 1) / CW_NODES;
    const int my = (int)blockIdx.x < nstages ? (nstages - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < CW_NSTAGE; ++s) {
            tc::mbar_init(&full_raw[s], 1);
            tc::mbar_init(&full_lo[s], 4);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(&done, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    pdl_wait();
    pdl_launch();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == 4) {
        if (lane == 0) {                                       // ---- producer
            for (int it = 0; it < my; ++it) {
                const int slot = it % CW_NSTAGE;
                if (it >= CW_NSTAGE) tc::mbar_wait(&empty[slot], (uint32_t)((it / CW_NSTAGE - 1) & 1));
                uint8_t* base = smem + slot * CW_STAGE;
                const int node0 = ((int)blockIdx.x + it * (int)gridDim.x) * CW_NODES;
                uint64_t* bar = &full_raw[slot];
                tc::mbar_expect_tx(bar, (uint32_t)CW_HALF);
#pragma unroll
                for (int p = 0; p < 4; ++p) cw_tma(base + (CW_P_DP + p) * CW_PANEL, &tm_dp, 32 * p, node0, bar);
#pragma unroll
                for (int p = 0; p < 4; ++p) cw_tma(base + (CW_P_ZH + p) * CW_PANEL, &tm_z, 32 * p, node0, bar);
                cw_tma(base + CW_P_SD1 * CW_PANEL, &tm_sd, 32, node0, bar);
                cw_tma(base + CW_P_H * CW_PANEL, &tm_h, 0, node0, bar);
                cw_tma(base + CW_P_SD0 * CW_PANEL, &tm_sd, 0, node0, bar);
#pragma unroll
                for (int p = 0; p < 4; ++p) cw_tma(base + (CW_P_DU + p) * CW_PANEL, &tm_du, 32 * p, node0, bar);
                cw_tma(base + CW_P_SG * CW_PANEL, &tm_sg, 0, node0, bar);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {                                       // ---- MMA issuer
            const uint32_t mn = (1u << 15) | (1u << 16);       // a_major = b_major = MN
            const uint32_t id1 = tc::make_idesc_tf32(128, CW_N1) | mn, id2 = tc::make_idesc_tf32(128, CW_N2) | mn;
            const uint32_t lbo = (uint32_t)CW_PANEL, sbo = 512u;
            for (int it = 0; it < my; ++it) {
                const int slot = it % CW_NSTAGE;
                tc::mbar_wait(&full_lo[slot], (uint32_t)((it / CW_NSTAGE) & 1));
                tc::fence_after_sync();
                const uint32_t hi = tc::smem_u32(smem + slot * CW_STAGE), lo = hi + CW_HALF;
#pragma unroll
                for (int ks = 0; ks < CW_NODES / 8; ++ks) {
                    const uint32_t off = (uint32_t)ks * 1024u;
                    const uint32_t acc = (it | ks) ? 1u : 0u;
                    const uint64_t a1h = cw_desc(hi + CW_P_DP * CW_PANEL + off, lbo, sbo), a1l = cw_desc(lo + CW_P_DP * CW_PANEL + off, lbo, sbo);
                    const uint64_t b1h = cw_desc(hi + CW_P_ZH * CW_PANEL + off, lbo, sbo), b1l = cw_desc(lo + CW_P_ZH * CW_PANEL + off, lbo, sbo);
                    tc::mma_tf32(tmem + CW_D1, a1h, b1h, id1, acc);
                    tc::mma_tf32(tmem + CW_D1, a1l, b1h, id1, 1u);
                    tc::mma_tf32(tmem + CW_D1, a1h, b1l, id1, 1u);
                    const uint64_t a2h = cw_desc(hi + CW_P_H * CW_PANEL + off, lbo, sbo), a2l = cw_desc(lo + CW_P_H * CW_PANEL + off, lbo, sbo);
                    const uint64_t b2h = cw_desc(hi + CW_P_DU * CW_PANEL + off, lbo, sbo), b2l = cw_desc(lo + CW_P_DU * CW_PANEL + off, lbo, sbo);
                    tc::mma_tf32(tmem + CW_D2, a2h, b2h, id2, acc);
                    tc::mma_tf32(tmem + CW_D2, a2l, b2h, id2, 1u);
                    tc::mma_tf32(tmem + CW_D2, a2h, b2l, id2, 1u);
                }
                tc::commit(&empty[slot]);                      // the stage is free once these MMAs have read it
            }
            if (my > 0) tc::commit(&done);
        }
    } else {
        // ---- split: hi in place (tf32, rounded to nearest), lo = x - hi into the second half of the stage
        for (int it = 0; it < my; ++it) {
            const int slot = it % CW_NSTAGE;
            tc::mbar_wait(&full_raw[slot], (uint32_t)((it / CW_NSTAGE) & 1));
            float4* hi = reinterpret_cast<float4*>(smem + slot * CW_STAGE);
            float4* lo = reinterpret_cast<float4*>(smem + slot * CW_STAGE + CW_HALF);
#pragma unroll 4
            for (int i = t; i < CW_HALF / 16; i += 128) {
                const float4 v = hi[i];
                float4 h, l;
                tc::split_tf32(v.x, h.x, l.x);
                tc::split_tf32(v.y, h.y, l.y);
                tc::split_tf32(v.z, h.z, l.z);
                tc::split_tf32(v.w, h.w, l.w);
                hi[i] = h;
                lo[i] = l;
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) cw_arrive(&full_lo[slot]);
        }
        if (my > 0) {
            // ---- flush: accumulator rows -> shared memory tiles -> vector reductions along the pack rows
            tc::mbar_wait(&done, 0);
            tc::fence_after_sync();
            float* T1 = reinterpret_cast<float*>(smem);
            float* T2 = T1 + 128 * CW_T1LD;
            const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < CW_N1; c0 += 8) {
                float v[8];
                tc::tmem_ld8(lane_base + CW_D1 + (uint32_t)c0, v);
                float* d = T1 + t * CW_T1LD + c0;
                *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (warp < 2) {
                for (int c0 = 0; c0 < CW_N2; c0 += 8) {
                    float v[8];
                    tc::tmem_ld8(lane_base + CW_D2 + (uint32_t)c0, v);
                    float* d = T2 + t * CW_T2LD + c0;
                    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            cw_sync_flush();
            using SA = ConvSizes<4>;
            using SB = ConvSizes<32>;
            constexpr int o1a = SA::W1, o2a = o1a + SA::B1, o3a = o2a + SA::W2, o4a = o3a + SA::W3;
            constexpr int o1b = SB::W1, o2b = o1b + SB::B1, o3b = o2b + SB::W2, o4b = o3b + SB::W3;
            // D1: row = gate s, output r.  20 vectors per row: 8 Zh_s | 1 (ze0 ze1 zs 0)_s | 8 h | 2 Zx_s | 1 x
            for (int idx = t; idx < 128 * 20; idx += 128) {
                const int row = idx / 20, k = idx - row * 20, s = row >> 5, r = row & 31;
                const float* src = T1 + row * CW_T1LD;
                float* gb = a.gwb + (size_t)s * SB::TOTAL;
                float* ga = a.gwa + (size_t)s * SA::TOTAL;
                if (k < 8) cw_red4(gb + o2b + r * 36 + 4 * k, *reinterpret_cast<const float4*>(src + 32 * s + 4 * k));
                else if (k == 8) cw_red4(gb + o2b + r * 36 + 32, *reinterpret_cast<const float4*>(src + 200 + 4 * s));
                else if (k < 17) cw_red4(gb + o3b + r * 32 + 4 * (k - 9), *reinterpret_cast<const float4*>(src + 160 + 4 * (k - 9)));
                else if (k < 19) cw_red4(ga + o2a + r * 8 + 4 * (k - 17), *reinterpret_cast<const float4*>(src + 128 + 8 * s + 4 * (k - 17)));
                else cw_red4(ga + o3a + r * 4, *reinterpret_cast<const float4*>(src + 192));
            }
            {                                                  // gb3 of both convs of gate s: the "1" column
                const int s = t >> 5, r = t & 31;
                const float v = T1[t * CW_T1LD + 196];
                atomicAdd(a.gwb + (size_t)s * SB::TOTAL + o4b + r, v);
                atomicAdd(a.gwa + (size_t)s * SA::TOTAL + o4a + r, v);
            }
            // D2: row = data component (h 0..31 | x 32..35 | "1" 36), column = [du_h 4 x 32 | dw_h 4 x 2 | du_x 4 x 4 | dw_x 4 x 2]
            for (int idx = t; idx < 136 * 8; idx += 128) {     // gW1 of the H convs: 4 consecutive h components per vector
                const int col = idx >> 3, cq = idx & 7;
                const int c = col < 128 ? col >> 5 : (col - 128) >> 1;
                const int R = col < 128 ? (col & 31) : 32 + ((col - 128) & 1);
                const float4 v = make_float4(T2[(4 * cq) * CW_T2LD + col], T2[(4 * cq + 1) * CW_T2LD + col], T2[(4 * cq + 2) * CW_T2LD + col],
                                             T2[(4 * cq + 3) * CW_T2LD + col]);
                cw_red4(a.gwb + (size_t)c * SB::TOTAL + R * 32 + 4 * cq, v);
            }
            if (t < 24) {                                      // gW1 of the X convs: the four x components
                const int col = 136 + t;
                const int c = t < 16 ? t >> 2 : (t - 16) >> 1;
                const int R = t < 16 ? (t & 3) : 4 + ((t - 16) & 1);
                const float4 v = make_float4(T2[32 * CW_T2LD + col], T2[33 * CW_T2LD + col], T2[34 * CW_T2LD + col], T2[35 * CW_T2LD + col]);
                cw_red4(a.gwa + (size_t)c * SA::TOTAL + R * 4, v);
            }
            for (int col = t; col < CW_N2; col += 128) {       // gb1: the "1" row
                const float v = T2[36 * CW_T2LD + col];
                float* dst;
                if (col < 128) dst = a.gwb + (size_t)(col >> 5) * SB::TOTAL + o1b + (col & 31);
                else if (col < 136) dst = a.gwb + (size_t)((col - 128) >> 1) * SB::TOTAL + o1b + 32 + ((col - 128) & 1);
                else if (col < 152) dst = a.gwa + (size_t)((col - 136) >> 2) * SA::TOTAL + o1a + ((col - 136) & 3);
                else dst = a.gwa + (size_t)((col - 152) >> 1) * SA::TOTAL + o1a + 4 + ((col - 152) & 1);
                atomicAdd(dst, v);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

typedef CUresult (*cw_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [N, cols] fp32 rows of pitch ld floats -> boxes of {32 components, CW_NODES nodes}, 128-byte swizzle with 32-byte atoms, rows
// past N read as zeros
static int cw_make_map(cw_encode_fn enc, CUtensorMap* m, const float* base, int cols, int N, int ld) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)CW_NODES};
    const cuuint32_t es[2] = {1u, 1u};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

}  // namespace qmp
using namespace qmp;

// Weight gradients of the decoder cell: accumulates into gwa [4, TOTAL(4)] (X convs) and gwb [4, TOTAL(32)] (H convs), forward
// pack layout (fused.cuh), caller zero-initialises.  h [N, ldh >= 32] is the hidden state the cell read; dP [N, lddp >= 128] the
// gate pre-activation gradients; zB / duB [N, 128], sd [N, 64], sg [N, 32] the rows written by qmp_fused_cell_bwd.
QMP_API int qmp_cell_wgrad(int N, const float* h, int ldh, const float* dP, int lddp, const float* zB, const float* duB, const float* sd,
                           const float* sg, float* gwa, float* gwb, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(ldh % 4 == 0 && ldh >= 32 && lddp % 4 == 0 && lddp >= 128 && al16(h) && al16(dP) && al16(zB) && al16(duB) && al16(sd) &&
                    al16(sg) && al16(gwa) && al16(gwb),
                "qmp_cell_wgrad: rows must be 16-byte aligned");
    static cw_encode_fn enc = nullptr;
    static int n_sm = 0;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        QMP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        QMP_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "qmp_cell_wgrad: cuTensorMapEncodeTiled is not available");
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QMP_CUDA(cudaFuncSetAttribute(cell_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM));
        enc = reinterpret_cast<cw_encode_fn>(fn);
    }
    CUtensorMap tm_dp, tm_z, tm_sd, tm_h, tm_du, tm_sg;
    int rc = 0;
    rc |= cw_make_map(enc, &tm_dp, dP, 128, N, lddp);
    rc |= cw_make_map(enc, &tm_z, zB, 128, N, 128);
    rc |= cw_make_map(enc, &tm_sd, sd, 64, N, 64);
    rc |= cw_make_map(enc, &tm_h, h, 32, N, ldh);
    rc |= cw_make_map(enc, &tm_du, duB, 128, N, 128);
    rc |= cw_make_map(enc, &tm_sg, sg, 32, N, 32);
    QMP_REQUIRE(rc == 0, "qmp_cell_wgrad: cuTensorMapEncodeTiled failed (%d)", rc);
    CwgArgs a{};
    a.N = N; a.gwa = gwa; a.gwb = gwb;
    const int nstages = cdiv(N, CW_NODES);
    QMP_CUDA(launch_pdl(cell_wgrad_kernel, dim3(nstages < n_sm ? nstages : n_sm), dim3(CW_THREADS), CW_SMEM, (cudaStream_t)stream, tm_dp, tm_z, tm_sd,
                        tm_h, tm_du, tm_sg, a));
    QMP_LAUNCH_CHECK("cell_wgrad_kernel");
    return 0;
}
