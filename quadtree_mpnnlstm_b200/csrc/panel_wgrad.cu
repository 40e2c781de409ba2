// Weight gradients of ONE TransformerConv (fused.cuh pack layout) as a streaming launch in the scheme of cell_wgrad.cu: the
// per-node rows are node-major in memory, a 2-D tensor map (box {32 components, 16 nodes}, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
// drops 32-component panels into shared memory in the MN-major operand layout of tcgen05.mma kind::tf32, and ONE wide
// instruction per 8 nodes (x3 for the TF32 split) forms every product the conv needs:
//
//     D [128 x 128] += [g | dU(0:32) | dU(32:40) | .] (x) [x(0:32) | x(32:DC), 1 | Z(0:32) | Z(32:DC+4)]
//
//     rows g      : gW3 = g (x) x,  gb3 = g (x) 1,  gW2 = g (x) Z          dU = [du | dw] (DC + 2 used),  Z = [z | ze0 ze1 zs 0]
//     rows dU     : gW1 = dU (x) x, gb1 = dU (x) 1
//
// Columns of a row beyond the tensor's width are zero-filled by the TMA unit, so the 36-wide input rows and the 40-wide dU / Z
// rows of the decoder's head conv fc_out1 (model/seq2seq.py:117-121) need no repacking; the "1" column is written into the
// shared-memory panel by the split warps.  Used for DC = 36 (the head conv, 90 launches per sample): 27-36 us with the
// per-problem kernel of fused_wgrad.cu (one thread per component, operands through registers) against ~10 us here.
// Roles as in cell_wgrad.cu: warp 4 TMA producer, warps 0-3 split hi / lo in place + flush, warp 5 MMA issuer.
#include "common.cuh"
#include "fused.cuh"
#include "tc.cuh"
#include <cuda.h>

namespace qmp {

constexpr int PW_NODES = 16, PW_PANEL = PW_NODES * 128, PW_NPANEL = 7;
constexpr int PW_HALF = PW_NPANEL * PW_PANEL, PW_STAGE = 2 * PW_HALF, PW_NSTAGE = 6;
constexpr int PW_SMEM = PW_NSTAGE * PW_STAGE + 1024;
constexpr int PW_P_G = 0, PW_P_DU0 = 1, PW_P_DU1 = 2, PW_P_X0 = 3, PW_P_X1 = 4, PW_P_Z0 = 5, PW_P_Z1 = 6;
constexpr int PW_TLD = 132;                                   // flush tile row stride (floats)
constexpr int PW_THREADS = 192;
static_assert(128 * PW_TLD * 4 <= PW_NSTAGE * PW_STAGE, "the flush tile reuses the ring");

constexpr int PW_MAXCONV = 12, PW_GROUP = 4;                  // convs per launch; convs whose accumulators share a CTA's tensor memory
constexpr int PW_ONE = 24;                                    // component of panel X1 that holds the constant 1 (column 56 of the x space)
struct PwConv {
    int seg;                                                  // 0: narrow-input conv (maps xa / dua / za), 1: wide (xb / dub / zb)
    int col_g, col_x, col_u;                                  // first column of the conv's g block, input block, dU / Z block
    int DC, x_hi;                                             // cap of the input width; x_hi: the input has columns 32.. (load panel X1)
    float* gw;                                                // [TOTAL(DC)] gradient of the conv's pack
};
struct PwArgs {
    int N, nconv, cpg;                                        // cpg: CTAs per conv group
    PwConv c[PW_MAXCONV];
};
struct PwMaps {
    CUtensorMap g, dua, za, xa, dub, zb, xb;
};

__device__ __forceinline__ uint64_t pw_desc(uint32_t smem_addr) {      // MN-major, SWIZZLE_128B_BASE32B: see cell_wgrad.cu
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((PW_PANEL >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((512u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ void pw_tma(void* dst, const CUtensorMap* tm, int col, int row, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     tc::smem_u32(dst)),
                 "l"(tm), "r"(col), "r"(row), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void pw_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pw_red4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(PW_THREADS, 1) panel_wgrad_kernel(const __grid_constant__ PwMaps m, const __grid_constant__ PwArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_raw[PW_NSTAGE], full_lo[PW_NSTAGE], empty[PW_NSTAGE], done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int nstages = (a.N + PW_NODES - 1) / PW_NODES;
    // CTA -> (conv group, rank in the group); work item it -> (stage rank + (it / cnt) * cpg, conv cbeg + it % cnt)
    const int group = (int)blockIdx.x / a.cpg, rank = (int)blockIdx.x - group * a.cpg;
    const int cbeg = group * PW_GROUP, cnt = a.nconv - cbeg < PW_GROUP ? a.nconv - cbeg : PW_GROUP;
    const int mystages = rank < nstages ? (nstages - 1 - rank) / a.cpg + 1 : 0;
    const int my = mystages * cnt;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < PW_NSTAGE; ++s) {
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
        if (lane == 0) {
            for (int it = 0; it < my; ++it) {
                const int slot = it % PW_NSTAGE;
                if (it >= PW_NSTAGE) tc::mbar_wait(&empty[slot], (uint32_t)((it / PW_NSTAGE - 1) & 1));
                uint8_t* base = smem + slot * PW_STAGE;
                const PwConv& cv = a.c[cbeg + it % cnt];
                const int node0 = (rank + (it / cnt) * a.cpg) * PW_NODES;
                uint64_t* bar = &full_raw[slot];
                const CUtensorMap* mu = cv.seg ? &m.dub : &m.dua;
                const CUtensorMap* mz = cv.seg ? &m.zb : &m.za;
                const CUtensorMap* mx = cv.seg ? &m.xb : &m.xa;
                tc::mbar_expect_tx(bar, (uint32_t)((cv.x_hi ? 7 : 6) * PW_PANEL));
                pw_tma(base + PW_P_G * PW_PANEL, &m.g, cv.col_g, node0, bar);
                pw_tma(base + PW_P_DU0 * PW_PANEL, mu, cv.col_u, node0, bar);
                pw_tma(base + PW_P_DU1 * PW_PANEL, mu, cv.col_u + 32, node0, bar);
                pw_tma(base + PW_P_X0 * PW_PANEL, mx, cv.col_x, node0, bar);
                if (cv.x_hi) pw_tma(base + PW_P_X1 * PW_PANEL, mx, cv.col_x + 32, node0, bar);
                pw_tma(base + PW_P_Z0 * PW_PANEL, mz, cv.col_u, node0, bar);
                pw_tma(base + PW_P_Z1 * PW_PANEL, mz, cv.col_u + 32, node0, bar);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(128, 128) | (1u << 15) | (1u << 16);
            for (int it = 0; it < my; ++it) {
                const int slot = it % PW_NSTAGE;
                tc::mbar_wait(&full_lo[slot], (uint32_t)((it / PW_NSTAGE) & 1));
                tc::fence_after_sync();
                const uint32_t hi = tc::smem_u32(smem + slot * PW_STAGE), lo = hi + PW_HALF;
                const uint32_t d = tmem + 128u * (uint32_t)(it % cnt);           // this conv's accumulator
#pragma unroll
                for (int ks = 0; ks < PW_NODES / 8; ++ks) {
                    const uint32_t off = (uint32_t)ks * 1024u;
                    const uint64_t ah = pw_desc(hi + PW_P_G * PW_PANEL + off), al = pw_desc(lo + PW_P_G * PW_PANEL + off);
                    const uint64_t bh = pw_desc(hi + PW_P_X0 * PW_PANEL + off), bl = pw_desc(lo + PW_P_X0 * PW_PANEL + off);
                    tc::mma_tf32(d, ah, bh, idesc, (it >= cnt || ks) ? 1u : 0u);
                    tc::mma_tf32(d, al, bh, idesc, 1u);
                    tc::mma_tf32(d, ah, bl, idesc, 1u);
                }
                tc::commit(&empty[slot]);
            }
            if (my > 0) tc::commit(&done);
        }
    } else {
        for (int it = 0; it < my; ++it) {
            const int slot = it % PW_NSTAGE;
            tc::mbar_wait(&full_raw[slot], (uint32_t)((it / PW_NSTAGE) & 1));
            uint8_t* base = smem + slot * PW_STAGE;
            if (!a.c[cbeg + it % cnt].x_hi)                    // panel X1 was not loaded: it only carries the constant 1
                reinterpret_cast<float4*>(base + PW_P_X1 * PW_PANEL)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t < PW_NODES) {                                // the constant-1 component of the valid nodes (32-byte chunks XOR (node & 3))
                const int node0 = (rank + (it / cnt) * a.cpg) * PW_NODES;
                constexpr int c = PW_ONE;
                if (node0 + t < a.N)
                    *reinterpret_cast<float*>(base + PW_P_X1 * PW_PANEL + t * 128 + (((c >> 3) ^ (t & 3)) << 5) + ((c & 7) << 2)) = 1.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float4* hi = reinterpret_cast<float4*>(base);
            float4* lo = reinterpret_cast<float4*>(base + PW_HALF);
#pragma unroll
            for (int i = t; i < PW_HALF / 16; i += 128) {
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
            if (lane == 0) pw_arrive(&full_lo[slot]);
        }
        if (my > 0) {
            tc::mbar_wait(&done, 0);
            tc::fence_after_sync();
            float* T = reinterpret_cast<float*>(smem);
            const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
            for (int k = 0; k < cnt; ++k) {
                const PwConv& cv = a.c[cbeg + k];
                for (int c0 = 0; c0 < 128; c0 += 8) {
                    float v[8];
                    tc::tmem_ld8(lane_base + 128u * (uint32_t)k + (uint32_t)c0, v);
                    float* d = T + t * PW_TLD + c0;
                    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int DC = cv.DC, W = DC + 4, q = DC / 4, qz = W / 4;
                const int o1 = (DC + 2) * DC, o2 = o1 + DC + 4, o3 = o2 + FC * W, o4 = o3 + FC * DC;
                // rows g (0..31): gW3 (q vectors) | gW2 (qz vectors);  rows dU (32 .. 32 + DC + 1): gW1 (q vectors)
                for (int idx = t; idx < FC * (q + qz); idx += 128) {
                    const int r = idx / (q + qz), kk = idx - r * (q + qz);
                    const float* src = T + r * PW_TLD;
                    if (kk < q) pw_red4(cv.gw + o3 + r * DC + 4 * kk, *reinterpret_cast<const float4*>(src + 4 * kk));
                    else pw_red4(cv.gw + o2 + r * W + 4 * (kk - q), *reinterpret_cast<const float4*>(src + 64 + 4 * (kk - q)));
                }
                for (int idx = t; idx < (DC + 2) * q; idx += 128) {
                    const int r = idx / q, kk = idx - r * q;
                    pw_red4(cv.gw + r * DC + 4 * kk, *reinterpret_cast<const float4*>(T + (32 + r) * PW_TLD + 4 * kk));
                }
                constexpr int oc = 32 + PW_ONE;                // the "1" column: gb3 and gb1
                if (t < FC) atomicAdd(cv.gw + o4 + t, T[t * PW_TLD + oc]);
                else if (t < FC + DC + 2) atomicAdd(cv.gw + o1 + (t - FC), T[t * PW_TLD + oc]);
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

typedef CUresult (*pw_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int pw_make_map(pw_encode_fn enc, CUtensorMap* m, const float* base, int cols, int N, int ld) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)PW_NODES};
    const cuuint32_t es[2] = {1u, 1u};
    return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace qmp
using namespace qmp;

// Weight gradients of one fused layer group: same arguments, results and accumulation semantics as qmp_fused_wgrad
// (fused_wgrad.cu), computed by the streaming kernel above -- every conv of the group is one [g | dU] (x) [x | 1 | Z] product
// per 8 nodes on TMA-staged panels; groups of four convs share a CTA's tensor memory, the CTAs are split between the groups.
QMP_API int qmp_fused_wgrad_tma(int N, const float* xa, int lda, int DA, int GA, const float* xb, int ldb, int DB, int GB,
                                int sharedB, int mode, int C, const float* dP, int lddp, const float* ZsA, const float* dUsA,
                                const float* ZsB, const float* dUsB, float* gwa, float* gwb, void* stream) {
    if (N <= 0) return 0;
    const int NC = GA + GB;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(NC >= 1 && NC <= PW_MAXCONV, "qmp_fused_wgrad_tma: at most %d convs per group", PW_MAXCONV);
    QMP_REQUIRE(DB >= 1 && DB <= 56 && DA >= 0 && DA <= 8 && C >= 1 && C <= FC, "qmp_fused_wgrad_tma: unsupported sizes");
    QMP_REQUIRE(ldb % 4 == 0 && lddp % 4 == 0 && (GA == 0 || lda % 4 == 0) && al16(xa) && al16(xb) && al16(dP) && al16(ZsA) && al16(dUsA) &&
                    al16(ZsB) && al16(dUsB) && al16(gwa) && al16(gwb),
                "qmp_fused_wgrad_tma: rows must be 16-byte aligned");
    const int dac = (GA == 0) ? 0 : (DA <= 4 ? 4 : 8);
    const int dbc = (DB <= 32) ? 32 : (DB + 3) / 4 * 4;
    static pw_encode_fn enc = nullptr;
    static int n_sm = 0;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        QMP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        QMP_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "qmp_fused_wgrad_tma: cuTensorMapEncodeTiled is not available");
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QMP_CUDA(cudaFuncSetAttribute(panel_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM));
        enc = reinterpret_cast<pw_encode_fn>(fn);
    }
    PwMaps m;
    int rc = 0;
    // g: the columns a conv may read are its own C outputs (mode 0) or its gate's 32 (mode 1); the tensor's width bounds the rest
    rc |= pw_make_map(enc, &m.g, dP, mode == 1 ? 4 * FC : NC * C, N, lddp);
    if (GA) {
        rc |= pw_make_map(enc, &m.dua, dUsA, GA * (dac + 4), N, GA * (dac + 4));
        rc |= pw_make_map(enc, &m.za, ZsA, GA * (dac + 4), N, GA * (dac + 4));
        rc |= pw_make_map(enc, &m.xa, xa, DA, N, lda);
    } else {
        m.dua = m.za = m.xa = m.g;
    }
    rc |= pw_make_map(enc, &m.dub, dUsB, GB * (dbc + 4), N, GB * (dbc + 4));
    rc |= pw_make_map(enc, &m.zb, ZsB, GB * (dbc + 4), N, GB * (dbc + 4));
    rc |= pw_make_map(enc, &m.xb, xb, sharedB ? DB : GB * DB, N, ldb);
    QMP_REQUIRE(rc == 0, "qmp_fused_wgrad_tma: cuTensorMapEncodeTiled failed (%d)", rc);
    QMP_REQUIRE(mode == 1 || C == FC || NC == 1, "qmp_fused_wgrad_tma: narrow outputs (C < 32) are supported for single convs only");
    PwArgs a{};
    a.N = N; a.nconv = NC;
    for (int c = 0; c < NC; ++c) {
        PwConv& v = a.c[c];
        const bool segA = c < GA;
        const int g = segA ? c : c - GA;
        v.seg = segA ? 0 : 1;
        v.DC = segA ? dac : dbc;
        v.col_g = mode == 1 ? (segA ? c : (g & 3)) * FC : c * C;
        v.col_x = segA ? 0 : (sharedB ? 0 : g * DB);
        v.col_u = g * (v.DC + 4);
        v.x_hi = (!segA && DB > 32) ? 1 : 0;
        const int total = (v.DC + 2) * v.DC + (v.DC + 4) + FC * (v.DC + 4) + FC * v.DC + FC;
        v.gw = (segA ? gwa : gwb) + (size_t)g * total;
    }
    const int ngroups = cdiv(NC, PW_GROUP);
    const int nstages = cdiv(N, PW_NODES);
    int cpg = n_sm / ngroups;
    if (cpg > nstages) cpg = nstages;
    a.cpg = cpg;
    QMP_CUDA(launch_pdl(panel_wgrad_kernel, dim3(cpg * ngroups), dim3(PW_THREADS), PW_SMEM, (cudaStream_t)stream, m, a));
    QMP_LAUNCH_CHECK("panel_wgrad_kernel");
    return 0;
}

// Weight gradients of ONE wide-input TransformerConv (the decoder's head conv fc_out1, DC = 36): qmp_fused_wgrad_tma for a
// group of one conv with the pointers spelled out.  x [N, ldx] are the conv's input rows (D valid columns), g [N, ldg] the
// gradient of its 32 outputs, Zs / dUs [N, DC + 4] the rows written by qmp_fused_bwd_onepass_tc; accumulates into gw [TOTAL(DC)].
QMP_API int qmp_panel_wgrad(int N, const float* x, int ldx, int D, int DC, const float* g, int ldg, const float* Zs, const float* dUs,
                            float* gw, void* stream) {
    QMP_REQUIRE(DC == (D <= 32 ? 32 : (D + 3) / 4 * 4), "qmp_panel_wgrad: DC must be the cap of D");
    return qmp_fused_wgrad_tma(N, nullptr, 0, 0, 0, x, ldx, D, 1, 1, 0, FC, g, ldg, nullptr, nullptr, Zs, dUs, nullptr, gw, stream);
}
