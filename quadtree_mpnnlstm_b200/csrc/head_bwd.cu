// Backward of the decoder's head conv fc_out1 -- ONE TransformerConv on the 36-wide head rows (32 normalised outputs | concat
// layer | 3 pads; model/seq2seq.py:117-121, 182-187), 90 launches per sample -- in the mapping of fused_cell_bwd.cu: one
// persistent CTA per SM, dense contractions on tcgen05 (3xTF32, A operands in tensor memory, thread = TMEM lane = node), the edge
// phase in OCTET layout (8 lanes per node row, a float4 of the 32 main columns per lane, lane 0 of the octet also carries the four
// extra columns), the source side of every edge in the same pass by 16-byte vector reductions.  It replaces the thread-per-node
// one-pass kernel fused_bwd_tc_kernel<0,36,1> for this conv (54 us per forecast step: 8 warps per SM, 369 tiles on 296 CTA
// slots = two rounds); same inputs, outputs and dropout mask, so the forward kernel and the weight-gradient kernel are unchanged.
//
//   G0  u  = x W1^T          (N = 48: 36 logit projections | 2 edge-attribute weights | pad)       recomputed, not saved
//   G1  dz = g W2, dx = g W3 (dz: 36 | dze0 dze1 dzs | pad)                                         g = gradient of the conv's output
//   edge phase per target i: alpha from the saved logits; d alpha_e = (dz . [x_j | e | 1]) keep_e; ds_e = alpha_e (d alpha_e - t);
//       du = sum ds x_j, dw = sum ds e, z = sum alpha keep [x_j | e | 1]  (rows dUs / Zs for the weight gradients);
//       source side: dx_j += ds_e u_i + alpha_e keep_e dz_i     (red.global.add.v4.f32)
//   G2  dx += [du | dw] W1
// Reference: autograd of PyG TransformerConv (heads 1, edge_dim 2, root weight) as convs.pack_tconv folds it.
#include "fused_fwd.inl"
#include "fused_cell.cuh"

namespace qmp {

struct HeadBwdLayout {
    static constexpr int DC = 36, KX = 40, NP = 48;               // row width, its K padding, N of every contraction (multiple of 16)
    static constexpr int BU = 2 * NP * KX * 4;                     // W1   as [n = r][k]: u_r = sum_k x_k W1[r][k]       (hi then lo)
    static constexpr int BZS = 2 * 2 * NP * FC * 4;                // [W2^T ; W3^T] as ONE operand of 96 rows [n][k = o]: dz_n = sum_o g_o W2[o][n]
                                                                   // (rows 0..47), dx_n = sum_o g_o W3[o][n] (rows 48..95) -- one chain, N = 96
    static constexpr int BD = 2 * NP * KX * 4;                     // W1^T as [n = k][r]: dx_k += sum_r dU_r W1[r][k]
    static constexpr int OU = 0, OZS = OU + BU, OD = OZS + BZS, OB1 = OD + BD;
    static constexpr int BYTES = OB1 + KX * 4;                     // + b1 (40 floats)
};
static_assert(HeadBwdLayout::BYTES % 16 == 0, "bulk copies move 16-byte units");

constexpr uint32_t HB_AG = 0;          // g rows: hi 32 | lo 32
constexpr uint32_t HB_AX = 64;         // x rows: hi 40 | lo 40; later [du | dw]
constexpr uint32_t HB_DU = 144;        // u, 48 columns
constexpr uint32_t HB_DZ = 192;        // dz, 48 columns
constexpr uint32_t HB_DX = 240;        // dx, 48 columns (= HB_DZ + 48: dz and the skip part of dx come out of one N = 96 chain)
static_assert(HB_DX == HB_DZ + 48, "one contraction writes dz | dx");
constexpr size_t HEADB_SMEM = HeadBwdLayout::BYTES + 2 * XPLANE * sizeof(float);
constexpr int HEADB_THREADS = CELL_WORKERS;

struct HeadBwdArgs {
    int N;
    const int* ptr; const int* nbr; const float* ea;
    const float* x; int ldx;                                   // [N, ldx >= 36] head rows
    const float* g; int ldg;                                   // [N, ldg >= 32] gradient of the conv's 32 outputs
    const float* logit; const float* mstat; const float* linv; // [E], [N], [N]
    float* Zs; float* dUs;                                     // [N, 40] rows for the weight gradients
    float* dx;                                                 // [N, ldx]: zero on entry
    float drop_p; unsigned long long seed; const unsigned long long* salt;
};

__device__ __forceinline__ void headb_sync() { asm volatile("bar.sync 0, %0;" ::"n"(HEADB_THREADS) : "memory"); }
__device__ __forceinline__ void headb_red4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void headb_fma4(float4& acc, float s, const float4& v) {
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
// sum over the 8 lanes of an octet of four values per lane: lane l8 ends with the total of v[l8 >> 1]
__device__ __forceinline__ float octet_reduce4(const float (&v)[4], int l8) {
    const bool b2 = l8 & 4, b1 = l8 & 2;
    float r2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b2 ? v[i + 2] : v[i], send = b2 ? v[i] : v[i + 2];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float keep = b1 ? r2[1] : r2[0], send = b1 ? r2[0] : r2[1];
    const float r1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    return r1 + __shfl_xor_sync(0xffffffffu, r1, 1);
}
// sum over the four (edge) lane pairs of an octet: lanes that differ in bits 1, 2 (the bit-0 partner holds the same value)
__device__ __forceinline__ float edge_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

__device__ __forceinline__ void headb_prefetch_tile(const HeadBwdArgs& a, int n0, int cnt) {
    const size_t n = (size_t)n0;
    tc::l2_prefetch(a.g + n * a.ldg, (long long)cnt * a.ldg * 4);
    tc::l2_prefetch(a.x + n * a.ldx, (long long)cnt * a.ldx * 4);
    tc::l2_prefetch(a.mstat + n, (long long)cnt * 4);
    tc::l2_prefetch(a.linv + n, (long long)cnt * 4);
    const int k0 = __ldg(a.ptr + n0), k1 = __ldg(a.ptr + n0 + cnt);
    tc::l2_prefetch(a.logit + k0, (long long)(k1 - k0) * 4);
    tc::l2_prefetch(a.nbr + k0, (long long)(k1 - k0) * 4);
    if (a.ea) tc::l2_prefetch(a.ea + (size_t)k0 * 2, (long long)(k1 - k0) * 8);
}

__global__ void __launch_bounds__(HEADB_THREADS, 1) head_bwd_kernel(const __grid_constant__ HeadBwdArgs a, const uint8_t* __restrict__ img,
                                                                     const int Q, const int R, const int T0) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[2];                 // 0: MMA groups, 1: image landed
    __shared__ uint32_t tmem_slot;
    using L = HeadBwdLayout;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    qmp_seed_init(a.seed, a.salt);
    float* exch = reinterpret_cast<float*>(smem + L::BYTES);          // plane 0: u, later [du | dw]; plane 1: dz
    if (t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_mbar_init();
        tc::mbar_expect_tx(&bars[1], (uint32_t)L::BYTES);      // the weight image starts moving before the tensor-memory allocation
        for (int off = 0; off < L::BYTES; off += 16384)
            tc::bulk_g2s(smem + off, img + off, (uint32_t)(L::BYTES - off < 16384 ? L::BYTES - off : 16384), &bars[1]);
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    pdl_wait();
    pdl_launch();
    tc::fence_before_sync();
    headb_sync();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int beg = (int)blockIdx.x * Q;
    int end = beg + Q;
    if (end > a.N) end = a.N;
    if (t == 64 && beg < end) headb_prefetch_tile(a, beg, end - beg < T0 ? end - beg : T0);

    const int q = warp & 3, cg = warp >> 2;
    const int nrow = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int o8 = lane >> 3, l8 = lane & 7, obase = lane & ~7, el = l8 >> 1;
    uint32_t par = 0;
    for (int r = 0; r < R; ++r) {
        const int tile0 = beg + r * T0;
        if (tile0 >= end) break;
        const int tcount = (end - tile0 < T0) ? end - tile0 : T0;
        if (t == 64 && tile0 + T0 < end) headb_prefetch_tile(a, tile0 + T0, end - tile0 - T0 < T0 ? end - tile0 - T0 : T0);

        // ---- g and x rows of node nrow -> tensor memory (column group cg: 8 columns; cg 0 also the four extra columns of x)
        {
            const int i = tile0 + nrow;
            const bool valid = nrow < tcount;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* gp = a.g + (size_t)i * a.ldg + 8 * cg;
            const float* xp = a.x + (size_t)i * a.ldx + 8 * cg;
            float4 g0 = zero, g1 = zero, x0 = zero, x1 = zero, x2 = zero;
            if (valid) {
                g0 = __ldg(reinterpret_cast<const float4*>(gp));
                g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
                x0 = __ldg(reinterpret_cast<const float4*>(xp));
                x1 = __ldg(reinterpret_cast<const float4*>(xp) + 1);
                if (cg == 0) x2 = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)i * a.ldx + 32));
            }
            float v[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            cell_stage8(lane_addr + HB_AG + 8 * (uint32_t)cg, lane_addr + HB_AG + 32 + 8 * (uint32_t)cg, v);
            float w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            cell_stage8(lane_addr + HB_AX + 8 * (uint32_t)cg, lane_addr + HB_AX + 40 + 8 * (uint32_t)cg, w);
            if (cg == 0) {
                float e[8] = {x2.x, x2.y, x2.z, x2.w, 0.f, 0.f, 0.f, 0.f};
                cell_stage8(lane_addr + HB_AX + 32, lane_addr + HB_AX + 72, e);
            }
            tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        headb_sync();
        if (t == 0) {                                          // G0 + G1: u, dz, skip part of dx
            tc::mbar_wait(&bars[1], 0);                        // weights in shared memory (returns at once after the first tile)
            tc::fence_after_sync();
            tc_mma3_at(0, tmem + HB_DU, tmem + HB_AX, tmem + HB_AX + 40, tc::smem_u32(smem + L::OU), tc::smem_u32(smem + L::OU + L::BU / 2),
                       L::NP, L::KX, false);
            tc_mma3_at(0, tmem + HB_DZ, tmem + HB_AG, tmem + HB_AG + 32, tc::smem_u32(smem + L::OZS), tc::smem_u32(smem + L::OZS + L::BZS / 2),
                       2 * L::NP, FC, false);                  // dz (columns HB_DZ ..) and the skip part of dx (HB_DX = HB_DZ + 48) in one chain
            tc::commit(&bars[0]);
        }
        __syncwarp();
        tc::mbar_wait(&bars[1], 0);                            // b1 is read below
        tc::mbar_wait(&bars[0], par);
        par ^= 1;
        tc::fence_after_sync();
        // ---- u (+ b1) -> exchange plane 0, dz -> plane 1: row nrow, columns 8 cg .. (cg 0 also 32..39)
        {
            const float* b1 = reinterpret_cast<const float*>(smem + L::OB1);
            float* ru = exch + nrow * XS;
            float* rz = exch + XPLANE + nrow * XS;
            float u8[8], z8[8];
            tc::tmem_ld8(lane_addr + HB_DU + 8 * (uint32_t)cg, u8);
            tc::tmem_ld8(lane_addr + HB_DZ + 8 * (uint32_t)cg, z8);
            const float4 ba = ld4(b1 + 8 * cg), bb = ld4(b1 + 8 * cg + 4);
            st4(ru + 8 * cg, u8[0] + ba.x, u8[1] + ba.y, u8[2] + ba.z, u8[3] + ba.w);
            st4(ru + 8 * cg + 4, u8[4] + bb.x, u8[5] + bb.y, u8[6] + bb.z, u8[7] + bb.w);
            st4(rz + 8 * cg, z8[0], z8[1], z8[2], z8[3]);
            st4(rz + 8 * cg + 4, z8[4], z8[5], z8[6], z8[7]);
            if (cg == 0) {
                tc::tmem_ld8(lane_addr + HB_DU + 32, u8);
                tc::tmem_ld8(lane_addr + HB_DZ + 32, z8);
                const float4 bc = ld4(b1 + 32), bd = ld4(b1 + 36);
                st4(ru + 32, u8[0] + bc.x, u8[1] + bc.y, u8[2] + bc.z, u8[3] + bc.w);
                st4(ru + 36, u8[4] + bd.x, u8[5] + bd.y, u8[6] + bd.z, u8[7] + bd.w);
                st4(rz + 32, z8[0], z8[1], z8[2], z8[3]);
                st4(rz + 36, z8[4], z8[5], z8[6], z8[7]);
            }
        }
        headb_sync();

        // ---- edge phase, octet layout, 4 nodes per warp pass
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            if (4 * (warp + 16 * p) >= tcount) continue;       // warp-uniform
            const int ln = 4 * (warp + 16 * p) + o8;
            const bool valid = ln < tcount;
            const int i = tile0 + ln;
            const int k0 = valid ? __ldg(a.ptr + i) : 0;
            const int deg = valid ? __ldg(a.ptr + i + 1) - k0 : 0;
            const int nq = __reduce_max_sync(0xffffffffu, (deg + 3) >> 2);
            float* r0 = exch + ln * XS;
            const float* r1 = exch + XPLANE + ln * XS;
            const float4 u = ld4(r0 + 4 * l8), dz = ld4(r1 + 4 * l8);
            const float4 ux = ld4(r0 + 32), dzx = ld4(r1 + 32), dzt = ld4(r1 + 36);      // extra columns; (dze0 dze1 dzs 0)
            const float m = valid ? __ldg(a.mstat + i) : 0.f, li = valid ? __ldg(a.linv + i) : 0.f;
            auto gather = [&](int qd, bool& on, int& kk, float& lg, float2& ev, float4 (&hr)[4], float4 (&hx)[4], int (&jx)[4]) {
                on = 4 * qd + el < deg;
                kk = k0 + 4 * qd + el;
                const int jj = on ? __ldg(a.nbr + kk) : -1;
                ev = make_float2(0.f, 0.f);
                if (on && a.ea) ev = __ldg(reinterpret_cast<const float2*>(a.ea) + kk);
                lg = on ? __ldg(a.logit + kk) : 0.f;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    jx[x] = __shfl_sync(0xffffffffu, jj, obase + 2 * x);
                    hr[x] = hx[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (jx[x] >= 0) {
                        hr[x] = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)jx[x] * a.ldx) + l8);
                        if (l8 == 0) hx[x] = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)jx[x] * a.ldx) + 8);
                    }
                }
            };
            auto coef = [&](const float4 (&hr)[4], const float4 (&hx)[4], bool on, int kk, float lg, const float2& ev, float& al, float& keep) {
                float v[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = dot4(dz, hr[x]) + dot4(dzx, hx[x]);      // hx is zero off lane 0 of the octet
                const float tot = octet_reduce4(v, l8);
                al = 0.f;
                keep = 0.f;
                if (on) {
                    al = fast_exp(lg - m) * li;
                    keep = fdropout_scale(QMP_SEED_SM, (long long)kk, a.drop_p);
                }
                return (tot + fmaf(dzt.x, ev.x, fmaf(dzt.y, ev.y, dzt.z))) * keep;
            };
            // first quad: kept in registers across both passes
            bool on0; int kk0, jx0[4]; float lg0; float2 ev0; float4 hr0[4], hx0[4];
            gather(0, on0, kk0, lg0, ev0, hr0, hx0, jx0);
            float al0, keep0;
            const float dal0 = coef(hr0, hx0, on0, kk0, lg0, ev0, al0, keep0);
            float tsum = edge_sum(al0 * dal0);
            for (int qd = 1; qd < nq; ++qd) {                  // larger in-degrees (quadtree meshes)
                bool on; int kk, jx[4]; float lg; float2 ev; float4 hr[4], hx[4];
                gather(qd, on, kk, lg, ev, hr, hx, jx);
                float al, keep;
                const float dal = coef(hr, hx, on, kk, lg, ev, al, keep);
                tsum += edge_sum(al * dal);
            }
            float4 du = make_float4(0.f, 0.f, 0.f, 0.f), z = du, dux = du, zx = du;
            float dw0 = 0.f, dw1 = 0.f, ze0 = 0.f, ze1 = 0.f, zs = 0.f;
            auto accumulate = [&](const float4 (&hr)[4], const float4 (&hx)[4], const int (&jx)[4], float al, float keep, float dal,
                                  const float2& ev) {
                const float dsv = al * (dal - tsum), alk = al * keep;
                dw0 += edge_sum(dsv * ev.x); dw1 += edge_sum(dsv * ev.y);
                ze0 += edge_sum(alk * ev.x); ze1 += edge_sum(alk * ev.y); zs += edge_sum(alk);
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const float dsb = __shfl_sync(0xffffffffu, dsv, obase + 2 * x);
                    const float alb = __shfl_sync(0xffffffffu, alk, obase + 2 * x);
                    headb_fma4(du, dsb, hr[x]);
                    headb_fma4(z, alb, hr[x]);
                    headb_fma4(dux, dsb, hx[x]);
                    headb_fma4(zx, alb, hx[x]);
                    if (jx[x] >= 0) {                          // source side: dx_j += ds u_i + alpha dz_i
                        float4 con = make_float4(0.f, 0.f, 0.f, 0.f);
                        headb_fma4(con, dsb, u);
                        headb_fma4(con, alb, dz);
                        headb_red4(a.dx + (size_t)jx[x] * a.ldx + 4 * l8, con);
                        if (l8 == 0) {
                            float4 cx = make_float4(0.f, 0.f, 0.f, 0.f);
                            headb_fma4(cx, dsb, ux);
                            headb_fma4(cx, alb, dzx);
                            headb_red4(a.dx + (size_t)jx[x] * a.ldx + 32, cx);
                        }
                    }
                }
            };
            accumulate(hr0, hx0, jx0, al0, keep0, dal0, ev0);
            for (int qd = 1; qd < nq; ++qd) {
                bool on; int kk, jx[4]; float lg; float2 ev; float4 hr[4], hx[4];
                gather(qd, on, kk, lg, ev, hr, hx, jx);
                float al, keep;
                const float dal = coef(hr, hx, on, kk, lg, ev, al, keep);
                accumulate(hr, hx, jx, al, keep, dal, ev);
            }
            __syncwarp();                                      // every lane of the octet has read u of this row
            st4(r0 + 4 * l8, du.x, du.y, du.z, du.w);          // [du | dw] for the last contraction
            if (l8 == 0) {
                st4(r0 + 32, dux.x, dux.y, dux.z, dux.w);
                st4(r0 + 36, dw0, dw1, 0.f, 0.f);
            }
            if (valid) {                                       // rows for the weight-gradient kernel
                float* zr = a.Zs + (size_t)i * 40;
                float* dr = a.dUs + (size_t)i * 40;
                st4(zr + 4 * l8, z.x, z.y, z.z, z.w);
                st4(dr + 4 * l8, du.x, du.y, du.z, du.w);
                if (l8 == 0) {
                    st4(zr + 32, zx.x, zx.y, zx.z, zx.w);
                    st4(zr + 36, ze0, ze1, zs, 0.f);
                    st4(dr + 32, dux.x, dux.y, dux.z, dux.w);
                    st4(dr + 36, dw0, dw1, 0.f, 0.f);
                }
            }
        }
        headb_sync();

        // ---- [du | dw] of row nrow -> tensor memory (K = 40), last contraction
        {
            const float* row = exch + nrow * XS;
            const bool valid = nrow < tcount;
            float v[8];
            ld8(v, row + 8 * cg);
            if (!valid) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = 0.f;
            }
            cell_stage8(lane_addr + HB_AX + 8 * (uint32_t)cg, lane_addr + HB_AX + 40 + 8 * (uint32_t)cg, v);
            if (cg == 0) {
                ld8(v, row + 32);
                if (!valid) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = 0.f;
                }
                cell_stage8(lane_addr + HB_AX + 32, lane_addr + HB_AX + 72, v);
            }
            tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        headb_sync();
        if (t == 0) {                                          // G2: dx += [du | dw] W1
            tc::fence_after_sync();
            tc_mma3_at(0, tmem + HB_DX, tmem + HB_AX, tmem + HB_AX + 40, tc::smem_u32(smem + L::OD), tc::smem_u32(smem + L::OD + L::BD / 2),
                       L::NP, L::KX, true);
            tc::commit(&bars[0]);
        }
        __syncwarp();
        tc::mbar_wait(&bars[0], par);
        par ^= 1;
        tc::fence_after_sync();
        // ---- self term dx_i: 8 columns per thread (cg 0 also the four extra columns), by reductions like the source-side terms
        {
            const int i = tile0 + nrow;
            const bool valid = nrow < tcount;
            float v[8], w[8];
            tc::tmem_ld8(lane_addr + HB_DX + 8 * (uint32_t)cg, v);
            if (cg == 0) tc::tmem_ld8(lane_addr + HB_DX + 32, w);
            if (valid) {
                float* d = a.dx + (size_t)i * a.ldx + 8 * cg;
                headb_red4(d, make_float4(v[0], v[1], v[2], v[3]));
                headb_red4(d + 4, make_float4(v[4], v[5], v[6], v[7]));
                if (cg == 0) headb_red4(a.dx + (size_t)i * a.ldx + 32, make_float4(w[0], w[1], w[2], w[3]));
            }
        }
        tc::fence_before_sync();
        headb_sync();                                          // exchange planes and tensor memory free for the next tile
    }
    tc::fence_before_sync();
    headb_sync();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

__device__ __forceinline__ void headb_put(uint8_t* img, int off, int half, int n, int k, int K, float v) {
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    *reinterpret_cast<float*>(img + off + img_off(n, k, K)) = hi;
    *reinterpret_cast<float*>(img + off + half + img_off(n, k, K)) = lo;
}

__global__ void __launch_bounds__(256) pack_head_bwd_kernel(const float* __restrict__ pack, uint8_t* __restrict__ img) {
    using L = HeadBwdLayout;
    using S = ConvSizes<36>;
    const int tid = blockIdx.x * 256 + threadIdx.x, nth = gridDim.x * 256;
    const float* W1 = pack;                                    // [38][36]
    const float* b1 = pack + S::W1;                            // [40]
    const float* W2 = b1 + S::B1;                              // [32][40]
    const float* W3 = W2 + S::W2;                              // [32][36]
    for (int idx = tid; idx < L::NP * L::KX; idx += nth) {     // W1 as [n = r][k] and W1^T as [n = k][r]
        const int n = idx / L::KX, k = idx % L::KX;
        headb_put(img, L::OU, L::BU / 2, n, k, L::KX, (n < 38 && k < 36) ? W1[n * 36 + k] : 0.f);
        headb_put(img, L::OD, L::BD / 2, n, k, L::KX, (n < 36 && k < 38) ? W1[k * 36 + n] : 0.f);
    }
    for (int idx = tid; idx < L::NP * FC; idx += nth) {        // W2^T, W3^T as [n][k = o]
        const int n = idx / FC, k = idx % FC;
        headb_put(img, L::OZS, L::BZS / 2, n, k, FC, n < 40 ? W2[k * 40 + n] : 0.f);
        headb_put(img, L::OZS, L::BZS / 2, L::NP + n, k, FC, n < 36 ? W3[k * 36 + n] : 0.f);
    }
    float* b = reinterpret_cast<float*>(img + L::OB1);
    for (int idx = tid; idx < L::KX; idx += nth) b[idx] = idx < 38 ? b1[idx] : 0.f;
}

}  // namespace qmp
using namespace qmp;

// Bytes of the head-conv backward weight image.
QMP_API long long qmp_head_bwd_image_bytes(void) { return HeadBwdLayout::BYTES; }

// pack [TOTAL(36)] (forward pack of the head conv, fused.cuh layout) -> out [qmp_head_bwd_image_bytes()]
QMP_API int qmp_pack_head_bwd(const float* pack, void* out, void* stream) {
    pack_head_bwd_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(pack, (uint8_t*)out);
    QMP_LAUNCH_CHECK("pack_head_bwd_kernel");
    qmp::after_producer();
    return 0;
}

// Backward of the head conv fc_out1 (one TransformerConv, 36-wide rows, 32 outputs): same contract as qmp_fused_bwd_onepass_tc
// for that group -- g [N, ldg] the gradient of the conv's output (relu mask already applied), logit [E] / mstat / linv [N] from
// the forward kernel; writes Zs / dUs [N, 40] (rows for qmp_fused_wgrad_tma) and dx [N, ldx] (zeroed here, accumulated with
// reductions).
QMP_API int qmp_head_bwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* x, int ldx, const void* image,
                         const float* g, int ldg, const float* logit, const float* mstat, const float* linv, float* Zs, float* dUs,
                         float* dx, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(ldx % 4 == 0 && ldx >= 36 && ldg % 4 == 0 && ldg >= 32 && al16(x) && al16(image) && al16(g) && al16(Zs) && al16(dUs) && al16(dx),
                "qmp_head_bwd: rows must be 16-byte aligned");
    QMP_REQUIRE(!ea || (reinterpret_cast<uintptr_t>(ea) & 7) == 0, "qmp_head_bwd: edge attributes must be 8-byte aligned");
    HeadBwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.x = x; a.ldx = ldx; a.g = g; a.ldg = ldg; a.logit = logit; a.mstat = mstat;
    a.linv = linv; a.Zs = Zs; a.dUs = dUs; a.dx = dx; a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QMP_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEADB_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    QMP_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * ldx * sizeof(float), st));
    const int G = cdiv(N, 128) < n_sm ? cdiv(N, 128) : n_sm;
    const int Q = (cdiv(N, G) + 3) & ~3;
    const int R = cdiv(Q, 128);
    const int T0 = (cdiv(Q, R) + 3) & ~3;
    QMP_CUDA(launch_pdl(head_bwd_kernel, dim3(cdiv(N, Q)), dim3(HEADB_THREADS), HEADB_SMEM, st, a, reinterpret_cast<const uint8_t*>(image), Q, R, T0));
    QMP_LAUNCH_CHECK("head_bwd_kernel");
    return 0;
}
