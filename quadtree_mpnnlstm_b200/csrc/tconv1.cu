// TransformerConv with ONE output channel (the decoder's fc_out2: hidden -> 1, model/seq2seq.py:117-121, 182-187; PyG
// TransformerConv heads = 1, edge_dim = 2, root weight).  With a single channel the query, key and value of a node are
// SCALARS:
//     q_i = Wq.x_i + bq    k_j = Wk.x_j + bk    v_j = Wv.x_j + bv    r_i = Ws.x_i + bs    e_ij = we . ea_ij
//     s_ij = q_i (k_j + e_ij)        alpha = softmax_j(s_ij)        out_i = sum_j drop(alpha_ij) (v_j + e_ij) + r_i
// so the edge phase gathers 16 bytes per neighbour (its [q k v r] record) instead of a 128-byte row, and the dense part
// is a 32 -> 4 projection per node.  The general fused kernels treat this conv as a 32-channel one (u_i = Wk^T q_i rows,
// 32-wide gathers, 128 x 32 tensor-core tiles): 107 us per forecast step for forward + backward + weight gradients at the
// bench mesh.  Here it is four small bandwidth-bound kernels, fp32 SIMT, no tensor cores (nothing GEMM-shaped is left):
//     forward   node kernel (octets: 8 lanes per node row) -> s4 [N,4];   edge kernel (thread = node, online softmax) -> out
//     backward  edge kernel (thread = node: dq_i; dk_j, dv_j by reductions onto the source nodes; d we)  -> ds4 [N,4]
//               node kernel (octets) -> dx = W4^T ds4,  weight gradients  dW4 += ds4 (x) x,  db += ds4
// Packed parameters P [136]: Wq (32) | Wk (32) | Wv (32) | Ws (32) | bq bk bv bs | we0 we1 | 0 0.   The softmax is
// recomputed from s4 in the backward pass (scalars: cheaper than saving logits).  Attention dropout uses the counter-based
// mask of the fused family (fused.cuh: same (seed, edge slot) -> same decision forward and backward).
#include "common.cuh"
#include "fused.cuh"

namespace qmp {

constexpr int T1_D = 32, T1_P = 136;

__device__ __forceinline__ float t1_oct_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}
__device__ __forceinline__ float t1_dot(const float4& a, const float4& b) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))); }

// s4[i] = (q, k, v, r) of node i: 8 lanes per row, one float4 of x and of each weight row per lane
__global__ void __launch_bounds__(256) tconv1_node_fwd_kernel(int N, const float* __restrict__ x, int ldx, const float* __restrict__ P,
                                                              float4* __restrict__ s4) {
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31, o8 = lane >> 3, l8 = lane & 7;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const float4 wq = __ldg(reinterpret_cast<const float4*>(P) + l8), wk = __ldg(reinterpret_cast<const float4*>(P + 32) + l8),
                 wv = __ldg(reinterpret_cast<const float4*>(P + 64) + l8), ws = __ldg(reinterpret_cast<const float4*>(P + 96) + l8);
    const float4 b = __ldg(reinterpret_cast<const float4*>(P + 128));
    for (int ps = warp; ps < (N + 3) / 4; ps += nwarps) {
        const int i = 4 * ps + o8;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < N) xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)i * ldx) + l8);
        const float q = t1_oct_sum(t1_dot(wq, xv)), k = t1_oct_sum(t1_dot(wk, xv)), v = t1_oct_sum(t1_dot(wv, xv)),
                    r = t1_oct_sum(t1_dot(ws, xv));
        if (i < N && l8 == 0) s4[i] = make_float4(q + b.x, k + b.y, v + b.z, r + b.w);
    }
}

// decoder head tail (model/seq2seq.py:167-178, 427-428) on the conv's result y_i: out = tanh(drop(y)) + x0 [-> sigmoid];
// x_next = [out, x[:, 1:]] -- the arithmetic of head_finish_fwd_kernel (lstm.cu), here as the epilogue of the edge kernel
struct T1Finish {
    const float* x; int F, binary; float drop_p; unsigned long long seed;      // step input [N, F], output dropout
    float* out; float* x_next;                                                 // [N], [N, F] (optional)
    const float* y; const float* d_out; const float* d_xnext; float* dx;       // backward: saved y / out, incoming gradients, d x [N, F]
};
__device__ __forceinline__ float t1_finish_keep(unsigned long long seed, int i, float drop_p) {
    if (drop_p <= 0.f) return 1.f;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((float)(z >> 40) * (1.0f / 16777216.0f) >= drop_p) ? 1.f / (1.f - drop_p) : 0.f;
}

// out_i = sum_j drop(alpha_ij) (v_j + e_ij) + r_i over the in-edges, online softmax; thread = node
template <bool FINISH>
__global__ void __launch_bounds__(256) tconv1_edge_fwd_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ nbr,
                                                              const float* __restrict__ ea, const float4* __restrict__ s4,
                                                              const float* __restrict__ P, float* __restrict__ out, float drop_p,
                                                              unsigned long long seed, const unsigned long long* __restrict__ salt,
                                                              const T1Finish fin) {
    pdl_wait();
    pdl_launch();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (drop_p > 0.f) seed = salted_seed(seed, salt);
    const float we0 = __ldg(P + 132), we1 = __ldg(P + 133);
    const float4 si = __ldg(s4 + i);
    float m = -INFINITY, l = 0.f, acc = 0.f;
    const int k1 = __ldg(ptr + i + 1);
    for (int kk = __ldg(ptr + i); kk < k1; ++kk) {
        const float4 sj = __ldg(s4 + __ldg(nbr + kk));
        float e = 0.f;
        if (ea) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(ea) + kk);
            e = fmaf(we0, a.x, we1 * a.y);
        }
        const float s = si.x * (sj.y + e);
        const float mn = fmaxf(m, s);
        const float sc = fast_exp(m - mn), pe = fast_exp(s - mn);
        l = fmaf(l, sc, pe);
        acc = fmaf(acc, sc, pe * fdropout_scale(seed, kk, drop_p) * (sj.z + e));
        m = mn;
    }
    const float y = (l > 0.f ? acc / l : 0.f) + si.w;
    out[i] = y;
    if constexpr (FINISH) {
        const float keep = t1_finish_keep(fin.drop_p > 0.f ? salted_seed(fin.seed, salt) : 0ull, i, fin.drop_p);
        const float* xr = fin.x + (size_t)i * fin.F;
        float o = tanhf(y * keep) + xr[0];
        if (fin.binary) o = 1.f / (1.f + expf(-o));
        fin.out[i] = o;
        if (fin.x_next) {
            float* xn = fin.x_next + (size_t)i * fin.F;
            xn[0] = o;
            for (int c = 1; c < fin.F; ++c) xn[c] = xr[c];
        }
    }
}

// Backward over the in-edges of node i (g_i = d out_i): softmax recomputed; dq_i, dr_i stored; dk_j, dv_j reduced onto the
// sources; d we accumulated per block.  ds4 must be zero on entry.
template <bool FINISH>
__global__ void __launch_bounds__(256) tconv1_edge_bwd_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ nbr,
                                                              const float* __restrict__ ea, const float4* __restrict__ s4,
                                                              const float* __restrict__ P, const float* __restrict__ g,
                                                              float* __restrict__ ds4, float* __restrict__ gP, float drop_p,
                                                              unsigned long long seed, const unsigned long long* __restrict__ salt,
                                                              const T1Finish fin) {
    pdl_wait();
    pdl_launch();
    if (drop_p > 0.f) seed = salted_seed(seed, salt);
    __shared__ float s_we[2];
    if (threadIdx.x < 2) s_we[threadIdx.x] = 0.f;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float dwe0 = 0.f, dwe1 = 0.f;
    if (i < N) {
        const float we0 = __ldg(P + 132), we1 = __ldg(P + 133);
        const float4 si = __ldg(s4 + i);
        float gi;
        if constexpr (FINISH) {                                // backward of the head tail: d y_i and the row d x_i
            const float keep = t1_finish_keep(fin.drop_p > 0.f ? salted_seed(fin.seed, salt) : 0ull, i, fin.drop_p);
            float gg = (fin.d_out ? fin.d_out[i] : 0.f) + (fin.d_xnext ? fin.d_xnext[(size_t)i * fin.F] : 0.f);
            if (fin.binary) gg *= fin.out[i] * (1.f - fin.out[i]);
            const float th = tanhf(fin.y[i] * keep);
            gi = gg * (1.f - th * th) * keep;
            float* dxr = fin.dx + (size_t)i * fin.F;
            dxr[0] = gg;
            for (int f = 1; f < fin.F; ++f) dxr[f] = fin.d_xnext ? fin.d_xnext[(size_t)i * fin.F + f] : 0.f;
        } else {
            gi = __ldg(g + i);
        }
        const int k0 = __ldg(ptr + i), k1 = __ldg(ptr + i + 1);
        float m = -INFINITY, l = 0.f, att = 0.f;               // pass 1: softmax statistics and the attention output
        for (int kk = k0; kk < k1; ++kk) {
            const float4 sj = __ldg(s4 + __ldg(nbr + kk));
            float e = 0.f;
            if (ea) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(ea) + kk);
                e = fmaf(we0, a.x, we1 * a.y);
            }
            const float s = si.x * (sj.y + e);
            const float mn = fmaxf(m, s);
            const float sc = fast_exp(m - mn), pe = fast_exp(s - mn);
            l = fmaf(l, sc, pe);
            att = fmaf(att, sc, pe * fdropout_scale(seed, kk, drop_p) * (sj.z + e));
            m = mn;
        }
        const float li = l > 0.f ? 1.f / l : 0.f;
        const float tsum = gi * att * li;                      // sum_j alpha_ij d alpha_ij
        float dq = 0.f;
        for (int kk = k0; kk < k1; ++kk) {                     // pass 2: gradients
            const int j = __ldg(nbr + kk);
            const float4 sj = __ldg(s4 + j);
            float2 a = make_float2(0.f, 0.f);
            if (ea) a = __ldg(reinterpret_cast<const float2*>(ea) + kk);
            const float e = fmaf(we0, a.x, we1 * a.y);
            const float key = sj.y + e;
            const float al = fast_exp(si.x * key - m) * li, keep = fdropout_scale(seed, kk, drop_p);
            const float dal = gi * keep * (sj.z + e);
            const float dsv = al * (dal - tsum);               // d s_ij
            const float dkey = dsv * si.x, dval = al * keep * gi;
            dq = fmaf(dsv, key, dq);
            atomicAdd(ds4 + (size_t)j * 4 + 1, dkey);
            atomicAdd(ds4 + (size_t)j * 4 + 2, dval);
            dwe0 = fmaf(dkey + dval, a.x, dwe0);
            dwe1 = fmaf(dkey + dval, a.y, dwe1);
        }
        ds4[(size_t)i * 4] = dq;
        ds4[(size_t)i * 4 + 3] = gi;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        dwe0 += __shfl_xor_sync(0xffffffffu, dwe0, d);
        dwe1 += __shfl_xor_sync(0xffffffffu, dwe1, d);
    }
    if ((threadIdx.x & 31) == 0 && (dwe0 != 0.f || dwe1 != 0.f)) {
        atomicAdd(&s_we[0], dwe0);
        atomicAdd(&s_we[1], dwe1);
    }
    __syncthreads();
    if (threadIdx.x < 2 && gP && s_we[threadIdx.x] != 0.f) atomicAdd(gP + 132 + threadIdx.x, s_we[threadIdx.x]);
}

// dx_i = W4^T ds4_i (optional) and the parameter gradients dW4 += ds4 (x) x, db += ds4; octets, accumulators in registers
__global__ void __launch_bounds__(256) tconv1_node_bwd_kernel(int N, const float* __restrict__ x, int ldx, const float* __restrict__ P,
                                                              const float4* __restrict__ ds4, float* __restrict__ dx, int lddx,
                                                              float* __restrict__ gP, int relu_mask) {
    pdl_wait();
    pdl_launch();
    __shared__ float s_g[T1_P];
    for (int t = threadIdx.x; t < T1_P; t += blockDim.x) s_g[t] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, o8 = lane >> 3, l8 = lane & 7;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const float4 wq = __ldg(reinterpret_cast<const float4*>(P) + l8), wk = __ldg(reinterpret_cast<const float4*>(P + 32) + l8),
                 wv = __ldg(reinterpret_cast<const float4*>(P + 64) + l8), ws = __ldg(reinterpret_cast<const float4*>(P + 96) + l8);
    float4 aq = make_float4(0.f, 0.f, 0.f, 0.f), ak = aq, av = aq, as = aq, ab = aq;
    for (int ps = warp; ps < (N + 3) / 4; ps += nwarps) {
        const int i = 4 * ps + o8;
        if (i >= N) continue;
        const float4 d = __ldg(ds4 + i);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)i * ldx) + l8);
        if (dx) {
            float4 o;
            o.x = fmaf(ws.x, d.w, fmaf(wv.x, d.z, fmaf(wk.x, d.y, wq.x * d.x)));
            o.y = fmaf(ws.y, d.w, fmaf(wv.y, d.z, fmaf(wk.y, d.y, wq.y * d.x)));
            o.z = fmaf(ws.z, d.w, fmaf(wv.z, d.z, fmaf(wk.z, d.y, wq.z * d.x)));
            o.w = fmaf(ws.w, d.w, fmaf(wv.w, d.z, fmaf(wk.w, d.y, wq.w * d.x)));
            if (relu_mask) {                               // x = relu(.): the gradient w.r.t. the pre-activation (model/seq2seq.py:184)
                o.x = xv.x > 0.f ? o.x : 0.f; o.y = xv.y > 0.f ? o.y : 0.f; o.z = xv.z > 0.f ? o.z : 0.f; o.w = xv.w > 0.f ? o.w : 0.f;
            }
            *(reinterpret_cast<float4*>(dx + (size_t)i * lddx) + l8) = o;
        }
        aq.x = fmaf(d.x, xv.x, aq.x); aq.y = fmaf(d.x, xv.y, aq.y); aq.z = fmaf(d.x, xv.z, aq.z); aq.w = fmaf(d.x, xv.w, aq.w);
        ak.x = fmaf(d.y, xv.x, ak.x); ak.y = fmaf(d.y, xv.y, ak.y); ak.z = fmaf(d.y, xv.z, ak.z); ak.w = fmaf(d.y, xv.w, ak.w);
        av.x = fmaf(d.z, xv.x, av.x); av.y = fmaf(d.z, xv.y, av.y); av.z = fmaf(d.z, xv.z, av.z); av.w = fmaf(d.z, xv.w, av.w);
        as.x = fmaf(d.w, xv.x, as.x); as.y = fmaf(d.w, xv.y, as.y); as.z = fmaf(d.w, xv.z, as.z); as.w = fmaf(d.w, xv.w, as.w);
        if (l8 == 0) { ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w; }
    }
    if (!gP) return;
    auto fold = [&](float4& v) {                                 // over the 4 octets of the warp
        v.x += __shfl_xor_sync(0xffffffffu, v.x, 8); v.x += __shfl_xor_sync(0xffffffffu, v.x, 16);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, 8); v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, 8); v.z += __shfl_xor_sync(0xffffffffu, v.z, 16);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, 8); v.w += __shfl_xor_sync(0xffffffffu, v.w, 16);
    };
    fold(aq); fold(ak); fold(av); fold(as); fold(ab);
    if (o8 == 0) {
        auto put = [&](int base, const float4& v) {
            atomicAdd(&s_g[base + 4 * l8], v.x); atomicAdd(&s_g[base + 4 * l8 + 1], v.y);
            atomicAdd(&s_g[base + 4 * l8 + 2], v.z); atomicAdd(&s_g[base + 4 * l8 + 3], v.w);
        };
        put(0, aq); put(32, ak); put(64, av); put(96, as);
        if (l8 == 0) {
            atomicAdd(&s_g[128], ab.x); atomicAdd(&s_g[129], ab.y); atomicAdd(&s_g[130], ab.z); atomicAdd(&s_g[131], ab.w);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 132; t += blockDim.x)
        if (s_g[t] != 0.f) atomicAdd(gP + t, s_g[t]);
}

static int t1_grid(int N) {
    const int want = (N + 31) / 32;              // 8 warps x 4 nodes per pass
    return want < 148 * 4 ? (want > 0 ? want : 1) : 148 * 4;
}

}  // namespace qmp
using namespace qmp;

// One-channel TransformerConv forward.  x [N, ldx] (32 columns used, 16-byte aligned rows), P [136] packed parameters (top of
// this file), in-CSR + edge attributes [E, 2] in CSR order (may be NULL).  Writes s4 [N, 4] (saved for the backward pass)
// and out [N].
QMP_API int qmp_tconv1_fwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* x, int ldx,
                           const float* P, float* s4, float* out, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(ldx % 4 == 0 && ldx >= T1_D && al16(x) && al16(P) && al16(s4), "qmp_tconv1_fwd: rows must be 16-byte aligned");
    QMP_REQUIRE(!ea || (reinterpret_cast<uintptr_t>(ea) & 7) == 0, "qmp_tconv1_fwd: edge attributes must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    QMP_CUDA(launch_pdl(tconv1_node_fwd_kernel, dim3(t1_grid(N)), dim3(256), 0, st, N, x, ldx, P, reinterpret_cast<float4*>(s4)));
    QMP_LAUNCH_CHECK("tconv1_node_fwd_kernel");
    QMP_CUDA(launch_pdl(tconv1_edge_fwd_kernel<false>, dim3(cdiv(N, 256)), dim3(256), 0, st, N, in_ptr, in_src, ea, reinterpret_cast<const float4*>(s4), P, out, drop_p, seed, qmp::dropout_salt(), T1Finish{}));
    QMP_LAUNCH_CHECK("tconv1_edge_fwd_kernel");
    return 0;
}

// Backward of qmp_tconv1_fwd: g [N] = d out.  ds4 [N, 4] is scratch.  Writes dx [N, lddx] (32 columns; may be NULL) and
// ACCUMULATES the parameter gradients into gP [136] (may be NULL).
QMP_API int qmp_tconv1_bwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* x, int ldx,
                           const float* P, const float* s4, const float* g, float* ds4, float* dx, int lddx, float* gP,
                           float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(ldx % 4 == 0 && ldx >= T1_D && al16(x) && al16(P) && al16(s4) && al16(ds4) && (!dx || (al16(dx) && lddx % 4 == 0)),
                "qmp_tconv1_bwd: rows must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    QMP_CUDA(cudaMemsetAsync(ds4, 0, (size_t)N * 4 * sizeof(float), st));
    QMP_CUDA(launch_pdl(tconv1_edge_bwd_kernel<false>, dim3(cdiv(N, 256)), dim3(256), 0, st, N, in_ptr, in_src, ea, reinterpret_cast<const float4*>(s4), P, g, ds4, gP,
                                                                drop_p, seed, qmp::dropout_salt(), T1Finish{}));
    QMP_LAUNCH_CHECK("tconv1_edge_bwd_kernel");
    QMP_CUDA(launch_pdl(tconv1_node_bwd_kernel, dim3(t1_grid(N) < 296 ? t1_grid(N) : 296), dim3(256), 0, st, N, x, ldx, P, reinterpret_cast<const float4*>(ds4), dx,
                                                                               lddx, gP, 0));
    QMP_LAUNCH_CHECK("tconv1_node_bwd_kernel");
    return 0;
}

// The whole tail of the decoder head in two launches: y = TransformerConv(32 -> 1)(h) (qmp_tconv1_fwd), then -- in the same edge
// kernel -- out = tanh(dropout(y)) + x[:, 0] (-> sigmoid if binary) and x_next = [out, x[:, 1:]] (qmp_head_finish_fwd;
// model/seq2seq.py:167-187, 427-428).  h [N, ldh] (32 columns), x [N, F] the step's input; writes s4 [N, 4], y [N] (both saved
// for the backward pass), out [N], x_next [N, F] (may be NULL).
QMP_API int qmp_head_tail_fwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* h, int ldh, const float* P,
                              const float* x, int F, int binary, float drop_attn, unsigned long long seed_attn, float drop_out,
                              unsigned long long seed_out, float* s4, float* y, float* out, float* x_next, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(ldh % 4 == 0 && ldh >= T1_D && al16(h) && al16(P) && al16(s4), "qmp_head_tail_fwd: rows must be 16-byte aligned");
    QMP_REQUIRE(!ea || (reinterpret_cast<uintptr_t>(ea) & 7) == 0, "qmp_head_tail_fwd: edge attributes must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    QMP_CUDA(launch_pdl(tconv1_node_fwd_kernel, dim3(t1_grid(N)), dim3(256), 0, st, N, h, ldh, P, reinterpret_cast<float4*>(s4)));
    QMP_LAUNCH_CHECK("tconv1_node_fwd_kernel");
    T1Finish fin{};
    fin.x = x; fin.F = F; fin.binary = binary; fin.drop_p = drop_out; fin.seed = seed_out; fin.out = out; fin.x_next = x_next;
    QMP_CUDA(launch_pdl(tconv1_edge_fwd_kernel<true>, dim3(cdiv(N, 256)), dim3(256), 0, st, N, in_ptr, in_src, ea, reinterpret_cast<const float4*>(s4), P, y, drop_attn,
                                                               seed_attn, qmp::dropout_salt(), fin));
    QMP_LAUNCH_CHECK("tconv1_edge_fwd_kernel");
    return 0;
}

// Backward of qmp_head_tail_fwd in two launches (+ one memset): the edge kernel first forms d y_i and the row d x_i [F] from
// d_out [N] / d_xnext [N, F] (either may be NULL) as qmp_head_finish_bwd does, then runs the conv's edge backward; the node
// kernel writes dh [N, lddh] -- masked by h > 0 when relu_mask != 0, i.e. the gradient w.r.t. the pre-activation of
// h = relu(.) (model/seq2seq.py:184) -- and accumulates the parameter gradients into gP [136] (may be NULL).  ds4 [N, 4] scratch.
QMP_API int qmp_head_tail_bwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* h, int ldh, const float* P,
                              const float* s4, const float* y, const float* out, const float* x, int F, int binary, float drop_attn,
                              unsigned long long seed_attn, float drop_out, unsigned long long seed_out, const float* d_out,
                              const float* d_xnext, float* ds4, float* dh, int lddh, int relu_mask, float* dx, float* gP, void* stream) {
    if (N <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    QMP_REQUIRE(ldh % 4 == 0 && ldh >= T1_D && al16(h) && al16(P) && al16(s4) && al16(ds4) && (!dh || (al16(dh) && lddh % 4 == 0)),
                "qmp_head_tail_bwd: rows must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    QMP_CUDA(cudaMemsetAsync(ds4, 0, (size_t)N * 4 * sizeof(float), st));
    T1Finish fin{};
    fin.x = x; fin.F = F; fin.binary = binary; fin.drop_p = drop_out; fin.seed = seed_out; fin.out = const_cast<float*>(out); fin.y = y;
    fin.d_out = d_out; fin.d_xnext = d_xnext; fin.dx = dx;
    QMP_CUDA(launch_pdl(tconv1_edge_bwd_kernel<true>, dim3(cdiv(N, 256)), dim3(256), 0, st, N, in_ptr, in_src, ea, reinterpret_cast<const float4*>(s4), P, nullptr, ds4,
                                                               gP, drop_attn, seed_attn, qmp::dropout_salt(), fin));
    QMP_LAUNCH_CHECK("tconv1_edge_bwd_kernel");
    QMP_CUDA(launch_pdl(tconv1_node_bwd_kernel, dim3(t1_grid(N) < 296 ? t1_grid(N) : 296), dim3(256), 0, st, N, h, ldh, P, reinterpret_cast<const float4*>(ds4), dh,
                                                                               lddh, gP, relu_mask));
    QMP_LAUNCH_CHECK("tconv1_node_bwd_kernel");
    return 0;
}
