// K2 / K3: pooling pixels into mesh nodes and back, driven by the per-pixel label image instead of
// the reference's dense one-hot [N, P] matrix (reference model/graph_functions.py:391-419 flatten,
// :451-468 unflatten, :383-389 / :460-468 pixel-wise variants, :555-587 get_mapping).
//
// A mesh is stored as: labels int32 [P] (-1 = masked), pix_ptr int32 [N+1] / pix_idx int32 [Pvalid]
// (the pixels of every node in raster order) and npix float [N].
//   pool   : out[b, v, c] = (sum over the node's pixels in raster order of img[b, p, c]) / npix[v]
//   unpool : img[b, p, c] = labels[p] >= 0 ? data[b, labels[p], c] : fill
// The backward of one is the other (with / without the division), so two kernels serve all four.
// The summation order of the pool is DEFINED (segment_sum_kernel below) and oracle/graph_ref.py:pool follows it bit for bit,
// which keeps the node positions -- and therefore the atan2 branch of the edge angle -- identical.
#include "common.cuh"

namespace qmp {

// pixels of each quadtree leaf in raster order, from the leaf rectangles
__global__ void pix_csr_from_rects_kernel(const int* __restrict__ labels, int P, int m, const int4* __restrict__ rect,
                                          const int* __restrict__ pix_ptr, int* __restrict__ pix_idx) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int v = labels[p];
    if (v < 0) return;
    const int4 rc = rect[v];
    const int r = p / m, c = p % m;
    pix_idx[pix_ptr[v] + (r - rc.x) * rc.w + (c - rc.y)] = p;
}

__global__ void npix_to_int_kernel(const float* __restrict__ npix, const int* __restrict__ n_nodes, int cap,
                                   int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    out[i] = (i < *n_nodes) ? (int)npix[i] : 0;
}

// pixel-wise mesh: node = raster rank of the unmasked pixel (graph_functions.py:511)
__global__ void keep_flags_kernel(const uint8_t* __restrict__ mask, int P, int* __restrict__ keep) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) keep[p] = (mask && mask[p]) ? 0 : 1;
}
__global__ void pixelwise_fill_kernel(const int* __restrict__ keep, const int* __restrict__ rank, int P,
                                      int* __restrict__ labels, int* __restrict__ pix_idx, int* __restrict__ pix_ptr,
                                      float* __restrict__ npix) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    if (p == P) {  // closing entry: pix_ptr[N] = N
        const int tot = (P > 0) ? rank[P - 1] + keep[P - 1] : 0;
        pix_ptr[tot] = tot;
        return;
    }
    if (keep[p]) {
        const int v = rank[p];
        labels[p] = v;
        pix_idx[v] = p;
        pix_ptr[v] = v;
        npix[v] = 1.0f;
    } else {
        labels[p] = -1;
    }
}

// out[b, v, c] = sum of img[b, p, c] over the node's pixels (/ npix when divide).  N is read from the device when
// n_nodes_dev != nullptr (capacity launch of the graph build).
//
// Summation order (the definition oracle/graph_ref.py:lane_tree_segment_sum restates, so that node positions -- and with
// them the edge attributes -- stay bit-identical): 32 partial sums, partial l = the node's pixels l, l + 32, l + 64, ...
// (raster order) added one after the other starting from 0; then the butterfly partial[l] += partial[l ^ off] for
// off = 16, 8, 4, 2, 1.  One warp per (frame, node): a 64 x 64 leaf is 128 rounds of coalesced loads instead of 4096
// dependent additions by one thread (693 us of the 977 us a mesh build at the ice grid took with a thread per output).
constexpr int SEG_CH = 8;          // channels per pass
__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ img, int B, int P, int C,
                                                          const int* __restrict__ pix_ptr, const int* __restrict__ pix_idx,
                                                          const float* __restrict__ npix, int n_cap,
                                                          const int* __restrict__ n_nodes_dev, int divide,
                                                          float* __restrict__ out) {
    const int n_nodes = n_nodes_dev ? min(*n_nodes_dev, n_cap) : n_cap;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = warp; w < (long long)B * n_nodes; w += nwarps) {
        const int b = (int)(w / n_nodes), v = (int)(w - (long long)b * n_nodes);
        const int a0 = pix_ptr[v], a1 = pix_ptr[v + 1];
        const float* src = img + (size_t)b * P * C;
        for (int c0 = 0; c0 < C; c0 += SEG_CH) {
            const int cw = C - c0 < SEG_CH ? C - c0 : SEG_CH;
            float acc[SEG_CH];
#pragma unroll
            for (int c = 0; c < SEG_CH; ++c) acc[c] = 0.f;
            // four rounds of loads in flight (pixel ids first, then the rows); the additions keep the defined order
            for (int k0 = a0 + lane; k0 < a1; k0 += 128) {
                int pid[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) pid[u] = (k0 + 32 * u < a1) ? pix_idx[k0 + 32 * u] : -1;
                float x[4][SEG_CH];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float* px = src + (size_t)(pid[u] < 0 ? 0 : pid[u]) * C + c0;
#pragma unroll
                    for (int c = 0; c < SEG_CH; ++c) x[u][c] = (pid[u] >= 0 && c < cw) ? px[c] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (pid[u] >= 0) {
#pragma unroll
                        for (int c = 0; c < SEG_CH; ++c) acc[c] += x[u][c];
                    }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int c = 0; c < SEG_CH; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
            }
            float mine = 0.f;
#pragma unroll
            for (int c = 0; c < SEG_CH; ++c)
                if (lane == c) mine = acc[c];
            if (lane < cw) out[((size_t)b * n_cap + v) * C + c0 + lane] = divide ? mine / npix[v] : mine;
        }
    }
}

// The same for a mesh whose nodes all own exactly ONE pixel (every pixel-wise mesh): 0 + x and 31 zero partials give x, so a
// thread per output computes the identical value.
__global__ void segment_single_kernel(const float* __restrict__ img, int B, int P, int C, const int* __restrict__ pix_ptr,
                                      const int* __restrict__ pix_idx, const float* __restrict__ npix, int n_cap,
                                      const int* __restrict__ n_nodes_dev, int divide, float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_nodes = n_nodes_dev ? *n_nodes_dev : n_cap;
    if (t >= (long long)B * n_cap * C) return;
    const int c = (int)(t % C);
    const int v = (int)((t / C) % n_cap);
    const int b = (int)(t / ((long long)C * n_cap));
    if (v >= n_nodes) return;
    const float s = 0.f + img[((size_t)b * P + pix_idx[pix_ptr[v]]) * C + c];
    out[((size_t)b * n_cap + v) * C + c] = divide ? s / npix[v] : s;
}

// img[b, p, c] = data[b, labels[p], c] (* 1/npix when divide) or fill
__global__ void gather_by_label_kernel(const float* __restrict__ data, int B, int P, int C, int n_stride,
                                       const int* __restrict__ labels, const float* __restrict__ npix, int divide,
                                       float fill, float* __restrict__ img) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * P * C) return;
    const int c = (int)(t % C);
    const int p = (int)((t / C) % P);
    const int b = (int)(t / ((long long)C * P));
    const int v = labels[p];
    float val = fill;
    if (v >= 0) {
        val = data[((size_t)b * n_stride + v) * C + c];
        if (divide) val = val / npix[v];
    }
    img[t] = val;
}


// ---- regrid: nodes of mesh S -> pixels -> nodes of mesh D in ONE pass, without the [P, C] image in between ----------------
// (model/seq2seq.py:434-491 do_remesh: unflatten(hidden, old mapping) then flatten(image, new mapping), per recurrent state and
// forecast step; and its autograd: pool backward = gather by the NEW labels with the division, unpool backward = segment sum
// over the OLD pixel lists.)
//   out[t][b, v, c] = (sum over the pixels p of D's node v, in segment_sum_kernel's DEFINED order, of val(p)) (/ npix_d[v])
//   val(p) = lab_s[p] >= 0 ? src[t][b, lab_s[p], c] (/ npix_s[lab_s[p]]) : fill        lab_s == null: src is an image [B, P, C]
// Bit-identical to gather_by_label_kernel followed by segment_sum_kernel: the gather is a copy (with the same division), the
// 32 partial sums and the xor butterfly are the ones defined there.  What changes is the lane mapping: a node with <= 4 (<= 16)
// pixels uses 4 (16) pixel slots x 8 (2) channel quads per warp pass instead of 32 slots x 8 channels of which 28 (16) idle;
// partial sums of the slots a node does not reach are +0 and x + 0 == x, so the skipped butterfly levels change nothing.
// Up to two sources / outputs (t): the hidden and the cell state are regridded by one launch.  Channels move as float4 when
// C % 4 == 0, one by one otherwise.
struct RegridArgs {
    const float* src[2]; float* out[2]; int nt;
    int B, P, C, n_src, n_dst;
    const int* lab_s; const float* npix_s; int src_divide; float fill;
    const int* pix_ptr; const int* pix_idx; const float* npix_d; int dst_divide;
};

template <int V> struct RgVec;
template <> struct RgVec<4> {
    float4 v;
    __device__ __forceinline__ void set(float f) { v = make_float4(f, f, f, f); }
    __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void div(float d) { v.x = v.x / d; v.y = v.y / d; v.z = v.z / d; v.w = v.w / d; }
    __device__ __forceinline__ void add(const RgVec& o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
    __device__ __forceinline__ void add_xor(int off) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off); v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, off); v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
    }
};
template <> struct RgVec<1> {
    float v;
    __device__ __forceinline__ void set(float f) { v = f; }
    __device__ __forceinline__ void load(const float* p) { v = *p; }
    __device__ __forceinline__ void store(float* p) const { *p = v; }
    __device__ __forceinline__ void div(float d) { v = v / d; }
    __device__ __forceinline__ void add(const RgVec& o) { v += o.v; }
    __device__ __forceinline__ void add_xor(int off) { v += __shfl_xor_sync(0xffffffffu, v, off); }
};

// One node, channel vectors [cv0, cv1) (a vector = V channels): S pixel slots x 32 / S channel vectors per warp pass.
template <int S, int V>
__device__ __forceinline__ void regrid_node(const RegridArgs& a, const float* __restrict__ src, float* __restrict__ out, int b,
                                            int v, int a0, int a1, int lane, int cv0, int cv1) {
    constexpr int Q = 32 / S;                    // channel vectors per pass
    constexpr int U = 8;                         // rounds of loads in flight; the additions keep the defined order
    const int slot = lane & (S - 1), q = lane / S;
    const size_t sbase = (size_t)b * (a.lab_s ? a.n_src : a.P) * a.C;
    for (int cvb = cv0; cvb < cv1; cvb += Q) {
        const int cq = (cvb + q) * V;
        const bool on = cvb + q < cv1;
        RgVec<V> acc;
        acc.set(0.f);
        for (int k0 = a0 + slot; k0 < a1; k0 += U * S) {
            int row[U];
            float dv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) row[u] = (k0 + S * u < a1) ? a.pix_idx[k0 + S * u] : -2;
            if (a.lab_s) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (row[u] >= 0) row[u] = a.lab_s[row[u]];
                if (a.src_divide) {
#pragma unroll
                    for (int u = 0; u < U; ++u) dv[u] = (row[u] >= 0) ? a.npix_s[row[u]] : 1.f;
                }
            }
            RgVec<V> x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                x[u].set(a.fill);
                if (row[u] >= 0 && on) x[u].load(src + sbase + (size_t)row[u] * a.C + cq);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (row[u] != -2) {
                    if (a.lab_s && a.src_divide && row[u] >= 0) x[u].div(dv[u]);
                    acc.add(x[u]);
                }
        }
#pragma unroll
        for (int off = S / 2; off > 0; off >>= 1) acc.add_xor(off);
        if (slot == 0 && on) {
            if (a.dst_divide) acc.div(a.npix_d[v]);
            acc.store(out + ((size_t)b * a.n_dst + v) * a.C + cq);
        }
    }
}

// Work item of a warp = 32 consecutive nodes of one frame b, both tensors (they share the index chain pix_ptr -> pix_idx ->
// lab_s).  Nodes with <= 4 pixels -- all of a pixel-level mesh, most of any quadtree mesh -- are summed LANE-SERIALLY: lane l owns
// node v0 + l, loads its <= 4 rows itself and adds them as the butterfly would ((p0 + p2) + (p1 + p3), absent partials +0), so 32
// index chains are in flight per warp instead of one.  Larger nodes are then taken one at a time by the whole warp
// (regrid_node: 16 or 32 pixel slots).
template <int V>
__global__ void __launch_bounds__(256) regrid_kernel(const RegridArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int groups = (a.n_dst + 31) / 32, ncv = a.C / V;
    for (long long w = warp; w < (long long)a.B * groups; w += nwarps) {
        const int b = (int)(w / groups), v = (int)(w - (long long)b * groups) * 32 + lane;
        const bool valid = v < a.n_dst;
        const int a0 = valid ? a.pix_ptr[v] : 0, a1 = valid ? a.pix_ptr[v + 1] : 0;
        const int np = a1 - a0;
        if (valid && np <= 4) {
            int r[4];
            float dv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = (k < np) ? a.pix_idx[a0 + k] : -2;
            if (a.lab_s) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (r[k] >= 0) r[k] = a.lab_s[r[k]];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) dv[k] = (a.lab_s && a.src_divide && r[k] >= 0) ? a.npix_s[r[k]] : 1.f;
            const float d = a.dst_divide ? a.npix_d[v] : 1.f;
            const size_t sbase = (size_t)b * (a.lab_s ? a.n_src : a.P) * a.C, obase = ((size_t)b * a.n_dst + v) * a.C;
            for (int t = 0; t < a.nt; ++t) {
                const float* src = (t ? a.src[1] : a.src[0]) + sbase;
                float* out = (t ? a.out[1] : a.out[0]) + obase;
                for (int cv = 0; cv < ncv; ++cv) {
                    RgVec<V> pk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        RgVec<V> x;
                        x.set(a.fill);
                        if (r[k] >= 0) x.load(src + (size_t)r[k] * a.C + cv * V);
                        if (a.lab_s && a.src_divide && r[k] >= 0) x.div(dv[k]);
                        pk[k].set(0.f);
                        if (r[k] != -2) pk[k].add(x);
                    }
                    pk[0].add(pk[2]);
                    pk[1].add(pk[3]);
                    pk[0].add(pk[1]);
                    if (a.dst_divide) pk[0].div(d);
                    pk[0].store(out + cv * V);
                }
            }
        }
        unsigned rest = __ballot_sync(0xffffffffu, valid && np > 4);
        while (rest) {
            const int n = __ffs(rest) - 1;
            rest &= rest - 1;
            const int b0 = __shfl_sync(0xffffffffu, a0, n), b1 = __shfl_sync(0xffffffffu, a1, n);
            const int vn = v - lane + n;
            for (int t = 0; t < a.nt; ++t) {
                if (b1 - b0 <= 16) regrid_node<16, V>(a, t ? a.src[1] : a.src[0], t ? a.out[1] : a.out[0], b, vn, b0, b1, lane, 0, ncv);
                else regrid_node<32, V>(a, t ? a.src[1] : a.src[0], t ? a.out[1] : a.out[0], b, vn, b0, b1, lane, 0, ncv);
            }
        }
    }
}

}  // namespace qmp
using namespace qmp;

// Build pix_ptr/pix_idx for a quadtree mesh from the leaf rectangles of qmp_quadtree_labels.
// cap = capacity of the node arrays (>= true N, e.g. P); scratch: tmp int32 [cap], blocksums [cap/1024+2].
QMP_API int qmp_mesh_pixels_from_rects(const int* labels, int n, int m, const int* node_rect, const float* npix,
                                       const int* n_nodes, int cap, int* pix_ptr, int* pix_idx, int* tmp, int* blocksums,
                                       void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int P = n * m;
    QMP_REQUIRE(cap >= 1, "qmp_mesh_pixels_from_rects: cap");
    npix_to_int_kernel<<<cdiv(cap, 256), 256, 0, st>>>(npix, n_nodes, cap, tmp);
    // pix_ptr has cap+1 entries: exclusive scan of cap counts, last entry = total
    int rc = exclusive_scan_i32(tmp, pix_ptr, cap, pix_ptr + cap, blocksums, st);
    if (rc) return rc;
    pix_csr_from_rects_kernel<<<cdiv(P, 256), 256, 0, st>>>(labels, P, m, (const int4*)node_rect, pix_ptr, pix_idx);
    QMP_LAUNCH_CHECK("qmp_mesh_pixels_from_rects");
    return 0;
}

// Pixel-wise mesh from a mask (NULL = keep everything).  Outputs sized for P nodes; n_nodes on device.
// scratch: keep int32 [P], rank int32 [P], blocksums [P/1024+2].
QMP_API int qmp_mesh_pixelwise(const uint8_t* mask, int P, int* labels, int* pix_ptr, int* pix_idx, float* npix,
                               int* n_nodes, int* keep, int* rank, int* blocksums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    keep_flags_kernel<<<cdiv(P, 256), 256, 0, st>>>(mask, P, keep);
    int rc = exclusive_scan_i32(keep, rank, P, n_nodes, blocksums, st);
    if (rc) return rc;
    pixelwise_fill_kernel<<<cdiv(P + 1, 256), 256, 0, st>>>(keep, rank, P, labels, pix_idx, pix_ptr, npix);
    QMP_LAUNCH_CHECK("qmp_mesh_pixelwise");
    return 0;
}

QMP_API int qmp_regrid(const float* src0, const float* src1, int B, int P, int C, int n_src, const int* lab_s,
                       const float* npix_s, int src_divide, float fill, const int* pix_ptr, const int* pix_idx,
                       const float* npix_d, int n_dst, int dst_divide, float* out0, float* out1, void* stream);

// pool forward (divide=1) / unpool backward (divide=0).  img [B,P,C] -> out [B,n_cap,C].  single = 1: the caller guarantees
// that every node owns exactly one pixel (a pixel-wise mesh) -- same values, a thread per output instead of a warp per node.
QMP_API int qmp_segment_sum(const float* img, int B, int P, int C, const int* pix_ptr, const int* pix_idx,
                            const float* npix, int n_cap, const int* n_nodes_dev, int divide, int single, float* out, void* stream) {
    const long long tot = (long long)B * n_cap * C;
    if (tot == 0) return 0;
    if (single) {
        segment_single_kernel<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(img, B, P, C, pix_ptr, pix_idx, npix, n_cap,
                                                                                n_nodes_dev, divide, out);
    } else if (!n_nodes_dev) {
        // the regrid kernel with an image as its source: same sums, lanes split over channel quads when a node is small
        return qmp_regrid(img, nullptr, B, P, C, 0, nullptr, nullptr, 0, 0.f, pix_ptr, pix_idx, npix, n_cap, divide, out, nullptr, stream);
    } else {
        // one warp per (frame, node), grid-stride (the node count of a capacity launch lives on the device)
        const long long items = (long long)B * n_cap;
        const int grid = (int)(items < 148 * 16 * 8 ? (items + 7) / 8 : 148 * 16);
        segment_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, B, P, C, pix_ptr, pix_idx, npix, n_cap, n_nodes_dev, divide, out);
    }
    QMP_LAUNCH_CHECK("qmp_segment_sum");
    return 0;
}

// unpool forward (divide=0, fill = 0 or NaN) / pool backward (divide=1, fill=0).  data [B,n_stride,C] -> img [B,P,C].
QMP_API int qmp_gather_by_label(const float* data, int B, int P, int C, int n_stride, const int* labels,
                                const float* npix, int divide, float fill, float* img, void* stream) {
    const long long tot = (long long)B * P * C;
    if (tot == 0) return 0;
    gather_by_label_kernel<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(data, B, P, C, n_stride, labels, npix,
                                                                             divide, fill, img);
    QMP_LAUNCH_CHECK("qmp_gather_by_label");
    return 0;
}

// Regrid node data from mesh S (labels lab_s, optional division by npix_s: pool backward) onto mesh D (pixel lists, optional
// division by npix_d: pool forward) without materialising the image; src1 / out1 may be null (one tensor).  lab_s == null:
// src is an image [B, P, C] (plain pooling with the small-node lane mapping).  Same values as qmp_gather_by_label followed by
// qmp_segment_sum, bit for bit.
QMP_API int qmp_regrid(const float* src0, const float* src1, int B, int P, int C, int n_src, const int* lab_s,
                       const float* npix_s, int src_divide, float fill, const int* pix_ptr, const int* pix_idx,
                       const float* npix_d, int n_dst, int dst_divide, float* out0, float* out1, void* stream) {
    RegridArgs a;
    a.src[0] = src0; a.src[1] = src1; a.out[0] = out0; a.out[1] = out1; a.nt = (src1 && out1) ? 2 : 1;
    a.B = B; a.P = P; a.C = C; a.n_src = n_src; a.n_dst = n_dst;
    a.lab_s = lab_s; a.npix_s = npix_s; a.src_divide = src_divide; a.fill = fill;
    a.pix_ptr = pix_ptr; a.pix_idx = pix_idx; a.npix_d = npix_d; a.dst_divide = dst_divide;
    if ((long long)B * n_dst == 0) return 0;
    const long long items = (long long)B * ((n_dst + 31) / 32);          // a warp per 32 nodes
    const int grid = (int)(items < 148 * 16 * 8 ? (items + 7) / 8 : 148 * 16);
    if (C % 4 == 0) regrid_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else regrid_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    QMP_LAUNCH_CHECK("qmp_regrid");
    return 0;
}
