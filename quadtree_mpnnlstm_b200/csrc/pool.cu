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

// pool forward (divide=1) / unpool backward (divide=0).  img [B,P,C] -> out [B,n_cap,C].  single = 1: the caller guarantees
// that every node owns exactly one pixel (a pixel-wise mesh) -- same values, a thread per output instead of a warp per node.
QMP_API int qmp_segment_sum(const float* img, int B, int P, int C, const int* pix_ptr, const int* pix_idx,
                            const float* npix, int n_cap, const int* n_nodes_dev, int divide, int single, float* out, void* stream) {
    const long long tot = (long long)B * n_cap * C;
    if (tot == 0) return 0;
    if (single) {
        segment_single_kernel<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(img, B, P, C, pix_ptr, pix_idx, npix, n_cap,
                                                                                n_nodes_dev, divide, out);
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
