// K4: edge emission in the reference's order, edge attributes, positional encoding.
//
// get_adj (reference model/graph_functions.py:291-345) walks pixels in raster order, puts the up /
// down / left / right labels into a Python set, removes -1, iterates the set and emits (node, nb)
// the first time nb is seen for node.  Two things fix the output order:
//   (1) within a pixel, CPython's iteration order of a <=4-element set of small ints: an 8-slot
//       open-addressed table, slot = hash & 7 (hash(-1) = -2), collisions resolved by
//       perturb >>= 5; i = (5 i + 1 + perturb) & 7, no linear probing at this size, no resize
//       (oracle/graph_ref.py:cpython_small_set_order, checked there against the real set);
//   (2) across pixels, the first pixel (raster order) of `node` that touches `nb`.
// (2) is resolved with a hash table keyed by (node, nb) holding the minimum of pixel*8 + slot
// (atomicMin is order independent, so the result is deterministic); a pixel emits exactly the
// candidates whose own code equals the table's minimum, then a scan compacts them in code order.
// get_adj_pixelwise (:471-493) emits per pixel [row+1, row-1, col+1, col-1] and drops pairs that
// touch -1; it goes through the same count / scan / emit path.
#include "common.cuh"

namespace qmp {

constexpr int SLOT_EMPTY = INT_MIN;
constexpr unsigned long long KEY_EMPTY = ~0ull;

struct PixelSet {
    int slot[8];
    __device__ void clear() {
#pragma unroll
        for (int i = 0; i < 8; ++i) slot[i] = SLOT_EMPTY;
    }
    __device__ void add(int v) {
        const long long h = (v == -1) ? -2ll : (long long)v;
        unsigned long long perturb = (unsigned long long)h;
        int i = (int)(h & 7);
        while (true) {
            if (slot[i] == SLOT_EMPTY) {
                slot[i] = v;
                return;
            }
            if (slot[i] == v) return;
            perturb >>= 5;
            i = (int)((5ull * (unsigned)i + 1ull + perturb) & 7ull);
        }
    }
};

__device__ __forceinline__ void quadtree_candidates(const int* __restrict__ labels, int rows, int cols, int i, int j,
                                                    PixelSet& s) {
    s.clear();
    if (i != 0) s.add(labels[(size_t)(i - 1) * cols + j]);
    if (i != rows - 1) s.add(labels[(size_t)(i + 1) * cols + j]);
    if (j != 0) s.add(labels[(size_t)i * cols + j - 1]);
    if (j != cols - 1) s.add(labels[(size_t)i * cols + j + 1]);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

__global__ void adj_insert_kernel(const int* __restrict__ labels, int rows, int cols, unsigned long long* __restrict__ keys,
                                  int* __restrict__ vals, unsigned cap_mask) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows * cols) return;
    const int v = labels[p];
    if (v < 0) return;
    PixelSet s;
    quadtree_candidates(labels, rows, cols, p / cols, p % cols, s);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int u = s.slot[k];
        if (u == SLOT_EMPTY || u == -1) continue;
        const unsigned long long key = ((unsigned long long)(unsigned)v << 32) | (unsigned)u;
        unsigned h = (unsigned)mix64(key) & cap_mask;
        while (true) {
            const unsigned long long old = atomicCAS(&keys[h], KEY_EMPTY, key);
            if (old == KEY_EMPTY || old == key) {
                atomicMin(&vals[h], p * 8 + k);
                break;
            }
            h = (h + 1) & cap_mask;
        }
    }
}

// per pixel: bit k of emit[p] set when slot k of this pixel is the first occurrence of its pair
__global__ void adj_flag_kernel(const int* __restrict__ labels, int rows, int cols,
                                const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                unsigned cap_mask, uint8_t* __restrict__ emit, int* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows * cols) return;
    const int v = labels[p];
    unsigned bits = 0;
    if (v >= 0) {
        PixelSet s;
        quadtree_candidates(labels, rows, cols, p / cols, p % cols, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int u = s.slot[k];
            if (u == SLOT_EMPTY || u == -1) continue;
            const unsigned long long key = ((unsigned long long)(unsigned)v << 32) | (unsigned)u;
            unsigned h = (unsigned)mix64(key) & cap_mask;
            while (keys[h] != key) h = (h + 1) & cap_mask;
            if (vals[h] == p * 8 + k) bits |= 1u << k;
        }
    }
    emit[p] = (uint8_t)bits;
    count[p] = __popc(bits);
}

__global__ void adj_emit_kernel(const int* __restrict__ labels, int rows, int cols, const uint8_t* __restrict__ emit,
                                const int* __restrict__ offset, long long* __restrict__ src64, long long* __restrict__ dst64,
                                int* __restrict__ src32, int* __restrict__ dst32) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows * cols) return;
    const unsigned bits = emit[p];
    if (!bits) return;
    const int v = labels[p];
    PixelSet s;
    quadtree_candidates(labels, rows, cols, p / cols, p % cols, s);
    int e = offset[p];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (bits & (1u << k)) {
            src64[e] = v;
            dst64[e] = s.slot[k];
            src32[e] = v;
            dst32[e] = s.slot[k];
            ++e;
        }
}

// pixel-wise: candidates [row+1, row-1, col+1, col-1]
__device__ __forceinline__ void pixelwise_candidates(const int* __restrict__ labels, int rows, int cols, int i, int j,
                                                     int nb[4]) {
    nb[0] = (i != rows - 1) ? labels[(size_t)(i + 1) * cols + j] : -1;
    nb[1] = (i != 0) ? labels[(size_t)(i - 1) * cols + j] : -1;
    nb[2] = (j != cols - 1) ? labels[(size_t)i * cols + j + 1] : -1;
    nb[3] = (j != 0) ? labels[(size_t)i * cols + j - 1] : -1;
}

__global__ void adjpx_count_kernel(const int* __restrict__ labels, int rows, int cols, int* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows * cols) return;
    int n = 0;
    if (labels[p] >= 0) {
        int nb[4];
        pixelwise_candidates(labels, rows, cols, p / cols, p % cols, nb);
#pragma unroll
        for (int k = 0; k < 4; ++k) n += (nb[k] >= 0);
    }
    count[p] = n;
}

__global__ void adjpx_emit_kernel(const int* __restrict__ labels, int rows, int cols, const int* __restrict__ offset,
                                  long long* __restrict__ src64, long long* __restrict__ dst64, int* __restrict__ src32,
                                  int* __restrict__ dst32) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows * cols) return;
    const int v = labels[p];
    if (v < 0) return;
    int nb[4];
    pixelwise_candidates(labels, rows, cols, p / cols, p % cols, nb);
    int e = offset[p];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (nb[k] >= 0) {
            src64[e] = v;
            dst64[e] = nb[k];
            src32[e] = v;
            dst32[e] = nb[k];
            ++e;
        }
}

// angle = atan2(dx, dy) mod 2pi / 2pi, dist = sqrt(dy^2 + dx^2), float32 like torch
// (graph_functions.py:358-370); xx = ii * W * res, yy = jj * H * res (:657 / :519)
__global__ void edge_attr_kernel(const int* __restrict__ src, const int* __restrict__ dst, int e_cap,
                                 const int* __restrict__ n_edges_dev, const float* __restrict__ pos_ii,
                                 const float* __restrict__ pos_jj, int pos_stride, float w_img, float h_img, float res,
                                 int two_cols, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_edges = n_edges_dev ? *n_edges_dev : e_cap;
    if (e >= e_cap || e >= n_edges) return;
    const int s = src[e], d = dst[e];
    const float xs = __fmul_rn(__fmul_rn(pos_ii[(size_t)s * pos_stride], w_img), res);
    const float xd = __fmul_rn(__fmul_rn(pos_ii[(size_t)d * pos_stride], w_img), res);
    const float ys = __fmul_rn(__fmul_rn(pos_jj[(size_t)s * pos_stride], h_img), res);
    const float yd = __fmul_rn(__fmul_rn(pos_jj[(size_t)d * pos_stride], h_img), res);
    const float dx = __fsub_rn(xs, xd), dy = __fsub_rn(ys, yd);
    const float dist = sqrtf(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
    if (two_cols) {
        const float two_pi = 6.283185307179586f;
        float a = atan2f(dx, dy);
        float r = fmodf(a, two_pi);
        if (r != 0.f && r < 0.f) r += two_pi;  // Python-style modulus (torch.remainder)
        out[(size_t)e * 2] = r / two_pi;
        out[(size_t)e * 2 + 1] = dist;
    } else {
        out[e] = dist;
    }
}

// out[b, r, c, :] = cat(x[b, r, c, :], c / W, r / H)  (model/utils.py:30-52; planes built in float64)
__global__ void add_pos_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Co = C + 2;
    if (t >= (long long)B * H * W * Co) return;
    const int c = (int)(t % Co);
    const long long px = t / Co;
    const int col = (int)(px % W), row = (int)((px / W) % H);
    float v;
    if (c < C) v = x[px * C + c];
    else if (c == C) v = (float)((double)col / (double)W);
    else v = (float)((double)row / (double)H);
    out[t] = v;
}

__global__ void fill_u64_kernel(unsigned long long* p, long long n, unsigned long long v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void fill_i32_kernel(int* p, long long n, int v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace qmp
using namespace qmp;

// Quadtree adjacency.  labels int32 [rows, cols].  Outputs (capacity e_cap each): src/dst as int64 and
// int32, n_edges (device).  Scratch: keys u64 [table_cap], vals int32 [table_cap] with table_cap a power of
// two >= 8 * rows * cols; emit uint8 [P]; count/offset int32 [P]; blocksums [P/1024+2].
QMP_API int qmp_adjacency_quadtree(const int* labels, int rows, int cols, long long* src64, long long* dst64, int* src32,
                                   int* dst32, int* n_edges, unsigned long long* keys, int* vals, long long table_cap,
                                   uint8_t* emit, int* count, int* offset, int* blocksums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int P = rows * cols;
    QMP_REQUIRE((table_cap & (table_cap - 1)) == 0 && table_cap >= 8ll * P && table_cap <= (1ll << 31),
                "qmp_adjacency_quadtree: table_cap must be a power of two >= 8*P");
    QMP_REQUIRE((long long)P * 8 < INT_MAX, "qmp_adjacency_quadtree: image too large");
    fill_u64_kernel<<<cdiv(table_cap, 256), 256, 0, st>>>(keys, table_cap, KEY_EMPTY);
    fill_i32_kernel<<<cdiv(table_cap, 256), 256, 0, st>>>(vals, table_cap, INT_MAX);
    const unsigned cap_mask = (unsigned)(table_cap - 1);
    adj_insert_kernel<<<cdiv(P, 128), 128, 0, st>>>(labels, rows, cols, keys, vals, cap_mask);
    adj_flag_kernel<<<cdiv(P, 128), 128, 0, st>>>(labels, rows, cols, keys, vals, cap_mask, emit, count);
    int rc = exclusive_scan_i32(count, offset, P, n_edges, blocksums, st);
    if (rc) return rc;
    adj_emit_kernel<<<cdiv(P, 128), 128, 0, st>>>(labels, rows, cols, emit, offset, src64, dst64, src32, dst32);
    QMP_LAUNCH_CHECK("qmp_adjacency_quadtree");
    return 0;
}

QMP_API int qmp_adjacency_pixelwise(const int* labels, int rows, int cols, long long* src64, long long* dst64, int* src32,
                                    int* dst32, int* n_edges, int* count, int* offset, int* blocksums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int P = rows * cols;
    adjpx_count_kernel<<<cdiv(P, 256), 256, 0, st>>>(labels, rows, cols, count);
    int rc = exclusive_scan_i32(count, offset, P, n_edges, blocksums, st);
    if (rc) return rc;
    adjpx_emit_kernel<<<cdiv(P, 256), 256, 0, st>>>(labels, rows, cols, offset, src64, dst64, src32, dst32);
    QMP_LAUNCH_CHECK("qmp_adjacency_pixelwise");
    return 0;
}

// pos_ii / pos_jj: pointers to the ii and jj feature of node 0 (frame 0), pos_stride floats between nodes.
// two_cols = 1 -> out [E, 2] = (angle, dist); 0 -> out [E] = dist.
QMP_API int qmp_edge_attrs(const int* src, const int* dst, int e_cap, const int* n_edges_dev, const float* pos_ii,
                           const float* pos_jj, int pos_stride, int img_w, int img_h, float resolution, int two_cols,
                           float* out, void* stream) {
    if (e_cap == 0) return 0;
    edge_attr_kernel<<<cdiv(e_cap, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, e_cap, n_edges_dev, pos_ii, pos_jj,
                                                                         pos_stride, (float)img_w, (float)img_h,
                                                                         resolution, two_cols, out);
    QMP_LAUNCH_CHECK("qmp_edge_attrs");
    return 0;
}

QMP_API int qmp_add_positional_encoding(const float* x, int B, int H, int W, int C, float* out, void* stream) {
    const long long tot = (long long)B * H * W * (C + 2);
    if (tot == 0) return 0;
    add_pos_kernel<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, out);
    QMP_LAUNCH_CHECK("qmp_add_positional_encoding");
    return 0;
}
