// K1-K4 in ONE cooperative launch: quadtree split pyramid -> leaf labels -> per-node pixel lists -> pooled node features
// -> adjacency in the reference's edge order -> edge attributes -> in / out CSR for the message-passing kernels.
// (reference model/graph_functions.py:145-356 quadtree_decompose / get_adj / dist_angle, :555-587 get_mapping, :391-419
// flatten, :590-681 image_to_graph; north_star (a): TMA-staged tiles, compacted node and CSR edge output.)
//
// The per-phase arithmetic is that of quadtree.cu / pool.cu / edges.cu / csr.cu (those entry points stay for the pixel-wise
// mesh, the generic-label paths and max_size > 64); what changes is the execution: one persistent grid of 2 CTAs per SM walks
// the phases with grid-wide barriers instead of ~50 dependent launches and three host-visible scans, node and edge counts
// stay on the device between phases (every later phase reads them after the barrier), and every output is written compacted:
// data [T, N, C+1], edge_index [2, E], attrs [E, 2], CSR rows -- prefixes of capacity buffers the host slices after ONE
// read-back of (N, E, NaN count).  The split pyramid of a 64 x 64 tile is staged into shared memory with bulk asynchronous
// copies (cp.async.bulk, one per tile row, completion on an mbarrier) from the level-0 planes a first grid-wide pass wrote.
//
// Summation order of the pooling: pool.cu (lane tree), bit-identical.  Edge order: edges.cu (CPython small-set order per pixel,
// first raster pixel per (node, neighbour) pair through a hash table of minima), bit-identical.
#include "common.cuh"
#include "tc.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace qmp {

constexpr int GB_THREADS = 256;
constexpr int GB_TS = 64;                                  // tile edge of the pyramid phase
constexpr int GB_SLOT_EMPTY = INT_MIN;
constexpr unsigned long long GB_KEY_EMPTY = ~0ull;

struct GbArgs {
    // inputs
    const float* img; int T, n, m, C;                      // frames [T, n, m, C]; the last two channels are the node positions
    const float* crit;                                     // [n_pad, m_pad] criterion image (padded, transformed)
    const uint8_t* mask; const uint8_t* hir;               // [n, m] or null
    int n_pad, m_pad, row_cap, L, cond; double thresh;
    float resolution; int two_cols; float size_div;        // cell size column = npix / size_div
    long long e_cap;
    // outputs (capacity buffers, compacted prefixes)
    int* labels; float* npix; int* pix_ptr; int* pix_idx; float* data;
    long long* ei64; int* src32; int* dst32; float* edge_attrs;
    int* in_ptr; int* in_src; int* in_eid; int* out_ptr; int* out_dst; int* out_kin; float* edge_attr_in;
    int* counts;                                           // [4]: n_nodes, n_edges, n_nan, 0  (zero on entry)
    // scratch
    float* e0; int* ka0;                                   // level-0 planes [n_pad, m_pad]: window extreme, cnt | any << 1
    int* cnt; uint8_t* split;                              // pyramid (levels 1..L; level l at level_off(l))
    int* base_off; int4* rect;
    unsigned long long* keys; int* vals; unsigned cap_mask; long long table_cap;
    uint8_t* emit; int* count;
    int* part_a; int* part_b;                              // per-CTA partial sums of the grid-wide scans [gridDim]
    int* cnt_in; int* cnt_out; int* cur_in; int* cur_out; int* eid_out; int* kin_of_edge; int* tmp_in; int* tmp_out;
};

__device__ __forceinline__ long long gb_level_off(int n_pad, int m_pad, int lvl) {
    long long off = 0;
    for (int k = 0; k < lvl; ++k) off += (long long)(n_pad >> k) * (m_pad >> k);
    return off;
}
__device__ __forceinline__ bool gb_split(float e, int cond, double thresh) {
    const double d = (double)e;            // the reference compares the float32 extreme with a float64 threshold
    return (cond == 0) ? (d > thresh) : (cond == 1) ? (d < thresh) : (cond == 2) ? (d > thresh) : (d < thresh);
}
__device__ __forceinline__ float gb_ext(float a, float b, bool is_max) { return is_max ? fmaxf(a, b) : fminf(a, b); }

// ---- block primitives (256 threads)
__device__ __forceinline__ int gb_block_sum(int v, int* s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) s_w[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < GB_THREADS / 32; ++w) t += s_w[w];
    return t;
}
// exclusive prefix of v over the block (thread order); total = block sum
__device__ __forceinline__ int gb_block_excl(int v, int* s_w, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < GB_THREADS / 32; ++w) {
        const int x = s_w[w];
        if (w < warp) woff += x;
        tot += x;
    }
    total = tot;
    return woff + incl - v;
}
// contiguous chunk of CTA b out of G over n items (the same split in both passes of a scan)
__device__ __forceinline__ void gb_chunk(int n, int& lo, int& hi) {
    const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
    lo = min(n, (int)blockIdx.x * per);
    hi = min(n, lo + per);
}
// sum of part[0 .. blockIdx.x - 1]
__device__ __forceinline__ int gb_base_of(const int* part, int* s_w) {
    int v = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += GB_THREADS) v += part[b];
    return gb_block_sum(v, s_w);
}

// ---- CPython iteration order of a <= 4-element set of small ints (edges.cu)
struct GbSet {
    int slot[8];
    __device__ void clear() {
#pragma unroll
        for (int i = 0; i < 8; ++i) slot[i] = GB_SLOT_EMPTY;
    }
    __device__ void add(int v) {
        const long long h = (v == -1) ? -2ll : (long long)v;
        unsigned long long perturb = (unsigned long long)h;
        int i = (int)(h & 7);
        while (true) {
            if (slot[i] == GB_SLOT_EMPTY) {
                slot[i] = v;
                return;
            }
            if (slot[i] == v) return;
            perturb >>= 5;
            i = (int)((5ull * (unsigned)i + 1ull + perturb) & 7ull);
        }
    }
};
__device__ __forceinline__ void gb_candidates(const int* __restrict__ labels, int rows, int cols, int i, int j, GbSet& s) {
    s.clear();
    if (i != 0) s.add(labels[(size_t)(i - 1) * cols + j]);
    if (i != rows - 1) s.add(labels[(size_t)(i + 1) * cols + j]);
    if (j != 0) s.add(labels[(size_t)i * cols + j - 1]);
    if (j != cols - 1) s.add(labels[(size_t)i * cols + j + 1]);
}
__device__ __forceinline__ unsigned long long gb_mix64(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// phase timeline: CTA 0 stamps the global timer (ns) after every grid-wide barrier into the arena's first 256 bytes, behind the
// four counts (bytes 64 ..: read by scripts/build_phases.py; 16 stores per launch)
__device__ __forceinline__ void gb_mark(const GbArgs& a, int k) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long*>(a.counts)[8 + k] = t;
    }
}

__global__ void __launch_bounds__(GB_THREADS, 2) quadtree_graph_kernel(const GbArgs a) {
    cg::grid_group grid = cg::this_grid();
    gb_mark(a, 0);
    __shared__ __align__(128) float s_e0[GB_TS * GB_TS];   // staged level-0 tile (bulk copies)
    __shared__ __align__(128) int s_k0[GB_TS * GB_TS];
    __shared__ float s_ext[1024 + 256];                    // levels >= 1: ping (1024) / pong (256)
    __shared__ int s_cnt[1024 + 256];
    __shared__ uint8_t s_any[1024 + 256];
    __shared__ int s_w[GB_THREADS / 32];
    __shared__ uint64_t s_bar;
    const int tid = threadIdx.x;
    const long long gtid = (long long)blockIdx.x * GB_THREADS + tid, gthreads = (long long)gridDim.x * GB_THREADS;
    const int P = a.n * a.m;
    const bool is_max = a.cond < 2;
    const float neutral = is_max ? -FLT_MAX : FLT_MAX;

    // ---------------------------------------------------------------- phase 0: level-0 planes, table fill, NaN count
    for (long long idx = gtid; idx < (long long)a.n_pad * a.m_pad; idx += gthreads) {
        const int r = (int)(idx / a.m_pad), c = (int)(idx % a.m_pad);
        float e = neutral;
        bool any = false;
        const int r1 = min(r + 2, a.row_cap), c1 = min(c + 2, a.m_pad);
        for (int rr = r; rr < r1; ++rr)
            for (int cc = c; cc < c1; ++cc) {
                e = gb_ext(e, a.crit[(size_t)rr * a.m_pad + cc], is_max);
                if (rr < a.n && cc < a.m) {
                    if (a.mask && a.mask[(size_t)rr * a.m + cc]) any = true;
                    if (a.hir && a.hir[(size_t)rr * a.m + cc]) any = true;
                }
            }
        const int k = (r < a.n && c < a.m && !(a.mask && a.mask[(size_t)r * a.m + c])) ? 1 : 0;
        a.e0[idx] = e;
        a.ka0[idx] = k | ((int)any << 1);
    }
    for (long long i = gtid; i < a.table_cap; i += gthreads) {
        a.keys[i] = GB_KEY_EMPTY;
        a.vals[i] = INT_MAX;
    }
    for (long long i = gtid; i < (long long)P + 2; i += gthreads) {
        a.cnt_in[i] = 0;
        a.cnt_out[i] = 0;
    }
    {
        int nan = 0;
        const long long tot = (long long)a.T * P * a.C;
        for (long long i = gtid; i < tot; i += gthreads) nan += (a.img[i] != a.img[i]);
        nan = gb_block_sum(nan, s_w);
        if (tid == 0 && nan) atomicAdd(&a.counts[2], nan);
    }
    if (tid == 0) {
        tc::mbar_init(&s_bar, 1);
        tc::fence_mbar_init();
    }
    grid.sync();
    gb_mark(a, 1);

    // ---------------------------------------------------------------- phase 1: split pyramid per 64 x 64 tile
    {
        const int tcols = (a.m_pad + GB_TS - 1) / GB_TS, trows = (a.n_pad + GB_TS - 1) / GB_TS;
        uint32_t bar_par = 0;
        for (int tile = blockIdx.x; tile < tcols * trows; tile += gridDim.x) {
            const int r0 = (tile / tcols) * GB_TS, c0 = (tile % tcols) * GB_TS;
            const int hv = min(GB_TS, a.n_pad - r0), wv = min(GB_TS, a.m_pad - c0);    // valid part (multiples of 4 or more)
            if (hv < GB_TS || wv < GB_TS) {                 // partial tile: neutral everywhere the copies do not land
                for (int t = tid; t < GB_TS * GB_TS; t += GB_THREADS) {
                    s_e0[t] = neutral;
                    s_k0[t] = 0;
                }
                tc::fence_async_smem();
            }
            __syncthreads();
            if ((a.m_pad & 3) == 0) {                       // rows are 16-byte multiples at 16-byte addresses: bulk copies
                if (tid == 0) {
                    tc::mbar_expect_tx(&s_bar, (uint32_t)(hv * wv * 8));
                    for (int r = 0; r < hv; ++r) {
                        const size_t g = (size_t)(r0 + r) * a.m_pad + c0;
                        tc::bulk_g2s(s_e0 + r * GB_TS, a.e0 + g, (uint32_t)(wv * 4), &s_bar);
                        tc::bulk_g2s(s_k0 + r * GB_TS, a.ka0 + g, (uint32_t)(wv * 4), &s_bar);
                    }
                }
                tc::mbar_wait(&s_bar, bar_par);
                bar_par ^= 1;
            } else {                                        // max_size 1 or 2 on an odd-width image: plain loads
                for (int t = tid; t < hv * wv; t += GB_THREADS) {
                    const int r = t / wv, c = t % wv;
                    const size_t g = (size_t)(r0 + r) * a.m_pad + c0 + c;
                    s_e0[r * GB_TS + c] = a.e0[g];
                    s_k0[r * GB_TS + c] = a.ka0[g];
                }
                __syncthreads();
            }
            // level l (1..L) of this tile from level l - 1; level 1 reads the staged planes
            int src = 0, dst = 0, w_prev = GB_TS;
            for (int lvl = 1; lvl <= a.L; ++lvl) {
                const int w = GB_TS >> lvl;
                const long long off = gb_level_off(a.n_pad, a.m_pad, lvl);
                const int gcols = a.m_pad >> lvl;
                dst = (lvl & 1) ? 0 : 1024;
                for (int t = tid; t < w * w; t += GB_THREADS) {
                    const int cr = t / w, cc = t % w;
                    const int a00 = (2 * cr) * w_prev + 2 * cc;
                    float e;
                    bool any;
                    int ksum;
                    if (lvl == 1) {
                        const int k0 = s_k0[a00], k1 = s_k0[a00 + 1], k2 = s_k0[a00 + w_prev], k3 = s_k0[a00 + w_prev + 1];
                        e = gb_ext(gb_ext(s_e0[a00], s_e0[a00 + 1], is_max), gb_ext(s_e0[a00 + w_prev], s_e0[a00 + w_prev + 1], is_max), is_max);
                        any = ((k0 | k1 | k2 | k3) & 2) != 0;
                        ksum = (k0 & 1) + (k1 & 1) + (k2 & 1) + (k3 & 1);
                    } else {
                        const int b00 = src + a00;
                        e = gb_ext(gb_ext(s_ext[b00], s_ext[b00 + 1], is_max), gb_ext(s_ext[b00 + w_prev], s_ext[b00 + w_prev + 1], is_max), is_max);
                        any = s_any[b00] | s_any[b00 + 1] | s_any[b00 + w_prev] | s_any[b00 + w_prev + 1];
                        ksum = s_cnt[b00] + s_cnt[b00 + 1] + s_cnt[b00 + w_prev] + s_cnt[b00 + w_prev + 1];
                    }
                    const int gr = (r0 >> lvl) + cr, gc = (c0 >> lvl) + cc;
                    const bool inside = (gr << lvl) < a.n && (gc << lvl) < a.m;
                    const bool sp = any || gb_split(e, a.cond, a.thresh);
                    const int k = inside ? (sp ? ksum : 1) : 0;
                    s_ext[dst + t] = e;
                    s_any[dst + t] = any;
                    s_cnt[dst + t] = k;
                    if ((gr << lvl) < a.n_pad && (gc << lvl) < a.m_pad) {
                        a.split[off + (size_t)gr * gcols + gc] = sp;
                        a.cnt[off + (size_t)gr * gcols + gc] = k;
                    }
                }
                __syncthreads();
                src = dst;
                w_prev = w;
            }
            __syncthreads();
        }
    }
    grid.sync();
    gb_mark(a, 2);

    // ---------------------------------------------------------------- phase 2: leaves before each base cell (reverse raster)
    if (blockIdx.x == 0) {
        const int nb = (a.n_pad >> a.L) * (a.m_pad >> a.L);
        const int* base_cnt = a.L == 0 ? nullptr : a.cnt + gb_level_off(a.n_pad, a.m_pad, a.L);
        const int per = (nb + GB_THREADS - 1) / GB_THREADS;
        const int lo = tid * per;                           // reversed positions [lo, lo + per): q <-> raster index nb - 1 - q
        int s = 0;
        for (int k = 0; k < per; ++k)
            if (lo + k < nb) {
                const int b = nb - 1 - (lo + k);
                s += a.L == 0 ? (a.ka0[b] & 1) : base_cnt[b];
            }
        int total;
        int run = gb_block_excl(s, s_w, total);
        for (int k = 0; k < per; ++k)
            if (lo + k < nb) {
                const int b = nb - 1 - (lo + k);
                a.base_off[b] = run;
                run += a.L == 0 ? (a.ka0[b] & 1) : base_cnt[b];
            }
        if (tid == 0) a.counts[0] = total;
    }
    grid.sync();
    gb_mark(a, 3);
    const int N = a.counts[0];

    // ---------------------------------------------------------------- phase 3: label of every pixel, leaf rectangles
    for (long long p = gtid; p < P; p += gthreads) {
        const int r = (int)(p / a.m), c = (int)(p % a.m);
        if (a.mask && a.mask[p]) {
            a.labels[p] = -1;
            continue;
        }
        int lvl = a.L;
        int off = a.base_off[(r >> lvl) * (a.m_pad >> lvl) + (c >> lvl)];
        while (lvl > 0) {
            const int cols = a.m_pad >> lvl;
            const long long lo = gb_level_off(a.n_pad, a.m_pad, lvl);
            if (!a.split[lo + (size_t)(r >> lvl) * cols + (c >> lvl)]) break;
            const int cl = lvl - 1, ccols = a.m_pad >> cl;
            const int br = (r >> lvl) << 1, bc = (c >> lvl) << 1;
            const int dr = (r >> cl) & 1, dc = (c >> cl) & 1;
            const int mine = (dr == 1 && dc == 1) ? 0 : (dr == 0 && dc == 1) ? 1 : (dr == 1 && dc == 0) ? 2 : 3;
            int k11, k01, k10;
            if (cl == 0) {                                  // level-0 counts live in the packed plane
                k11 = a.ka0[(size_t)(br + 1) * ccols + bc + 1] & 1;
                k01 = a.ka0[(size_t)(br)*ccols + bc + 1] & 1;
                k10 = a.ka0[(size_t)(br + 1) * ccols + bc] & 1;
            } else {
                const long long clo = gb_level_off(a.n_pad, a.m_pad, cl);
                k11 = a.cnt[clo + (size_t)(br + 1) * ccols + bc + 1];
                k01 = a.cnt[clo + (size_t)(br)*ccols + bc + 1];
                k10 = a.cnt[clo + (size_t)(br + 1) * ccols + bc];
            }
            if (mine > 0) off += k11;
            if (mine > 1) off += k01;
            if (mine > 2) off += k10;
            lvl = cl;
        }
        a.labels[p] = off;
        const int s = 1 << lvl;
        const int x0 = (r >> lvl) << lvl, y0 = (c >> lvl) << lvl;
        if (r == x0 && c == y0) {
            const int hh = min(x0 + s, a.n) - x0, ww = min(y0 + s, a.m) - y0;
            a.rect[off] = make_int4(x0, y0, hh, ww);
            a.npix[off] = (float)(hh * ww);
        }
    }
    grid.sync();
    gb_mark(a, 4);

    // ---------------------------------------------------------------- phase 4a: pixel-list scan (pass 1), adjacency inserts
    {
        int lo, hi;
        gb_chunk(N, lo, hi);
        int s = 0;
        for (int v = lo + tid; v < hi; v += GB_THREADS) s += (int)a.npix[v];
        s = gb_block_sum(s, s_w);
        if (tid == 0) a.part_a[blockIdx.x] = s;
    }
    for (long long p = gtid; p < P; p += gthreads) {
        const int v = a.labels[p];
        if (v < 0) continue;
        GbSet s;
        gb_candidates(a.labels, a.n, a.m, (int)(p / a.m), (int)(p % a.m), s);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int u = s.slot[k];
            if (u == GB_SLOT_EMPTY || u == -1) continue;
            const unsigned long long key = ((unsigned long long)(unsigned)v << 32) | (unsigned)u;
            unsigned h = (unsigned)gb_mix64(key) & a.cap_mask;
            while (true) {
                const unsigned long long old = atomicCAS(&a.keys[h], GB_KEY_EMPTY, key);
                if (old == GB_KEY_EMPTY || old == key) {
                    atomicMin(&a.vals[h], (int)p * 8 + k);
                    break;
                }
                h = (h + 1) & a.cap_mask;
            }
        }
    }
    grid.sync();
    gb_mark(a, 5);

    // ---------------------------------------------------------------- phase 4b: pix_ptr; first-occurrence flags + edge scan (pass 1)
    {
        int lo, hi;
        gb_chunk(N, lo, hi);
        int carry = gb_base_of(a.part_a, s_w);
        for (int v0 = lo; v0 < hi; v0 += GB_THREADS) {
            const int v = v0 + tid;
            const int x = v < hi ? (int)a.npix[v] : 0;
            int tot;
            const int ex = gb_block_excl(x, s_w, tot);
            if (v < hi) a.pix_ptr[v] = carry + ex;
            carry += tot;
        }
        if (blockIdx.x == gridDim.x - 1 && tid == 0) a.pix_ptr[N] = carry;     // the last CTA's carry is the grand total
    }
    {
        int lo, hi;
        gb_chunk(P, lo, hi);
        int s = 0;
        for (int p = lo + tid; p < hi; p += GB_THREADS) {
            const int v = a.labels[p];
            unsigned bits = 0;
            if (v >= 0) {
                GbSet st;
                gb_candidates(a.labels, a.n, a.m, p / a.m, p % a.m, st);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int u = st.slot[k];
                    if (u == GB_SLOT_EMPTY || u == -1) continue;
                    const unsigned long long key = ((unsigned long long)(unsigned)v << 32) | (unsigned)u;
                    unsigned h = (unsigned)gb_mix64(key) & a.cap_mask;
                    while (a.keys[h] != key) h = (h + 1) & a.cap_mask;
                    if (a.vals[h] == p * 8 + k) bits |= 1u << k;
                }
            }
            a.emit[p] = (uint8_t)bits;
            const int c = __popc(bits);
            a.count[p] = c;
            s += c;
        }
        s = gb_block_sum(s, s_w);
        if (tid == 0) a.part_b[blockIdx.x] = s;
    }
    grid.sync();
    gb_mark(a, 6);

    // ---------------------------------------------------------------- phase 5: pixel lists; edge offsets + emission
    for (long long p = gtid; p < P; p += gthreads) {
        const int v = a.labels[p];
        if (v < 0) continue;
        const int4 rc = a.rect[v];
        const int r = (int)(p / a.m), c = (int)(p % a.m);
        a.pix_idx[a.pix_ptr[v] + (r - rc.x) * rc.w + (c - rc.y)] = (int)p;
    }
    {
        int lo, hi;
        gb_chunk(P, lo, hi);
        int carry = gb_base_of(a.part_b, s_w);
        for (int p0 = lo; p0 < hi; p0 += GB_THREADS) {
            const int p = p0 + tid;
            const int x = p < hi ? a.count[p] : 0;
            int tot;
            int e = carry + gb_block_excl(x, s_w, tot);
            carry += tot;
            if (p < hi && x) {
                const unsigned bits = a.emit[p];
                const int v = a.labels[p];
                GbSet st;
                gb_candidates(a.labels, a.n, a.m, p / a.m, p % a.m, st);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (bits & (1u << k)) {
                        a.src32[e] = v;
                        a.dst32[e] = st.slot[k];
                        ++e;
                    }
            }
        }
        if (blockIdx.x == gridDim.x - 1 && tid == 0) a.counts[1] = carry;
    }
    grid.sync();
    gb_mark(a, 7);
    const int E = a.counts[1];
    const int Cd = a.C + 1;                                 // data row: pooled channels | cell size

    // ---------------------------------------------------------------- phase 6: pooled node features (pool.cu order), edge_index, CSR counts
    {
        const int lane = tid & 31;
        const long long warp = gtid >> 5, nwarps = gthreads >> 5;
        // a warp takes 32 consecutive leaves of one frame.  Leaves of <= 4 pixels (all of a pixel-level mesh, most of any quadtree)
        // are summed lane-serially -- lane l owns leaf v0 + l and adds its pixels as the butterfly would, (p0 + p2) + (p1 + p3) with
        // absent partial sums +0 -- so 32 leaves are in flight per warp instead of one; larger leaves are then taken one at a time
        // by the whole warp.
        const int groups = (N + 31) / 32;
        for (long long w = warp; w < (long long)a.T * groups; w += nwarps) {
            const int b = (int)(w / groups), v0 = (int)(w - (long long)b * groups) * 32;
            const float* src = a.img + (size_t)b * P * a.C;
            const int vl = v0 + lane;
            int4 rcl = make_int4(0, 0, 0, 0);
            if (vl < N) rcl = a.rect[vl];
            const int cntl = rcl.z * rcl.w;
            if (vl < N && cntl <= 4) {
                const float np = a.npix[vl];
                float* orow = a.data + ((size_t)b * N + vl) * Cd;
                const float* px[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int kk = k < cntl ? k : 0;
                    const int rr = kk / rcl.w;
                    px[k] = src + ((size_t)(rcl.x + rr) * a.m + rcl.y + (kk - rr * rcl.w)) * a.C;
                }
                for (int c = 0; c < a.C; ++c) {
                    float pk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) pk[k] = (k < cntl) ? 0.f + px[k][c] : 0.f;
                    orow[c] = ((pk[0] + pk[2]) + (pk[1] + pk[3])) / np;
                }
                orow[a.C] = np / a.size_div;
            }
            // larger leaves: queued (once, by the warps of frame 0) and summed below, a warp per (frame, leaf)
            if (b == 0 && vl < N && cntl > 4) a.tmp_in[atomicAdd(&a.counts[3], 1)] = vl;
        }
    }
    for (long long e = gtid; e < E; e += gthreads) {
        const int s = a.src32[e], d = a.dst32[e];
        a.ei64[e] = s;
        a.ei64[(size_t)E + e] = d;
        atomicAdd(&a.cnt_in[d], 1);
        atomicAdd(&a.cnt_out[s], 1);
    }
    grid.sync();                                            // the queue of larger leaves is complete
    {
        const int lane = tid & 31;
        const long long warp = gtid >> 5, nwarps = gthreads >> 5;
        const int n_large = a.counts[3];
        for (long long w = warp; w < (long long)a.T * n_large; w += nwarps) {
            const int b = (int)(w / n_large), v = a.tmp_in[w - (long long)b * n_large];
            const float* src = a.img + (size_t)b * P * a.C;
            // the pixels of a leaf in raster order come straight from its rectangle (no pixel-list indirection: every load
            // address is known up front, so the loads of several rounds are in flight together; the additions stay in the
            // defined order).  (Eight rounds in flight, float4 loads, or a warp per group of four channels did not measure
            // faster: the phase stays at 35 - 65 us for the 64 x 64 leaves of the ice grid, run-to-run spread included.)
            const int4 rc = a.rect[v];
            const int cnt = rc.z * rc.w;
            const float np = a.npix[v];
            float* orow = a.data + ((size_t)b * N + v) * Cd;
            for (int c0 = 0; c0 < a.C; c0 += 8) {
                const int cw = a.C - c0 < 8 ? a.C - c0 : 8;
                float acc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = 0.f;
                for (int k0 = lane; k0 < cnt; k0 += 128) {
                    float x[4][8];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + 32 * u;
                        const bool ok = k < cnt;
                        const int rr = ok ? k / rc.w : 0;
                        const float* px = src + ((size_t)(rc.x + rr) * a.m + rc.y + (ok ? k - rr * rc.w : 0)) * a.C + c0;
#pragma unroll
                        for (int c = 0; c < 8; ++c) x[u][c] = (ok && c < cw) ? px[c] : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (k0 + 32 * u < cnt) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) acc[c] += x[u][c];
                        }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
                }
                float mine = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (lane == c) mine = acc[c];
                if (lane < cw) orow[c0 + lane] = mine / np;
            }
            if (lane == 0) orow[a.C] = np / a.size_div;
        }
    }
    grid.sync();
    gb_mark(a, 8);

    // ---------------------------------------------------------------- phase 7: edge attributes; CSR row-pointer scans (pass 1)
    for (long long e = gtid; e < E; e += gthreads) {
        const int s = a.src32[e], d = a.dst32[e];
        const float* ps = a.data + (size_t)s * Cd + (a.C - 2);      // frame 0: ii, jj of the node
        const float* pd = a.data + (size_t)d * Cd + (a.C - 2);
        const float w_img = (float)a.m, h_img = (float)a.n;
        const float xs = __fmul_rn(__fmul_rn(ps[0], w_img), a.resolution), xd = __fmul_rn(__fmul_rn(pd[0], w_img), a.resolution);
        const float ys = __fmul_rn(__fmul_rn(ps[1], h_img), a.resolution), yd = __fmul_rn(__fmul_rn(pd[1], h_img), a.resolution);
        const float dx = __fsub_rn(xs, xd), dy = __fsub_rn(ys, yd);
        const float dist = sqrtf(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
        if (a.two_cols) {
            const float two_pi = 6.283185307179586f;
            float r = fmodf(atan2f(dx, dy), two_pi);
            if (r != 0.f && r < 0.f) r += two_pi;
            a.edge_attrs[(size_t)e * 2] = r / two_pi;
            a.edge_attrs[(size_t)e * 2 + 1] = dist;
        } else {
            a.edge_attrs[e] = dist;
        }
    }
    {
        int lo, hi;
        gb_chunk(N + 1, lo, hi);
        int si = 0, so = 0;
        for (int v = lo + tid; v < hi; v += GB_THREADS) {
            si += a.cnt_in[v];
            so += a.cnt_out[v];
        }
        si = gb_block_sum(si, s_w);
        so = gb_block_sum(so, s_w);
        if (tid == 0) {
            a.part_a[blockIdx.x] = si;
            a.part_b[blockIdx.x] = so;
        }
    }
    grid.sync();
    gb_mark(a, 9);

    // ---------------------------------------------------------------- phase 8: in_ptr / out_ptr (+ cursors)
    {
        int lo, hi;
        gb_chunk(N + 1, lo, hi);
        int ci = gb_base_of(a.part_a, s_w), co = gb_base_of(a.part_b, s_w);
        for (int v0 = lo; v0 < hi; v0 += GB_THREADS) {
            const int v = v0 + tid;
            const int xi = v < hi ? a.cnt_in[v] : 0, xo = v < hi ? a.cnt_out[v] : 0;
            int ti, to;
            const int ei = gb_block_excl(xi, s_w, ti);
            const int eo = gb_block_excl(xo, s_w, to);
            if (v < hi) {
                a.in_ptr[v] = ci + ei;
                a.cur_in[v] = ci + ei;
                a.out_ptr[v] = co + eo;
                a.cur_out[v] = co + eo;
            }
            ci += ti;
            co += to;
        }
    }
    grid.sync();
    gb_mark(a, 10);

    // ---------------------------------------------------------------- phase 9: rows filled in arrival order ...
    for (long long e = gtid; e < E; e += gthreads) {
        a.tmp_in[atomicAdd(&a.cur_in[a.dst32[e]], 1)] = (int)e;
        a.tmp_out[atomicAdd(&a.cur_out[a.src32[e]], 1)] = (int)e;
    }
    grid.sync();
    gb_mark(a, 11);

    // ---------------------------------------------------------------- phase 10: ... then ordered by edge id (fixed summation order):
    // a warp per row, every element's rank = the number of smaller ids in its row (rows are short: a leaf's perimeter)
    {
        const int lane = tid & 31;
        const long long warp = gtid >> 5, nwarps = gthreads >> 5;
        // a warp takes 32 consecutive rows of one CSR: rows of <= 8 entries (a small leaf's neighbours) are ranked by their own lane,
        // longer rows afterwards by the whole warp
        const int groups = (N + 31) / 32;
        for (long long t = warp; t < 2ll * groups; t += nwarps) {
            const int which = (int)(t & 1), v0 = (int)(t >> 1) * 32, vl = v0 + lane;
            const int* ptr = which ? a.out_ptr : a.in_ptr;
            const int* tmp = which ? a.tmp_out : a.tmp_in;
            int* eid = which ? a.eid_out : a.in_eid;
            const int lol = vl < N ? ptr[vl] : 0, rl = vl < N ? ptr[vl + 1] - lol : 0;
            if (rl > 0 && rl <= 8) {
                int x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = i < rl ? tmp[lol + i] : INT_MAX;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < rl) {
                        int rank = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) rank += (x[j] < x[i]);
                        eid[lol + rank] = x[i];
                    }
                }
            }
            unsigned rest = __ballot_sync(0xffffffffu, rl > 8);
            while (rest) {
                const int n = __ffs(rest) - 1;
                rest &= rest - 1;
                const int lo = __shfl_sync(0xffffffffu, lol, n), r = __shfl_sync(0xffffffffu, rl, n);
                for (int i = lane; i < r; i += 32) {
                    const int x = tmp[lo + i];
                    int rank = 0;
                    for (int j = 0; j < r; ++j) rank += (tmp[lo + j] < x);
                    eid[lo + rank] = x;
                }
            }
        }
    }
    grid.sync();
    gb_mark(a, 12);

    // ---------------------------------------------------------------- phase 11: in-CSR payload
    {
        const int width = a.two_cols ? 2 : 1;
        for (long long k = gtid; k < E; k += gthreads) {
            const int e = a.in_eid[k];
            a.in_src[k] = a.src32[e];
            a.kin_of_edge[e] = (int)k;
            for (int c = 0; c < width; ++c) a.edge_attr_in[(size_t)k * width + c] = a.edge_attrs[(size_t)e * width + c];
        }
    }
    grid.sync();
    gb_mark(a, 13);

    // ---------------------------------------------------------------- phase 12: out-CSR payload
    for (long long k = gtid; k < E; k += gthreads) {
        const int e = a.eid_out[k];
        a.out_dst[k] = a.dst32[e];
        a.out_kin[k] = a.kin_of_edge[e];
    }
    gb_mark(a, 14);
}

struct GbScratch {
    size_t e0, ka0, cnt, split, base_off, rect, keys, vals, emit, count, part_a, part_b, cnt_in, cnt_out, cur_in, cur_out, eid_out,
        kin_of_edge, tmp_in, tmp_out, counts;
    // results at capacity (P pixels, e_cap = 4 P edges); qmp_quadtree_graph_export copies their prefixes out
    size_t labels, npix, pix_ptr, pix_idx, data, ei64, src32, dst32, edge_attrs, in_ptr, in_src, in_eid, out_ptr, out_dst, out_kin,
        edge_attr_in;
    size_t total;
};

static long long gb_table_cap(long long P) {
    long long cap = 1024;
    while (cap < 8 * P) cap <<= 1;
    return cap;
}

static GbScratch gb_layout(int n, int m, int max_size, int T, int C) {
    const long long n_pad = (n + max_size - 1) / max_size * max_size, m_pad = (m + max_size - 1) / max_size * max_size;
    const long long P = (long long)n * m, np = n_pad * m_pad, e_cap = 4 * P;
    int L = 0;
    while ((1 << L) < max_size) ++L;
    long long cells = 0;
    for (int k = 0; k <= L; ++k) cells += (n_pad >> k) * (m_pad >> k);
    const long long nb = (n_pad >> L) * (m_pad >> L), cap = gb_table_cap(P);
    GbScratch s{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) & ~(size_t)255;
        return at;
    };
    s.counts = take(256);                                   // 4 counts | phase timeline (gb_mark)
    s.e0 = take(4 * np);
    s.ka0 = take(4 * np);
    s.cnt = take(4 * cells);
    s.split = take(cells);
    s.base_off = take(4 * nb);
    s.rect = take(16 * P);
    s.keys = take(8 * cap);
    s.vals = take(4 * cap);
    s.emit = take(P);
    s.count = take(4 * P);
    s.part_a = take(4 * 4096);
    s.part_b = take(4 * 4096);
    s.cnt_in = take(4 * (P + 2));
    s.cnt_out = take(4 * (P + 2));
    s.cur_in = take(4 * (P + 2));
    s.cur_out = take(4 * (P + 2));
    s.eid_out = take(4 * e_cap);
    s.kin_of_edge = take(4 * e_cap);
    s.tmp_in = take(4 * e_cap);
    s.tmp_out = take(4 * e_cap);
    s.labels = take(4 * P);
    s.npix = take(4 * P);
    s.pix_ptr = take(4 * (P + 1));
    s.pix_idx = take(4 * P);
    s.data = take(4 * (size_t)T * P * (C + 1));
    s.ei64 = take(8 * 2 * e_cap);
    s.src32 = take(4 * e_cap);
    s.dst32 = take(4 * e_cap);
    s.edge_attrs = take(4 * 2 * e_cap);
    s.in_ptr = take(4 * (P + 1));
    s.in_src = take(4 * e_cap);
    s.in_eid = take(4 * e_cap);
    s.out_ptr = take(4 * (P + 1));
    s.out_dst = take(4 * e_cap);
    s.out_kin = take(4 * e_cap);
    s.edge_attr_in = take(4 * 2 * e_cap);
    s.total = off;
    return s;
}

// one launch copies the result prefixes out of the arena: up to 16 segments (source, destination, 4-byte words)
struct GbExport {
    const uint32_t* src[16];
    uint32_t* dst[16];
    long long words[16];
    int n;
};
__global__ void __launch_bounds__(256) gb_export_kernel(const GbExport x) {
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;
    for (int s = 0; s < x.n; ++s)
        for (long long i = gtid; i < x.words[s]; i += gthreads) x.dst[s][i] = x.src[s][i];
}

}  // namespace qmp
using namespace qmp;

// Bytes of the arena qmp_quadtree_graph works in for T frames of an n x m x C image (scratch AND the results at capacity;
// reusable across calls on one stream -- qmp_quadtree_graph_export copies a result out before the next build overwrites it).
QMP_API long long qmp_quadtree_graph_scratch_bytes(int n, int m, int max_size, int T, int C) {
    return (long long)gb_layout(n, m, max_size, T, C).total;
}

// The whole quadtree graph build of image_to_graph in one cooperative launch + one read-back.
//   img [T, n, m, C] frames (pooled into the node features; channels C-2, C-1 are the positional encoding), crit
//   [n_pad, m_pad] the criterion image (qmp_frame_max_pad, then the caller's transform), mask / hir [n, m] uint8 or NULL.
// Everything is produced inside `arena` (qmp_quadtree_graph_scratch_bytes), compacted: labels int32 [P]; npix [N]; pix_ptr
// [N+1]; pix_idx [<= P]; data [T, N, C+1] (last column npix / (max_size/2)^2); edge_index int64 [2, E]; src32 / dst32 [E];
// edge_attrs [E, 2] (two_cols) or [E]; in_ptr / out_ptr [N+1]; in_src / in_eid / out_dst / out_kin [E]; edge_attr_in =
// edge_attrs in in-CSR order.  counts_host (pinned int32 [4]) receives (N, E, number of NaNs in img, 0) and the call returns
// after the stream has drained -- the one host synchronisation of a mesh build.  max_size <= 64 (larger leaves: the per-kernel
// entry points).
QMP_API int qmp_quadtree_graph(const float* img, int T, int n, int m, int C, const float* crit, const uint8_t* mask, const uint8_t* hir,
                               int max_size, int cond, double thresh, float resolution, int two_cols, int* counts_host, void* arena,
                               void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QMP_REQUIRE(max_size > 0 && (max_size & (max_size - 1)) == 0 && max_size <= 64, "qmp_quadtree_graph: max_size must be a power of two <= 64");
    QMP_REQUIRE(cond >= 0 && cond < 4, "qmp_quadtree_graph: unknown condition");
    QMP_REQUIRE(T > 0 && n > 0 && m > 0 && C >= 2, "qmp_quadtree_graph: bad shape");
    QMP_REQUIRE((long long)n * m * 8 < INT_MAX, "qmp_quadtree_graph: image too large");
    const GbScratch lay = gb_layout(n, m, max_size, T, C);
    uint8_t* base = reinterpret_cast<uint8_t*>(arena);
    GbArgs a{};
    a.img = img; a.T = T; a.n = n; a.m = m; a.C = C;
    a.crit = crit; a.mask = mask; a.hir = hir;
    a.n_pad = (n + max_size - 1) / max_size * max_size;
    a.m_pad = (m + max_size - 1) / max_size * max_size;
    a.row_cap = a.n_pad < a.m_pad ? a.n_pad : a.m_pad;
    QMP_REQUIRE(n <= a.m_pad, "qmp_quadtree_graph: image taller than its padded width (the reference reads out of bounds)");
    a.L = 0;
    while ((1 << a.L) < max_size) ++a.L;
    a.cond = cond; a.thresh = thresh; a.resolution = resolution; a.two_cols = two_cols;
    a.size_div = (float)((max_size / 2.0) * (max_size / 2.0));
    a.e_cap = 4ll * n * m;
#define GB_AT(T_, f) reinterpret_cast<T_*>(base + lay.f)
    a.labels = GB_AT(int, labels); a.npix = GB_AT(float, npix); a.pix_ptr = GB_AT(int, pix_ptr); a.pix_idx = GB_AT(int, pix_idx);
    a.data = GB_AT(float, data); a.ei64 = GB_AT(long long, ei64); a.src32 = GB_AT(int, src32); a.dst32 = GB_AT(int, dst32);
    a.edge_attrs = GB_AT(float, edge_attrs); a.in_ptr = GB_AT(int, in_ptr); a.in_src = GB_AT(int, in_src); a.in_eid = GB_AT(int, in_eid);
    a.out_ptr = GB_AT(int, out_ptr); a.out_dst = GB_AT(int, out_dst); a.out_kin = GB_AT(int, out_kin);
    a.edge_attr_in = GB_AT(float, edge_attr_in);
    a.counts = GB_AT(int, counts); a.e0 = GB_AT(float, e0); a.ka0 = GB_AT(int, ka0); a.cnt = GB_AT(int, cnt); a.split = GB_AT(uint8_t, split);
    a.base_off = GB_AT(int, base_off); a.rect = GB_AT(int4, rect); a.keys = GB_AT(unsigned long long, keys); a.vals = GB_AT(int, vals);
    a.table_cap = gb_table_cap((long long)n * m);
    a.cap_mask = (unsigned)(a.table_cap - 1);
    a.emit = GB_AT(uint8_t, emit); a.count = GB_AT(int, count); a.part_a = GB_AT(int, part_a); a.part_b = GB_AT(int, part_b);
    a.cnt_in = GB_AT(int, cnt_in); a.cnt_out = GB_AT(int, cnt_out); a.cur_in = GB_AT(int, cur_in); a.cur_out = GB_AT(int, cur_out);
    a.eid_out = GB_AT(int, eid_out); a.kin_of_edge = GB_AT(int, kin_of_edge);
    a.tmp_in = GB_AT(int, tmp_in); a.tmp_out = GB_AT(int, tmp_out);

    static int grid = 0;
    if (grid == 0) {
        int dev = 0, n_sm = 0, per_sm = 0;
        QMP_CUDA(cudaGetDevice(&dev));
        QMP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QMP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quadtree_graph_kernel, GB_THREADS, 0));
        QMP_REQUIRE(per_sm >= 1, "qmp_quadtree_graph: the kernel does not fit on an SM");
        grid = n_sm * (per_sm < 2 ? per_sm : 2);
        if (grid > 4096) grid = 4096;
    }
    QMP_CUDA(cudaMemsetAsync(a.counts, 0, 16, st));
    void* kargs[] = {(void*)&a};
    QMP_CUDA(cudaLaunchCooperativeKernel((const void*)quadtree_graph_kernel, dim3(grid), dim3(GB_THREADS), kargs, 0, st));
    QMP_LAUNCH_CHECK("quadtree_graph_kernel");
    if (counts_host) {
        QMP_CUDA(cudaMemcpyAsync(counts_host, a.counts, 16, cudaMemcpyDeviceToHost, st));
        QMP_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// Copy the result of the last qmp_quadtree_graph on `arena` into exact-size buffers (one launch), N and E as read back:
//   ipack int32: labels [P] | pix_ptr [N+1] | pix_idx [P] | src32 [E] | dst32 [E] | in_ptr [N+1] | in_src [E] | in_eid [E] |
//                out_ptr [N+1] | out_dst [E] | out_kin [E]
//   fpack float: npix [N] | data [T, N, C+1] | edge_attrs [E, w] | edge_attr_in [E, w]      (w = 2 with two_cols, else 1)
//   edge_index int64 [2, E]
// every segment starts at a multiple of 4 elements (16-byte rows for the vector loads of the conv kernels).
QMP_API int qmp_quadtree_graph_export(const void* arena, int n, int m, int max_size, int T, int C, int two_cols, int N, int E,
                                      int* ipack, float* fpack, long long* edge_index, void* stream) {
    const GbScratch lay = gb_layout(n, m, max_size, T, C);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(arena);
    const long long P = (long long)n * m, w = two_cols ? 2 : 1;
    GbExport x{};
    auto r4 = [](long long v) { return (v + 3) & ~3ll; };
    long long io = 0, fo = 0;
    auto seg_i = [&](size_t at, long long words) {
        x.src[x.n] = reinterpret_cast<const uint32_t*>(base + at);
        x.dst[x.n] = reinterpret_cast<uint32_t*>(ipack + io);
        x.words[x.n] = words;
        io += r4(words);
        ++x.n;
    };
    auto seg_f = [&](size_t at, long long words) {
        x.src[x.n] = reinterpret_cast<const uint32_t*>(base + at);
        x.dst[x.n] = reinterpret_cast<uint32_t*>(fpack + fo);
        x.words[x.n] = words;
        fo += r4(words);
        ++x.n;
    };
    seg_i(lay.labels, P); seg_i(lay.pix_ptr, N + 1); seg_i(lay.pix_idx, P); seg_i(lay.src32, E); seg_i(lay.dst32, E);
    seg_i(lay.in_ptr, N + 1); seg_i(lay.in_src, E); seg_i(lay.in_eid, E); seg_i(lay.out_ptr, N + 1); seg_i(lay.out_dst, E);
    seg_i(lay.out_kin, E);
    seg_f(lay.npix, N); seg_f(lay.data, (long long)T * N * (C + 1)); seg_f(lay.edge_attrs, E * w); seg_f(lay.edge_attr_in, E * w);
    x.src[x.n] = reinterpret_cast<const uint32_t*>(base + lay.ei64);
    x.dst[x.n] = reinterpret_cast<uint32_t*>(edge_index);
    x.words[x.n] = 4ll * E;                                 // 2 E int64
    ++x.n;
    gb_export_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(x);
    QMP_LAUNCH_CHECK("gb_export_kernel");
    return 0;
}
