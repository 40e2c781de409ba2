// K1: quadtree split decision pyramid + label assignment, bit-exact against the reference's
// LIFO-stack traversal (reference model/graph_functions.py:145-259; oracle/graph_ref.py:quadtree_labels).
//
// The reference pops cells from a stack; whether a cell splits depends only on the data inside its
// (size+1)^2 window, and that window is the union of its four children's windows, so every split
// flag is computed bottom-up in parallel (a max/min + "any mask/HIR" pyramid).  A leaf's label is
// its rank in the reference's pop order = reverse raster over base cells, children visited
// (x+s,y+s) (x,y+s) (x+s,y) (x,y); the rank comes from per-cell leaf counts (bottom-up) and a
// per-pixel top-down walk that adds the counts of the siblings visited first.
#include "common.cuh"

namespace qmp {

// ---- criterion frame: max over time of channel 0, edge-replicated to the padded extent
// (graph_functions.py:632 and :190)
__global__ void frame_max_pad_kernel(const float* __restrict__ x, int T, int H, int W, int C, int n_pad, int m_pad,
                                     float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pad * m_pad) return;
    const int r = min(idx / m_pad, H - 1), c = min(idx % m_pad, W - 1);
    const float* p = x + ((size_t)r * W + c) * C;
    const size_t tstride = (size_t)H * W * C;
    float v = p[0];
    for (int t = 1; t < T; ++t) v = fmaxf(v, p[t * tstride]);
    out[idx] = v;
}

struct QtParams {
    int n, m;          // image rows, cols
    int n_pad, m_pad;  // padded to multiples of max_size
    int row_cap;       // min(n_pad, m_pad): the reference clips both axes with shape[1]
    int L;             // log2(max_size)
    int cond;          // 0 max>, 1 max<, 2 min>, 3 min<
    double thresh;
};

__device__ __forceinline__ bool crit_split(float e, int cond, double thresh) {
    // reference compares the float32 extreme with a float64 threshold (numba promotes)
    const double d = (double)e;
    return (cond == 0) ? (d > thresh) : (cond == 1) ? (d < thresh) : (cond == 2) ? (d > thresh) : (d < thresh);
}

__device__ __forceinline__ float ext_op(float a, float b, bool is_max) { return is_max ? fmaxf(a, b) : fminf(a, b); }

__host__ __device__ inline long long level_offset(int n_pad, int m_pad, int lvl) {
    long long off = 0;
    for (int k = 0; k < lvl; ++k) off += (long long)(n_pad >> k) * (m_pad >> k);
    return off;
}

// One CTA per 64 x 64 tile; levels 0..LT = min(log2(max_size), 6) in shared memory.
constexpr int QT_TS = 64;
__global__ void __launch_bounds__(256) qt_tile_kernel(const float* __restrict__ crit, const uint8_t* __restrict__ mask,
                                                      const uint8_t* __restrict__ hir, QtParams P, int ts, int LT,
                                                      uint8_t* __restrict__ split, int* __restrict__ cnt,
                                                      float* __restrict__ ext_top, uint8_t* __restrict__ any_top) {
    __shared__ float s_ext[QT_TS * QT_TS + QT_TS * QT_TS / 4];
    __shared__ int s_cnt[QT_TS * QT_TS + QT_TS * QT_TS / 4];
    __shared__ uint8_t s_any[QT_TS * QT_TS + QT_TS * QT_TS / 4];
    const bool is_max = P.cond < 2;
    const float neutral = is_max ? -FLT_MAX : FLT_MAX;
    const int r0 = blockIdx.y * ts, c0 = blockIdx.x * ts;

    // level 0: the 2x2 window of every pixel (size 1 -> size+1 = 2)
    for (int t = threadIdx.x; t < ts * ts; t += blockDim.x) {
        const int r = r0 + t / ts, c = c0 + t % ts;
        float e = neutral;
        bool a = false;
        int k = 0;
        if (r < P.n_pad && c < P.m_pad) {
            const int r1 = min(r + 2, P.row_cap), c1 = min(c + 2, P.m_pad);
            for (int rr = r; rr < r1; ++rr)
                for (int cc = c; cc < c1; ++cc) {
                    e = ext_op(e, crit[(size_t)rr * P.m_pad + cc], is_max);
                    if (rr < P.n && cc < P.m) {
                        if (mask && mask[(size_t)rr * P.m + cc]) a = true;
                        if (hir && hir[(size_t)rr * P.m + cc]) a = true;
                    }
                }
            k = (r < P.n && c < P.m && !(mask && mask[(size_t)r * P.m + c])) ? 1 : 0;
            cnt[(size_t)r * P.m_pad + c] = k;  // level 0 lives at offset 0
        }
        s_ext[t] = e;
        s_any[t] = a;
        s_cnt[t] = k;
    }
    __syncthreads();

    int src = 0, dst = QT_TS * QT_TS, w_prev = ts;
    for (int lvl = 1; lvl <= LT; ++lvl) {
        const int w = ts >> lvl;
        const long long off = level_offset(P.n_pad, P.m_pad, lvl);
        const int gcols = P.m_pad >> lvl;
        for (int t = threadIdx.x; t < w * w; t += blockDim.x) {
            const int cr = t / w, cc = t % w;
            const int a00 = src + (2 * cr) * w_prev + 2 * cc;
            const float e = ext_op(ext_op(s_ext[a00], s_ext[a00 + 1], is_max),
                                   ext_op(s_ext[a00 + w_prev], s_ext[a00 + w_prev + 1], is_max), is_max);
            const bool a = s_any[a00] | s_any[a00 + 1] | s_any[a00 + w_prev] | s_any[a00 + w_prev + 1];
            const int ksum = s_cnt[a00] + s_cnt[a00 + 1] + s_cnt[a00 + w_prev] + s_cnt[a00 + w_prev + 1];
            const int gr = (r0 >> lvl) + cr, gc = (c0 >> lvl) + cc;  // global cell coords at this level
            const bool inside = (gr << lvl) < P.n && (gc << lvl) < P.m;
            const bool sp = a || crit_split(e, P.cond, P.thresh);
            const int k = inside ? (sp ? ksum : 1) : 0;
            s_ext[dst + t] = e;
            s_any[dst + t] = a;
            s_cnt[dst + t] = k;
            if ((gr << lvl) < P.n_pad && (gc << lvl) < P.m_pad) {
                split[off + (size_t)gr * gcols + gc] = sp;
                cnt[off + (size_t)gr * gcols + gc] = k;
                if (lvl == LT && ext_top) {
                    ext_top[(size_t)gr * gcols + gc] = e;
                    any_top[(size_t)gr * gcols + gc] = a;
                }
            }
        }
        __syncthreads();
        // ping-pong: next level reads what was just written
        const int tmp = src;
        src = dst;
        dst = tmp;
        w_prev = w;
    }
}

// Levels above the tile (max_size > 64): one thread per parent cell, children read from global.
__global__ void qt_level_up_kernel(QtParams P, int lvl, const float* __restrict__ ext_in, const uint8_t* __restrict__ any_in,
                                   float* __restrict__ ext_out, uint8_t* __restrict__ any_out, uint8_t* __restrict__ split,
                                   int* __restrict__ cnt) {
    const int rows = P.n_pad >> lvl, cols = P.m_pad >> lvl;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    const bool is_max = P.cond < 2;
    const int gr = t / cols, gc = t % cols, ccols = cols * 2;
    const size_t a00 = (size_t)(2 * gr) * ccols + 2 * gc;
    const float e = ext_op(ext_op(ext_in[a00], ext_in[a00 + 1], is_max),
                           ext_op(ext_in[a00 + ccols], ext_in[a00 + ccols + 1], is_max), is_max);
    const bool a = any_in[a00] | any_in[a00 + 1] | any_in[a00 + ccols] | any_in[a00 + ccols + 1];
    const long long offc = level_offset(P.n_pad, P.m_pad, lvl - 1), off = level_offset(P.n_pad, P.m_pad, lvl);
    const int ksum = cnt[offc + a00] + cnt[offc + a00 + 1] + cnt[offc + a00 + ccols] + cnt[offc + a00 + ccols + 1];
    const bool inside = (gr << lvl) < P.n && (gc << lvl) < P.m;
    const bool sp = a || crit_split(e, P.cond, P.thresh);
    ext_out[t] = e;
    any_out[t] = a;
    split[off + t] = sp;
    cnt[off + t] = inside ? (sp ? ksum : 1) : 0;
}

// Exclusive SUFFIX sum over base cells in raster order (= prefix in the reference's pop order).
__global__ void __launch_bounds__(1024) qt_base_scan_kernel(const int* __restrict__ base_cnt, int nb, int* __restrict__ base_off,
                                                            int* __restrict__ n_nodes) {
    __shared__ int part[1024];
    const int per = (nb + 1023) / 1024;
    // thread t owns reversed positions [t*per, t*per+per): reversed position q <-> raster index nb-1-q
    const int lo = threadIdx.x * per;
    int s = 0;
    for (int k = 0; k < per; ++k)
        if (lo + k < nb) s += base_cnt[nb - 1 - (lo + k)];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        int t = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int k = 0; k < per; ++k)
        if (lo + k < nb) {
            const int b = nb - 1 - (lo + k);
            base_off[b] = run;
            run += base_cnt[b];
        }
    if (threadIdx.x == 1023) *n_nodes = part[1023];
}

// Per-pixel top-down walk.  Also records each leaf's rectangle at its origin pixel.
__global__ void qt_assign_kernel(QtParams P, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ split,
                                 const int* __restrict__ cnt, const int* __restrict__ base_off, int* __restrict__ labels,
                                 int4* __restrict__ node_rect, float* __restrict__ npix) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n * P.m) return;
    const int r = p / P.m, c = p % P.m;
    if (mask && mask[p]) {
        labels[p] = -1;
        return;
    }
    int lvl = P.L;
    int off = base_off[(r >> lvl) * (P.m_pad >> lvl) + (c >> lvl)];
    while (lvl > 0) {
        const int cols = P.m_pad >> lvl;
        const long long lo = level_offset(P.n_pad, P.m_pad, lvl);
        if (!split[lo + (size_t)(r >> lvl) * cols + (c >> lvl)]) break;
        // descend: children visited in order (1,1) (0,1) (1,0) (0,0) as (row-half, col-half)
        const int cl = lvl - 1, ccols = P.m_pad >> cl;
        const long long clo = level_offset(P.n_pad, P.m_pad, cl);
        const int br = (r >> lvl) << 1, bc = (c >> lvl) << 1;
        const int dr = (r >> cl) & 1, dc = (c >> cl) & 1;
        const int mine = (dr == 1 && dc == 1) ? 0 : (dr == 0 && dc == 1) ? 1 : (dr == 1 && dc == 0) ? 2 : 3;
        const int k11 = cnt[clo + (size_t)(br + 1) * ccols + bc + 1];
        const int k01 = cnt[clo + (size_t)(br)*ccols + bc + 1];
        const int k10 = cnt[clo + (size_t)(br + 1) * ccols + bc];
        if (mine > 0) off += k11;
        if (mine > 1) off += k01;
        if (mine > 2) off += k10;
        lvl = cl;
    }
    labels[p] = off;
    const int s = 1 << lvl;
    const int x0 = (r >> lvl) << lvl, y0 = (c >> lvl) << lvl;
    if (r == x0 && c == y0) {
        const int hh = min(x0 + s, P.n) - x0, ww = min(y0 + s, P.m) - y0;
        node_rect[off] = make_int4(x0, y0, hh, ww);
        npix[off] = (float)(hh * ww);
    }
}

}  // namespace qmp

using namespace qmp;

// crit_out[n_pad, m_pad] = edge-padded max over T of x[T,H,W,C][..., 0]
QMP_API int qmp_frame_max_pad(const float* x, int T, int H, int W, int C, int n_pad, int m_pad, float* crit_out,
                              void* stream) {
    QMP_REQUIRE(T > 0 && H > 0 && W > 0 && C > 0 && n_pad >= H && m_pad >= W, "qmp_frame_max_pad: bad shape");
    frame_max_pad_kernel<<<cdiv((long long)n_pad * m_pad, 256), 256, 0, (cudaStream_t)stream>>>(x, T, H, W, C, n_pad,
                                                                                                 m_pad, crit_out);
    QMP_LAUNCH_CHECK("qmp_frame_max_pad");
    return 0;
}

QMP_API long long qmp_quadtree_pyramid_cells(int n, int m, int max_size) {
    const int n_pad = (n + max_size - 1) / max_size * max_size, m_pad = (m + max_size - 1) / max_size * max_size;
    int L = 0;
    while ((1 << L) < max_size) ++L;
    return level_offset(n_pad, m_pad, L + 1);
}

// crit: [n_pad, m_pad] float32 criterion image (already padded and transformed).
// mask / hir: [n, m] uint8 or NULL.  cond: index into CONDITIONS.  Outputs: labels int32 [n, m]
// (-1 masked), node_rect int4 [<= n*m] (x0, y0, rows, cols), npix float [<= n*m], n_nodes int (device).
// Scratch: split uint8 [cells], cnt int32 [cells] with cells = qmp_quadtree_pyramid_cells();
// base_off int32 [(n_pad/max_size)*(m_pad/max_size)]; top_f float/top_b uint8 [2 * (n_pad/64)*(m_pad/64)]
// (only touched when max_size > 64).
QMP_API int qmp_quadtree_labels(const float* crit, const uint8_t* mask, const uint8_t* hir, int n, int m, int max_size,
                                int cond, double thresh, int* labels, int* node_rect, float* npix, int* n_nodes,
                                uint8_t* split, int* cnt, int* base_off, float* top_f, uint8_t* top_b, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QMP_REQUIRE(max_size > 0 && (max_size & (max_size - 1)) == 0, "qmp_quadtree_labels: max_size must be a power of two");
    QMP_REQUIRE(cond >= 0 && cond < 4, "qmp_quadtree_labels: unknown condition");
    QtParams P;
    P.n = n;
    P.m = m;
    P.n_pad = (n + max_size - 1) / max_size * max_size;
    P.m_pad = (m + max_size - 1) / max_size * max_size;
    P.row_cap = P.n_pad < P.m_pad ? P.n_pad : P.m_pad;
    QMP_REQUIRE(n <= P.m_pad, "qmp_quadtree_labels: image taller than its padded width (the reference reads out of bounds)");
    P.L = 0;
    while ((1 << P.L) < max_size) ++P.L;
    P.cond = cond;
    P.thresh = thresh;
    // tiles are always 64 x 64 (a multiple of any max_size <= 64); levels above min(L, 6) go level by level
    const int ts = QT_TS;
    const int LT = P.L < 6 ? P.L : 6;
    dim3 grid(cdiv(P.m_pad, ts), cdiv(P.n_pad, ts));
    const bool tall = P.L > LT;
    qt_tile_kernel<<<grid, 256, 0, st>>>(crit, mask, hir, P, ts, LT, split, cnt, tall ? top_f : nullptr,
                                         tall ? top_b : nullptr);
    QMP_LAUNCH_CHECK("qt_tile_kernel");
    if (tall) {
        const int topcells = (P.n_pad >> LT) * (P.m_pad >> LT);
        float* fin = top_f;
        float* fout = top_f + topcells;
        uint8_t* bin = top_b;
        uint8_t* bout = top_b + topcells;
        for (int lvl = LT + 1; lvl <= P.L; ++lvl) {
            const int cells = (P.n_pad >> lvl) * (P.m_pad >> lvl);
            qt_level_up_kernel<<<cdiv(cells, 256), 256, 0, st>>>(P, lvl, fin, bin, fout, bout, split, cnt);
            QMP_LAUNCH_CHECK("qt_level_up_kernel");
            float* tf = fin; fin = fout; fout = tf;
            uint8_t* tb = bin; bin = bout; bout = tb;
        }
    }
    const int nb = (P.n_pad >> P.L) * (P.m_pad >> P.L);
    qt_base_scan_kernel<<<1, 1024, 0, st>>>(cnt + level_offset(P.n_pad, P.m_pad, P.L), nb, base_off, n_nodes);
    QMP_LAUNCH_CHECK("qt_base_scan_kernel");
    qt_assign_kernel<<<cdiv((long long)n * m, 256), 256, 0, st>>>(P, mask, split, cnt, base_off, labels,
                                                                  (int4*)node_rect, npix);
    QMP_LAUNCH_CHECK("qt_assign_kernel");
    return 0;
}
