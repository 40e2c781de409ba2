// Error string, version, and the int32 exclusive scan shared by the graph kernels.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace qmp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static const unsigned long long* g_salt = nullptr;
const unsigned long long* dropout_salt() { return g_salt; }

static int g_pdl = -1;      // -1: not decided yet (QMP_PDL; default OFF: measured below)
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("QMP_PDL");
        g_pdl = e ? (atoi(e) != 0) : 0;
    }
    return g_pdl != 0;
}
static bool g_producer = true;
void after_producer() { g_producer = true; }
bool pdl_allowed_now() {
    const bool ok = pdl_enabled() && !g_producer;
    g_producer = false;
    return ok;
}

// ---- scan: 1024 items per block (256 threads x 4), then a single block scans the block sums.
__global__ void __launch_bounds__(256) scan_block_kernel(const int* __restrict__ in, int* __restrict__ out, int n,
                                                         int* __restrict__ blocksums) {
    __shared__ int warp_tot[8];
    const int base = blockIdx.x * 1024 + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (base + k < n) ? in[base + k] : 0;
    int tsum = v[0] + v[1] + v[2] + v[3];
    // inclusive warp scan of per-thread sums
    int incl = tsum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    int excl = woff + incl - tsum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 255) blocksums[blockIdx.x] = woff + incl;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int* __restrict__ blocksums, int nb, int* __restrict__ total) {
    // each thread owns a contiguous chunk; Hillis-Steele over the 1024 chunk totals
    __shared__ int part[1024];
    const int per = (nb + 1023) / 1024;
    const int lo = threadIdx.x * per;
    int s = 0;
    for (int k = 0; k < per; ++k)
        if (lo + k < nb) s += blocksums[lo + k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        int t = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;  // exclusive prefix of my chunk
    for (int k = 0; k < per; ++k)
        if (lo + k < nb) {
            int t = blocksums[lo + k];
            blocksums[lo + k] = run;
            run += t;
        }
    if (threadIdx.x == 1023 && total) *total = part[1023];
}

__global__ void __launch_bounds__(256) scan_add_kernel(int* __restrict__ out, int n, const int* __restrict__ blocksums) {
    const int base = blockIdx.x * 1024 + threadIdx.x * 4;
    const int add = blocksums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (base + k < n) out[base + k] += add;
}

int exclusive_scan_i32(const int* in, int* out, int n, int* total, int* blocksums, cudaStream_t st) {
    if (n <= 0) {
        if (total) QMP_CUDA(cudaMemsetAsync(total, 0, sizeof(int), st));
        return 0;
    }
    const int nb = cdiv(n, 1024);
    QMP_REQUIRE(nb <= 4096, "exclusive_scan_i32: n=%d too large", n);
    scan_block_kernel<<<nb, 256, 0, st>>>(in, out, n, blocksums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(blocksums, nb, total);
    if (nb > 1) scan_add_kernel<<<nb, 256, 0, st>>>(out, n, blocksums);
    QMP_LAUNCH_CHECK("exclusive_scan_i32");
    return 0;
}

}  // namespace qmp

QMP_API const char* qmp_last_error(void) { return qmp::g_err; }
QMP_API int qmp_version(void) { return 100; }

// Exposed for tests: out[i] = sum_{k<i} in[k]; *total = sum.  scratch >= n/1024 + 2 ints.
QMP_API int qmp_exclusive_scan_i32(const int* in, int* out, int n, int* total, int* scratch, void* stream) {
    return qmp::exclusive_scan_i32(in, out, n, total, scratch, (cudaStream_t)stream);
}

// Dropout salt: a device uint64 (or NULL = none, the default) that every seeded kernel launched AFTER this call mixes into its
// by-value seed: effective seed = seed + 0x9E3779B97F4A7C15 * (*salt), read on the device when the kernel runs.  Forward and
// backward kernels of one step see the same value as long as the caller changes *salt only between steps.  This is what lets
// a captured CUDA graph (whose kernel arguments are frozen) resample its dropout masks every replay (the reference resamples
// per call: torch.nn.functional.dropout in PyG TransformerConv.message, nn.Dropout in model/seq2seq.py:169).  Process-wide
// host state, read at launch time; returns 0.
// Programmatic dependent launch of the hot kernels (common.cuh, launch_pdl): on = 1 (the prologue of a kernel overlaps the tail
// of its predecessor in the stream / captured graph), off = 0 (default: plain stream order).  Measured on B200
// (profiles/r03_pdl.txt): a captured chain of decoder-cell forward launches runs 51.2 -> 49.9 us per launch with it, the
// whole 10 + 90-frame training step 40.0 -> 40.5 ms (two A/B runs, either order): the default stays off.
// Process-wide host state, read at launch time; returns the previous setting.
QMP_API int qmp_set_pdl(int on) {
    const int prev = qmp::pdl_enabled() ? 1 : 0;
    qmp::g_pdl = on ? 1 : 0;
    return prev;
}

QMP_API int qmp_set_dropout_salt(const unsigned long long* salt) {
    qmp::g_salt = salt;
    return 0;
}
