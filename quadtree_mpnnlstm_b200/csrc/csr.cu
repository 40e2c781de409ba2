// CSR views of an edge list for the message-passing kernels.
//
// The reference hands PyG an ``edge_index[2, E]`` (row 0 = source j, row 1 = target i) and PyG
// scatters per edge.  The kernels here gather instead, so each graph gets, once:
//   in-CSR  (by target): in_ptr[N+1], in_src[E], in_eid[E]   -- edges entering node i, by edge id
//   out-CSR (by source): out_ptr[N+1], out_dst[E], out_kin[E] -- edges leaving node j; out_kin is the
//                         position of that edge in the in-CSR (where per-edge values are stored)
// Rows are ordered by ascending edge id, so every sum over a neighbourhood runs in a fixed order.
#include "common.cuh"

namespace qmp {

__global__ void csr_count_kernel(const int* __restrict__ key, int E, int N, int* __restrict__ cnt, int* __restrict__ bad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int k = key[e];
    if (k < 0 || k >= N) {
        atomicAdd(bad, 1);
        return;
    }
    atomicAdd(&cnt[k], 1);
}

__global__ void csr_fill_kernel(const int* __restrict__ key, int E, int N, int* __restrict__ cursor, int* __restrict__ eid) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int k = key[e];
    if (k < 0 || k >= N) return;
    eid[atomicAdd(&cursor[k], 1)] = e;
}

// insertion sort of each row's edge ids (rows are short; a quadtree node has at most a few hundred)
__global__ void csr_sort_rows_kernel(const int* __restrict__ ptr, int N, int* __restrict__ eid) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int lo = ptr[v], hi = ptr[v + 1];
    for (int a = lo + 1; a < hi; ++a) {
        const int x = eid[a];
        int b = a - 1;
        while (b >= lo && eid[b] > x) {
            eid[b + 1] = eid[b];
            --b;
        }
        eid[b + 1] = x;
    }
}

__global__ void csr_in_finish_kernel(const int* __restrict__ in_eid, int E, const int* __restrict__ valid_total,
                                     const int* __restrict__ src, int* __restrict__ in_src,
                                     int* __restrict__ kin_of_edge) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E || k >= *valid_total) return;  // fewer than E only when endpoints were out of range
    const int e = in_eid[k];
    in_src[k] = src[e];
    kin_of_edge[e] = k;
}

__global__ void csr_out_finish_kernel(const int* __restrict__ out_eid, int E, const int* __restrict__ valid_total,
                                      const int* __restrict__ dst, const int* __restrict__ kin_of_edge,
                                      int* __restrict__ out_dst, int* __restrict__ out_kin) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E || k >= *valid_total) return;
    const int e = out_eid[k];
    out_dst[k] = dst[e];
    out_kin[k] = kin_of_edge[e];
}

__global__ void cast_i64_i32_kernel(const long long* __restrict__ in, long long n, int* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

// rows permuted by an index list: out[k, :] = in[idx[k], :]
__global__ void gather_rows_kernel(const float* __restrict__ in, const int* __restrict__ idx, long long n, int width,
                                   float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * width) return;
    out[t] = in[(size_t)idx[t / width] * width + t % width];
}

static int build_one(const int* key, int E, int N, int* ptr, int* eid, int* tmp, int* blocksums, int* bad, cudaStream_t st) {
    QMP_CUDA(cudaMemsetAsync(tmp, 0, sizeof(int) * (size_t)(N + 1), st));
    if (E > 0) csr_count_kernel<<<cdiv(E, 256), 256, 0, st>>>(key, E, N, tmp, bad);
    int rc = exclusive_scan_i32(tmp, ptr, N + 1, nullptr, blocksums, st);  // tmp[N] = 0 -> ptr[N] = E
    if (rc) return rc;
    QMP_CUDA(cudaMemcpyAsync(tmp, ptr, sizeof(int) * (size_t)(N + 1), cudaMemcpyDeviceToDevice, st));
    if (E > 0) {
        csr_fill_kernel<<<cdiv(E, 256), 256, 0, st>>>(key, E, N, tmp, eid);
        csr_sort_rows_kernel<<<cdiv(N, 128), 128, 0, st>>>(ptr, N, eid);
    }
    QMP_LAUNCH_CHECK("csr build");
    return 0;
}

}  // namespace qmp
using namespace qmp;

// edge_index: int64 [2, E] row-major (the reference layout).  Outputs as described at the top of the file;
// src32/dst32 [E] are int32 copies.  bad (device int, zeroed by the caller) counts out-of-range endpoints.
// Scratch: tmp int32 [N+2], blocksums [(N+1)/1024+2], eid_out int32 [E], kin_of_edge int32 [E].
QMP_API int qmp_csr_from_edge_index(const long long* edge_index, long long E_, int N, int* src32, int* dst32, int* in_ptr,
                                    int* in_src, int* in_eid, int* out_ptr, int* out_dst, int* out_kin, int* bad, int* tmp,
                                    int* blocksums, int* eid_out, int* kin_of_edge, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QMP_REQUIRE(E_ >= 0 && E_ < INT_MAX && N >= 0, "qmp_csr_from_edge_index: bad sizes");
    const int E = (int)E_;
    if (E > 0) {
        cast_i64_i32_kernel<<<cdiv(E, 256), 256, 0, st>>>(edge_index, E, src32);
        cast_i64_i32_kernel<<<cdiv(E, 256), 256, 0, st>>>(edge_index + E, E, dst32);
    }
    int rc = build_one(dst32, E, N, in_ptr, in_eid, tmp, blocksums, bad, st);
    if (rc) return rc;
    rc = build_one(src32, E, N, out_ptr, eid_out, tmp, blocksums, bad, st);
    if (rc) return rc;
    if (E > 0) {
        csr_in_finish_kernel<<<cdiv(E, 256), 256, 0, st>>>(in_eid, E, in_ptr + N, src32, in_src, kin_of_edge);
        csr_out_finish_kernel<<<cdiv(E, 256), 256, 0, st>>>(eid_out, E, out_ptr + N, dst32, kin_of_edge, out_dst, out_kin);
    }
    QMP_LAUNCH_CHECK("qmp_csr_from_edge_index");
    return 0;
}

// out[k, 0:width] = in[idx[k], 0:width]  (edge payload into in-CSR order)
QMP_API int qmp_gather_rows(const float* in, const int* idx, long long n, int width, float* out, void* stream) {
    if (n * width == 0) return 0;
    gather_rows_kernel<<<cdiv(n * width, 256), 256, 0, (cudaStream_t)stream>>>(in, idx, n, width, out);
    QMP_LAUNCH_CHECK("qmp_gather_rows");
    return 0;
}
