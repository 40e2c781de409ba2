// Shared helpers for the qmp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <limits.h>
#include <float.h>

#define QMP_API extern "C" __attribute__((visibility("default")))

namespace qmp {

void set_error(const char* fmt, ...);

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#define QMP_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            qmp::set_error(__VA_ARGS__);            \
            return -1;                              \
        }                                           \
    } while (0)

#define QMP_LAUNCH_CHECK(name)                                                      \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            qmp::set_error("%s: %s", name, cudaGetErrorString(e__));                \
            return (int)e__;                                                        \
        }                                                                           \
    } while (0)

#define QMP_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            qmp::set_error("%s: %s", #call, cudaGetErrorString(e__));               \
            return (int)e__;                                                        \
        }                                                                           \
    } while (0)

// Dropout salt (qmp_set_dropout_salt): a DEVICE uint64 the seeded kernels mix into their by-value seed at run time, so that
// a captured CUDA graph draws a new mask every replay (the host bumps the device value; kernel arguments stay frozen).
const unsigned long long* dropout_salt();
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long salted_seed(unsigned long long seed, const unsigned long long* salt) {
    return salt ? seed + 0x9E3779B97F4A7C15ull * __ldg(salt) : seed;
}
#ifdef QMP_NO_SALT      // timing experiment only: what the run-time salt costs
#define QMP_SEED(a) ((a).seed)
#else
#define QMP_SEED(a) qmp::salted_seed((a).seed, (a).salt)
#endif
#endif

// Exclusive scan of n int32 (n <= 4M).  `blocksums` is caller scratch of >= cdiv(n,1024)+1 ints.
// Writes the grand total to *total (device) when total != nullptr.
int exclusive_scan_i32(const int* in, int* out, int n, int* total, int* blocksums, cudaStream_t st);

}  // namespace qmp
