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

// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may start while its predecessor in the stream
// (or captured graph) is still running: its prologue -- barrier init, tensor-memory allocation, the bulk copy of the weight
// image -- overlaps the predecessor's tail.  Rules kept by every kernel that is launched this way:
//   * nothing an earlier kernel of the same step wrote is read, and no global memory is written, before pdl_wait()
//     (griddepcontrol.wait: the predecessor grid has completed and its writes are visible);
//   * pdl_launch() (griddepcontrol.launch_dependents) comes AFTER the kernel's own pdl_wait(), so when a dependent starts, the
//     kernel before its predecessor is complete: a prologue may read data that is two or more launches old (weight images,
//     packed parameters, the dropout salt).
// qmp_set_pdl(1) (or QMP_PDL=1) turns the launch attribute on; the default is off (core.cu: the captured training step
// measured 1 % slower with it, a chain of one kernel 2.5 % faster): the same kernels then run in plain stream order.
// The kernels that BUILD those step constants (weight packs and images) call after_producer(): the next launch_pdl() of the
// process is then an ordinary stream-ordered launch, so a prologue never reads an image its direct predecessor wrote.
bool pdl_enabled();
void after_producer();
bool pdl_allowed_now();         // pdl_enabled() and no producer since the last launch_pdl(); clears the producer mark
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_allowed_now() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Exclusive scan of n int32 (n <= 4M).  `blocksums` is caller scratch of >= cdiv(n,1024)+1 ints.
// Writes the grand total to *total (device) when total != nullptr.
int exclusive_scan_i32(const int* in, int* out, int n, int* total, int* blocksums, cudaStream_t st);

}  // namespace qmp
