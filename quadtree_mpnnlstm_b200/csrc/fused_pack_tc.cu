// Weight images for the tcgen05 fused kernels: from the padded fp32 pack of a conv group (layout: fused.cuh,
// W1 | b1 | W2 | W3 | b3 per conv) to the shared-memory images the kernels stream (fused_tc.cuh): every matrix as a
// K-major B operand (rows = outputs of the contraction, columns = its reduction index), split into hi / lo TF32
// parts, in the canonical no-swizzle layout, plus the plain fp32 biases.  One tiny launch per group and optimizer
// step (the weights are constant over the timesteps of a forward / backward pass).
#include "common.cuh"
#include "fused_tc.cuh"

namespace qmp {

__device__ __forceinline__ void img_put(uint8_t* img, int off_hi, int off_lo, int n, int k, int K, float v) {
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    *reinterpret_cast<float*>(img + off_hi + img_off(n, k, K)) = hi;
    *reinterpret_cast<float*>(img + off_lo + img_off(n, k, K)) = lo;
}

// which = 0: forward image (TcFwdLayout); 1: backward target side (TcBwdTLayout); 2: backward source side (TcBwdSLayout)
__global__ void __launch_bounds__(256) fused_pack_tc_kernel(const float* __restrict__ pack, int G, int DC, int which,
                                                            uint8_t* __restrict__ out) {
    const int g = blockIdx.x;
    const int o1 = (DC + 2) * DC, o2 = o1 + DC + 4, o3 = o2 + FC * (DC + 4), o4 = o3 + FC * DC, total = o4 + FC;
    const float* W1 = pack + (size_t)g * total;
    const float* b1 = W1 + o1;
    const float* W2 = W1 + o2;
    const float* W3 = W1 + o3;
    const float* b3 = W1 + o4;
    if (which == 0) {
        const TcFwdLayout L(DC);
        uint8_t* img = out + (size_t)g * L.BYTES;
        for (int idx = threadIdx.x; idx < L.N1 * L.K1; idx += 256) {          // W1: rows u (DC) | w (2), columns x
            const int n = idx / L.K1, k = idx % L.K1;
            img_put(img, L.W1H, L.W1L, n, k, L.K1, (n < DC + 2 && k < DC) ? W1[n * DC + k] : 0.f);
        }
        for (int idx = threadIdx.x; idx < FC * L.K1; idx += 256) {            // W3 (skip): rows outputs, columns x
            const int n = idx / L.K1, k = idx % L.K1;
            img_put(img, L.W3H, L.W3L, n, k, L.K1, (k < DC) ? W3[n * DC + k] : 0.f);
        }
        for (int idx = threadIdx.x; idx < FC * L.K2; idx += 256) {            // W2: rows outputs, columns z | ze | zs
            const int n = idx / L.K2, k = idx % L.K2;
            img_put(img, L.W2H, L.W2L, n, k, L.K2, (k < DC + 4) ? W2[n * (DC + 4) + k] : 0.f);
        }
        for (int idx = threadIdx.x; idx < 48; idx += 256) reinterpret_cast<float*>(img + L.B1)[idx] = (idx < DC + 4) ? b1[idx] : 0.f;
        for (int idx = threadIdx.x; idx < FC; idx += 256) reinterpret_cast<float*>(img + L.B3)[idx] = b3[idx];
    } else if (which == 1) {
        const TcBwdTLayout L(DC);
        uint8_t* img = out + (size_t)g * L.BYTES;
        for (int idx = threadIdx.x; idx < L.N2 * FC; idx += 256) {            // rows = z | ze | zs index, columns = outputs o
            const int n = idx / FC, o = idx % FC;
            img_put(img, L.W2TH, L.W2TL, n, o, FC, (n < DC + 4) ? W2[o * (DC + 4) + n] : 0.f);
        }
        for (int idx = threadIdx.x; idx < L.N1P * FC; idx += 256) {           // rows = x index, columns = outputs o
            const int n = idx / FC, o = idx % FC;
            img_put(img, L.W3TH, L.W3TL, n, o, FC, (n < DC) ? W3[o * DC + n] : 0.f);
        }
        for (int idx = threadIdx.x; idx < L.N1P * L.K2; idx += 256) {         // rows = x index, columns = [du | dw] index r
            const int n = idx / L.K2, r = idx % L.K2;
            img_put(img, L.W1TH, L.W1TL, n, r, L.K2, (n < DC && r < DC + 2) ? W1[r * DC + n] : 0.f);
        }
        for (int idx = threadIdx.x; idx < DC * L.K1; idx += 256) {            // plain fp32: row = x index m, column = u index k
            const int m = idx / L.K1, k = idx % L.K1;
            reinterpret_cast<float*>(img + L.W1P)[idx] = (k < DC) ? W1[k * DC + m] : 0.f;
        }
        for (int idx = threadIdx.x; idx < L.K1; idx += 256) reinterpret_cast<float*>(img + L.B1P)[idx] = (idx < DC) ? b1[idx] : 0.f;
    } else {
        const TcBwdSLayout L(DC);
        uint8_t* img = out + (size_t)g * L.BYTES;
        for (int idx = threadIdx.x; idx < L.N1P * L.KS; idx += 256) {         // rows = x index k, columns = av | bv | sum ds
            const int n = idx / L.KS, col = idx % L.KS;
            float v = 0.f;
            if (n < DC) {
                if (col < FC) v = W2[col * (DC + 4) + n];
                else if (col < FC + DC) v = W1[n * DC + (col - FC)];
                else if (col == FC + L.K1) v = b1[n];
            }
            img_put(img, L.BSH, L.BSL, n, col, L.KS, v);
        }
    }
}

}  // namespace qmp
using namespace qmp;

// Bytes of one conv's image (which = 0: forward).  Returns -1 for an unknown kind.
QMP_API long long qmp_fused_tc_image_bytes(int DC, int which) {
    if (which == 0) return TcFwdLayout(DC).BYTES;
    if (which == 1) return TcBwdTLayout(DC).BYTES;
    if (which == 2) return TcBwdSLayout(DC).BYTES;
    return -1;
}

// pack [G, TOTAL(DC)] (fused.cuh layout, DC in {4, 8, 32, 36}) -> out [G, image bytes]
QMP_API int qmp_fused_pack_tc(const float* pack, int G, int DC, int which, void* out, void* stream) {
    if (G <= 0) return 0;
    QMP_REQUIRE(DC == 4 || DC == 8 || DC == 32 || DC == 36, "qmp_fused_pack_tc: DC must be 4, 8, 32 or 36");
    QMP_REQUIRE(which >= 0 && which <= 2, "qmp_fused_pack_tc: unknown image kind %d", which);
    fused_pack_tc_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(pack, G, DC, which, (uint8_t*)out);
    QMP_LAUNCH_CHECK("fused_pack_tc_kernel");
    qmp::after_producer();
    return 0;
}
