// C-ABI entries of the tcgen05 fused backward (kernels in fused_bwd_tc.inl, instantiated in fused_bwd_tc_k*_i*.cu).
#include "fused_bwd_tc.inl"

namespace qmp {
#define QMP_DECL(K)                                                                     \
    extern template int launch_bwd_tc<0, 32, K>(const FusedBwdArgs&, cudaStream_t);     \
    extern template int launch_bwd_tc<0, 36, K>(const FusedBwdArgs&, cudaStream_t);     \
    extern template int launch_bwd_tc<4, 32, K>(const FusedBwdArgs&, cudaStream_t);     \
    extern template int launch_bwd_tc<8, 32, K>(const FusedBwdArgs&, cudaStream_t);
QMP_DECL(1)
QMP_DECL(2)
#undef QMP_DECL

template <int KIND>
static int dispatch_bwd_tc(const FusedBwdArgs& a, cudaStream_t st) {
    const int dac = (a.GA == 0) ? 0 : (a.DA <= 4 ? 4 : 8);
    const int dbc = (a.DB <= 32) ? 32 : 36;
    if (dac == 0 && dbc == 32) return launch_bwd_tc<0, 32, KIND>(a, st);
    if (dac == 0 && dbc == 36) return launch_bwd_tc<0, 36, KIND>(a, st);
    if (dac == 4 && dbc == 32) return launch_bwd_tc<4, 32, KIND>(a, st);
    if (dac == 8 && dbc == 32) return launch_bwd_tc<8, 32, KIND>(a, st);
    set_error("qmp_fused_bwd_tc: no kernel variant for DA=%d DB=%d", a.DA, a.DB);
    return -1;
}
}  // namespace qmp
using namespace qmp;

// Same contract as qmp_fused_bwd_target, on the tensor cores: wa / wb are weight images of kind 1 (qmp_fused_pack_tc).
QMP_API int qmp_fused_bwd_target_tc(int N, const int* in_ptr, const int* in_src, const float* ea, const float* xa, int lda,
                                    int DA, int GA, const void* wa, const float* xb, int ldb, int DB, int GB, int sharedB,
                                    const void* wb, int mode, int C, const float* dP, int lddp, const float* logit,
                                    const float* mstat, const float* linv, float* ds, float* ZsA, float* dUsA, float* ZsB,
                                    float* dUsB, float* dxa, float* dxb, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    QMP_REQUIRE(GB >= 1 && DB >= 1 && DB <= 36 && DA >= 0 && DA <= 8 && C >= 1 && C <= FC, "qmp_fused_bwd_target_tc: unsupported sizes");
    QMP_REQUIRE((DB == 32 || DB == 36) && ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0 &&
                    (GA == 0 || (DA % 4 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(xa) & 15) == 0)) &&
                    (reinterpret_cast<uintptr_t>(dP) & 15) == 0 && (mode == 0 || lddp % 4 == 0),
                "qmp_fused_bwd_*_tc: rows must be 16-byte aligned with a multiple of 4 columns (pad them)");
    FusedBwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.xa = xa; a.lda = lda; a.DA = DA; a.GA = GA;
    a.wa = reinterpret_cast<const float*>(wa);
    a.xb = xb; a.ldb = ldb; a.DB = DB; a.GB = GB; a.sharedB = sharedB; a.wb = reinterpret_cast<const float*>(wb);
    a.NC = GA + GB; a.mode = mode; a.C = C;
    a.dP = dP; a.lddp = lddp; a.logit = logit; a.mstat = mstat; a.linv = linv; a.ds = ds; a.ZsA = ZsA; a.dUsA = dUsA;
    a.ZsB = ZsB; a.dUsB = dUsB; a.dxa = dxa; a.dxb = dxb; a.need_dxa = dxa != nullptr; a.need_dxb = dxb != nullptr;
    a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    return dispatch_bwd_tc<1>(a, (cudaStream_t)stream);
}

// Target AND source side in one launch (fused_bwd_tc.inl, one-pass mode): same arguments as qmp_fused_bwd_target_tc, but dxa / dxb
// (those that are not null; [N, lda] / [N, ldb], zeroed here) receive the complete input gradient -- every in-edge's contribution
// to its source row by 16-byte vector reductions -- so qmp_fused_bwd_source_tc, the out-CSR and the kind-2 images are not
// needed.  ds is scratch ([E, GA + GB]).
QMP_API int qmp_fused_bwd_onepass_tc(int N, const int* in_ptr, const int* in_src, const float* ea, const float* xa, int lda,
                                     int DA, int GA, const void* wa, const float* xb, int ldb, int DB, int GB, int sharedB,
                                     const void* wb, int mode, int C, const float* dP, int lddp, const float* logit,
                                     const float* mstat, const float* linv, float* ds, float* ZsA, float* dUsA, float* ZsB,
                                     float* dUsB, float* dxa, float* dxb, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    QMP_REQUIRE(GB >= 1 && DB >= 1 && DB <= 36 && DA >= 0 && DA <= 8 && C >= 1 && C <= FC, "qmp_fused_bwd_onepass_tc: unsupported sizes");
    QMP_REQUIRE((DB == 32 || DB == 36) && ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0 &&
                    (GA == 0 || (DA % 4 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(xa) & 15) == 0)) &&
                    (reinterpret_cast<uintptr_t>(dP) & 15) == 0 && (mode == 0 || lddp % 4 == 0),
                "qmp_fused_bwd_*_tc: rows must be 16-byte aligned with a multiple of 4 columns (pad them)");
    QMP_REQUIRE((dxa == nullptr || (reinterpret_cast<uintptr_t>(dxa) & 15) == 0) && (dxb == nullptr || (reinterpret_cast<uintptr_t>(dxb) & 15) == 0),
                "qmp_fused_bwd_onepass_tc: gradient rows must be 16-byte aligned (vector reductions)");
    FusedBwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.xa = xa; a.lda = lda; a.DA = DA; a.GA = GA;
    a.wa = reinterpret_cast<const float*>(wa);
    a.xb = xb; a.ldb = ldb; a.DB = DB; a.GB = GB; a.sharedB = sharedB; a.wb = reinterpret_cast<const float*>(wb);
    a.NC = GA + GB; a.mode = mode; a.C = C;
    a.dP = dP; a.lddp = lddp; a.logit = logit; a.mstat = mstat; a.linv = linv; a.ds = ds; a.ZsA = ZsA; a.dUsA = dUsA;
    a.ZsB = ZsB; a.dUsB = dUsB; a.dxa = dxa; a.dxb = dxb; a.need_dxa = dxa != nullptr; a.need_dxb = dxb != nullptr;
    a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt(); a.onepass = 1;
    if (dxa) QMP_CUDA(cudaMemsetAsync(dxa, 0, (size_t)N * lda * sizeof(float), (cudaStream_t)stream));
    if (dxb) QMP_CUDA(cudaMemsetAsync(dxb, 0, (size_t)N * ldb * sizeof(float), (cudaStream_t)stream));
    return dispatch_bwd_tc<1>(a, (cudaStream_t)stream);
}

// Same contract as qmp_fused_bwd_source, on the tensor cores: wa / wb are weight images of kind 2.
QMP_API int qmp_fused_bwd_source_tc(int N, const int* out_ptr, const int* out_dst, const int* out_kin, const float* xa, int lda,
                                    int DA, int GA, const void* wa, const float* xb, int ldb, int DB, int GB, int sharedB,
                                    const void* wb, int mode, int C, const float* dP, int lddp, const float* logit,
                                    const float* mstat, const float* linv, const float* ds, float* dxa, float* dxb,
                                    float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0 || (dxa == nullptr && dxb == nullptr)) return 0;
    QMP_REQUIRE((DB == 32 || DB == 36) && ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0 &&
                    (GA == 0 || (DA % 4 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(xa) & 15) == 0)) &&
                    (reinterpret_cast<uintptr_t>(dP) & 15) == 0 && (mode == 0 || lddp % 4 == 0),
                "qmp_fused_bwd_*_tc: rows must be 16-byte aligned with a multiple of 4 columns (pad them)");
    FusedBwdArgs a{};
    a.N = N; a.ptr = out_ptr; a.nbr = out_dst; a.kin = out_kin; a.xa = xa; a.lda = lda; a.DA = DA; a.GA = GA;
    a.wa = reinterpret_cast<const float*>(wa);
    a.xb = xb; a.ldb = ldb; a.DB = DB; a.GB = GB; a.sharedB = sharedB; a.wb = reinterpret_cast<const float*>(wb);
    a.NC = GA + GB; a.mode = mode; a.C = C;
    a.dP = dP; a.lddp = lddp; a.logit = logit; a.mstat = mstat; a.linv = linv; a.ds = const_cast<float*>(ds);
    a.dxa = dxa; a.dxb = dxb; a.need_dxa = dxa != nullptr; a.need_dxb = dxb != nullptr; a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    return dispatch_bwd_tc<2>(a, (cudaStream_t)stream);
}
