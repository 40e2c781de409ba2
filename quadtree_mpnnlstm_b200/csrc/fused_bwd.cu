// C-ABI entries of the fused backward (kernels in fused_bwd.inl, instantiated in fused_bwd_i*.cu).
#include "fused_bwd.inl"

namespace qmp {
extern template int launch_bwd<0, 32>(const FusedBwdArgs&, int, cudaStream_t);
extern template int launch_bwd<0, 36>(const FusedBwdArgs&, int, cudaStream_t);
extern template int launch_bwd<4, 32>(const FusedBwdArgs&, int, cudaStream_t);
extern template int launch_bwd<8, 32>(const FusedBwdArgs&, int, cudaStream_t);

static int dispatch_bwd(const FusedBwdArgs& a, int which, cudaStream_t st) {
    const int dac = (a.GA == 0) ? 0 : (a.DA <= 4 ? 4 : 8);
    const int dbc = (a.DB <= 32) ? 32 : 36;
    if (dac == 0 && dbc == 32) return launch_bwd<0, 32>(a, which, st);
    if (dac == 0 && dbc == 36) return launch_bwd<0, 36>(a, which, st);
    if (dac == 4 && dbc == 32) return launch_bwd<4, 32>(a, which, st);
    if (dac == 8 && dbc == 32) return launch_bwd<8, 32>(a, which, st);
    set_error("qmp_fused_bwd: no kernel variant for DA=%d DB=%d", a.DA, a.DB);
    return -1;
}
}  // namespace qmp
using namespace qmp;

// Target side of the fused backward.  dP [N, lddp]: gate mode -> [N, 4*32] from qmp_lstm_gates_bwd (conv c feeds
// gate c (segment A) / (c - GA) % 4 (segment B)); plain mode -> the upstream gradient [N, NC*C].  wa / wb: backward
// weight packs (fused_bwd.inl).  Writes ds [E, NC], Zs* / dUs* [N, G, cap+4] (operands of the weight-gradient
// reductions) and, when asked, the self part of dxa [N, lda] / dxb [N, ldb].
QMP_API int qmp_fused_bwd_target(int N, const int* in_ptr, const int* in_src, const float* ea, const float* xa, int lda,
                                 int DA, int GA, const float* wa, const float* xb, int ldb, int DB, int GB, int sharedB,
                                 const float* wb, int mode, int C, const float* dP, int lddp, const float* logit,
                                 const float* mstat, const float* linv, float* ds, float* ZsA, float* dUsA, float* ZsB,
                                 float* dUsB, float* dxa, float* dxb, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    FusedBwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.xa = xa; a.lda = lda; a.DA = DA; a.GA = GA; a.wa = wa;
    a.xb = xb; a.ldb = ldb; a.DB = DB; a.GB = GB; a.sharedB = sharedB; a.wb = wb; a.NC = GA + GB; a.mode = mode; a.C = C;
    a.dP = dP; a.lddp = lddp; a.logit = logit; a.mstat = mstat; a.linv = linv; a.ds = ds; a.ZsA = ZsA; a.dUsA = dUsA;
    a.ZsB = ZsB; a.dUsB = dUsB; a.dxa = dxa; a.dxb = dxb; a.need_dxa = dxa != nullptr; a.need_dxb = dxb != nullptr;
    a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    return dispatch_bwd(a, 0, (cudaStream_t)stream);
}

// Source side: adds to dxa / dxb (which hold the self part) the contributions through the edges leaving each node.
QMP_API int qmp_fused_bwd_source(int N, const int* out_ptr, const int* out_dst, const int* out_kin, const float* xa, int lda,
                                 int DA, int GA, const float* wa, const float* xb, int ldb, int DB, int GB, int sharedB,
                                 const float* wb, int mode, int C, const float* dP, int lddp, const float* logit,
                                 const float* mstat, const float* linv, const float* ds, float* dxa, float* dxb,
                                 float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0 || (dxa == nullptr && dxb == nullptr)) return 0;
    FusedBwdArgs a{};
    a.N = N; a.ptr = out_ptr; a.nbr = out_dst; a.kin = out_kin; a.xa = xa; a.lda = lda; a.DA = DA; a.GA = GA; a.wa = wa;
    a.xb = xb; a.ldb = ldb; a.DB = DB; a.GB = GB; a.sharedB = sharedB; a.wb = wb; a.NC = GA + GB; a.mode = mode; a.C = C;
    a.dP = dP; a.lddp = lddp; a.logit = logit; a.mstat = mstat; a.linv = linv; a.ds = const_cast<float*>(ds);
    a.dxa = dxa; a.dxb = dxb; a.need_dxa = dxa != nullptr; a.need_dxb = dxb != nullptr; a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    return dispatch_bwd(a, 1, (cudaStream_t)stream);
}
