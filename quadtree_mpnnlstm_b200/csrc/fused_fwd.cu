// C-ABI entry of the fused forward (kernel template in fused_fwd.inl, instantiated in fused_fwd_i*.cu so the
// heavily unrolled variants compile in parallel).
#include "fused_fwd.inl"

namespace qmp {
extern template int launch_fwd<0, 32>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd<0, 36>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd<4, 32>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd<8, 32>(const FusedFwdArgs&, cudaStream_t);
}  // namespace qmp
using namespace qmp;

// Fused conv-layer-group forward (see the top of this file).  Weight buffers wa / wb are the padded packs of
// fused.cuh, GA / GB convs back to back, padded to DA_cap in {4, 8} and DB_cap in {32, 36} (the caps are
// derived from DA / DB).  GA may be 0 (no segment A).  mode 1 needs GA in {0, 4} and GB in {4, 8}; hidden size 32.
// Saves logit [E, GA+GB] (in-CSR order), mstat / linv [N, GA+GB] for the backward kernels.
QMP_API int qmp_fused_fwd(int N, const int* in_ptr, const int* in_src, const float* ea, const float* xa, int lda, int DA,
                          int GA, const float* wa, const float* xb, int ldb, int DB, int GB, int sharedB, const float* wb,
                          int mode, int relu_out, int C, float* out, int ldo, const float* Cprev, const float* params,
                          int norm_h, int norm_c, int norm_o, float eps, float* gates, float* Craw, float* Oout,
                          float* Hout, float* Cout, float* head_in, int ldh, const float* concat, float* logit,
                          float* mstat, float* linv, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    QMP_REQUIRE(GB >= 1 && DB >= 1 && DB <= 36 && DA >= 0 && DA <= 8 && C >= 1 && C <= FC, "qmp_fused_fwd: unsupported sizes");
    QMP_REQUIRE(mode == 0 || ((GA == 0 || GA == 4) && (GB == 4 || GB == 8) && C == FC), "qmp_fused_fwd: gate mode needs 4 gates");
    QMP_REQUIRE(sharedB || GB == 8 || mode == 0, "qmp_fused_fwd: own-input gate mode needs 8 convs");
    FusedFwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.xa = xa; a.lda = lda; a.DA = DA; a.GA = GA; a.wa = wa;
    a.xb = xb; a.ldb = ldb; a.DB = DB; a.GB = GB; a.sharedB = sharedB; a.wb = wb; a.NC = GA + GB; a.mode = mode;
    a.relu_out = relu_out; a.C = C; a.out = out; a.ldo = ldo; a.Cprev = Cprev; a.params = params; a.norm_h = norm_h;
    a.norm_c = norm_c; a.norm_o = norm_o; a.eps = eps; a.gates = gates; a.Craw = Craw; a.Oout = Oout; a.Hout = Hout;
    a.Cout = Cout; a.head_in = head_in; a.ldh = ldh; a.concat = concat; a.logit = logit; a.mstat = mstat; a.linv = linv;
    a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    cudaStream_t st = (cudaStream_t)stream;
    const int dac = (GA == 0) ? 0 : (DA <= 4 ? 4 : 8);
    const int dbc = (DB <= 32) ? 32 : 36;
    if (dac == 0 && dbc == 32) return launch_fwd<0, 32>(a, st);
    if (dac == 0 && dbc == 36) return launch_fwd<0, 36>(a, st);
    if (dac == 4 && dbc == 32) return launch_fwd<4, 32>(a, st);
    if (dac == 8 && dbc == 32) return launch_fwd<8, 32>(a, st);
    qmp::set_error("qmp_fused_fwd: no kernel variant for DA=%d DB=%d", DA, DB);
    return -1;
}
