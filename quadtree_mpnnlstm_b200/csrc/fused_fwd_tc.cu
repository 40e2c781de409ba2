// C-ABI entry of the tcgen05 fused forward (kernel template in fused_fwd_tc.inl, instantiated in fused_fwd_tc_i*.cu).
#include "fused_fwd_tc.inl"

namespace qmp {
extern template int launch_fwd_tc<0, 32>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd_tc<0, 36>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd_tc<4, 32>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd_tc<8, 32>(const FusedFwdArgs&, cudaStream_t);
template <int DAC> int launch_fwd_pw(const FusedFwdArgs&, cudaStream_t);          // fused_fwd_pw.inl (two threads per node)
extern template int launch_fwd_pw<0>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd_pw<4>(const FusedFwdArgs&, cudaStream_t);
extern template int launch_fwd_pw<8>(const FusedFwdArgs&, cudaStream_t);
static int g_paired = 1;
}  // namespace qmp
using namespace qmp;

// 1 (default): groups of 32-wide convs run the paired-warp kernel (two threads per node); 2: also groups with narrow X
// convs; 0: always one thread per node.  Returns the old value.  Both compute the same thing; the switch exists for A/B timing and cross-checks.
QMP_API int qmp_set_fused_paired(int enable) {
    const int old = g_paired;
    g_paired = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    return old;
}

// Same contract as qmp_fused_fwd, on the tensor cores: wa / wb are the weight IMAGES built by qmp_fused_pack_tc
// (kind 0) from the padded packs, [GA, image bytes(cap DA)] and [GB, image bytes(cap DB)].
QMP_API int qmp_fused_fwd_tc(int N, const int* in_ptr, const int* in_src, const float* ea, const float* xa, int lda, int DA,
                             int GA, const void* wa, const float* xb, int ldb, int DB, int GB, int sharedB, const void* wb,
                             int mode, int relu_out, int C, float* out, int ldo, const float* Cprev, const float* params,
                             int norm_h, int norm_c, int norm_o, float eps, float* gates, float* Craw, float* Oout,
                             float* Hout, float* Cout, float* head_in, int ldh, const float* concat, float* logit,
                             float* mstat, float* linv, float drop_p, unsigned long long seed, void* stream) {
    if (N <= 0) return 0;
    QMP_REQUIRE(GB >= 1 && DB >= 1 && DB <= 36 && DA >= 0 && DA <= 8 && C >= 1 && C <= FC, "qmp_fused_fwd_tc: unsupported sizes");
    QMP_REQUIRE(mode == 0 || ((GA == 0 || GA == 4) && (GB == 4 || GB == 8) && C == FC), "qmp_fused_fwd_tc: gate mode needs 4 gates");
    QMP_REQUIRE(sharedB || GB == 8 || mode == 0, "qmp_fused_fwd_tc: own-input gate mode needs 8 convs");
    QMP_REQUIRE(DB % 4 == 0 && ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0 &&
                    (GA == 0 || (DA % 4 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(xa) & 15) == 0)),
                "qmp_fused_fwd_tc: input rows must be 16-byte aligned with a multiple of 4 columns (pad them)");
    QMP_REQUIRE(DB == 32 || DB == 36, "qmp_fused_fwd_tc: segment B rows are 32 or 36 floats wide");
    FusedFwdArgs a{};
    a.N = N; a.ptr = in_ptr; a.nbr = in_src; a.ea = ea; a.xa = xa; a.lda = lda; a.DA = DA; a.GA = GA;
    a.wa = reinterpret_cast<const float*>(wa);
    a.xb = xb; a.ldb = ldb; a.DB = DB; a.GB = GB; a.sharedB = sharedB; a.wb = reinterpret_cast<const float*>(wb);
    a.NC = GA + GB; a.mode = mode;
    a.relu_out = relu_out; a.C = C; a.out = out; a.ldo = ldo; a.Cprev = Cprev; a.params = params; a.norm_h = norm_h;
    a.norm_c = norm_c; a.norm_o = norm_o; a.eps = eps; a.gates = gates; a.Craw = Craw; a.Oout = Oout; a.Hout = Hout;
    a.Cout = Cout; a.head_in = head_in; a.ldh = ldh; a.concat = concat; a.logit = logit; a.mstat = mstat; a.linv = linv;
    a.drop_p = drop_p; a.seed = seed; a.salt = qmp::dropout_salt();
    cudaStream_t st = (cudaStream_t)stream;
    const int dac = (GA == 0) ? 0 : (DA <= 4 ? 4 : 8);
    const int dbc = (DB <= 32) ? 32 : 36;
    // Groups made of 32-wide convs only (encoder layers >= 1, decoder head fc_out2): the paired-warp kernel, two threads
    // per node (measured 127 / 133 us against 148 / 154 us).  With narrow X convs in the group (g_paired == 2 forces it)
    // the second warp of a pair idles through them and the one-thread-per-node kernel is as fast (133 vs 139 us).
    if (dbc == 32 && ((g_paired == 1 && GA == 0) || g_paired == 2)) {
        if (dac == 0) return launch_fwd_pw<0>(a, st);
        if (dac == 4) return launch_fwd_pw<4>(a, st);
        return launch_fwd_pw<8>(a, st);
    }
    if (dac == 0 && dbc == 32) return launch_fwd_tc<0, 32>(a, st);
    if (dac == 0 && dbc == 36) return launch_fwd_tc<0, 36>(a, st);
    if (dac == 4 && dbc == 32) return launch_fwd_tc<4, 32>(a, st);
    if (dac == 8 && dbc == 32) return launch_fwd_tc<8, 32>(a, st);
    qmp::set_error("qmp_fused_fwd_tc: no kernel variant for DA=%d DB=%d", DA, DB);
    return -1;
}
