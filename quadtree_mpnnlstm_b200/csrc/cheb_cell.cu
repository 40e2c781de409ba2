// The eight ChebConv / GCNConv stacks of one GConvLSTM step, forward and backward, as ONE entry point each
// (reference model/model.py:394-463 gate pre-activations conv_x_g(X) + conv_h_g(H); model/model.py:60-97 GraphConv;
// PyG 2.2.0 ChebConv K = 3 sym lambda_max = 2 and GCNConv; Seq2Seq's default conv, model/seq2seq.py:203).
//
// On the meshes this conv type runs on (BASELINE configs[0]: 64 x 64 frames, a few hundred to a few thousand nodes) every launch
// is a few microseconds of device time; what bounded a step was the host: one Python call, one allocator round trip and one
// autograd node per launch.  These two functions issue the whole launch sequence of a cell step -- propagation over the in- /
// out-CSR (qmp_spmm), grouped contractions (qmp_gemm), in-place weight-gradient accumulation (qmp_gemm_tn_acc) -- back to back
// on the caller's stream into one workspace laid out by quadtree_mpnnlstm_b200/cheb_cell.py:layout, which also holds the
// Python restatement of the same sequence (_fwd_py / _bwd_py, the cross-check).
//
// Packs: Wb [G, C, K w + 1] = per conv the K weight blocks [C, w] side by side and the bias as last column; layer 0 has one pack
// for the four x stacks (w = F) and one for the four h stacks (w = C), layers >= 1 one pack of eight (w = C).
#include "common.cuh"

QMP_API int qmp_spmm(int N, int width, const int* ptr, const int* nbr, const int* vidx, const float* val, const float* x,
                     int ldx, float alpha, float beta, const float* z, int ldz, float* y, int ldy, void* stream);
QMP_API int qmp_gemm(const float* A, const float* B, const float* bias, float* C, int n, int m, int k, int lda, int ldb,
                     int ldc, long long sA, long long sB, long long sC, long long sBias, int batch, int b_is_kxm,
                     int accumulate, int relu, void* stream);
QMP_API int qmp_gemm_tn_acc(const float* A, const float* B, float* C, int n, int ma, int mb, int lda, int ldb, int ldc,
                            long long sA, long long sB, long long sC, int batch, int b_ones, void* stream);

namespace qmp {

// y[i, c] = a[i, c] (op 0) | y[i, c] -= a[i, c] (op 1) | y[i, c] = a[i, c] + b[i, c] (op 2), rows of width w, leading dimensions
__global__ void cc_rows_kernel(int op, long long n, int w, const float* __restrict__ a, int lda, const float* __restrict__ b,
                               int ldb, float* __restrict__ y, int ldy) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * w) return;
    const long long i = t / w;
    const int c = (int)(t - i * w);
    const float v = a[i * lda + c];
    float* dst = y + i * ldy + c;
    *dst = (op == 0) ? v : (op == 1) ? *dst - v : v + b[i * ldb + c];
}

static inline void cc_rows(int op, int N, int w, const float* a, int lda, const float* b, int ldb, float* y, int ldy, cudaStream_t st) {
    const long long tot = (long long)N * w;
    if (tot == 0) return;
    cc_rows_kernel<<<cdiv(tot, 256), 256, 0, st>>>(op, N, w, a, lda, b, ldb, y, ldy);
}

struct CcGraph { const int* ptr; const int* nbr; const int* vidx; const float* val; };

// gradient of the basis with respect to its input; dT[k] = block k (leading dimension ld, changed in place), out [N, w]
static int cc_basis_bwd(const CcGraph& g, bool cheb, int K, int N, int w, float* const* dT, int ld, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (!cheb) return qmp_spmm(N, w, g.ptr, g.nbr, g.vidx, g.val, dT[0], ld, 1.f, 0.f, nullptr, w, out, w, stream);
    for (int k = K - 1; k > 1; --k) {
        rc = qmp_spmm(N, w, g.ptr, g.nbr, g.vidx, g.val, dT[k], ld, 2.f, 1.f, dT[k - 1], ld, dT[k - 1], ld, stream);
        if (rc) return rc;
        cc_rows(1, N, w, dT[k], ld, nullptr, 0, dT[k - 2], ld, st);
    }
    if (K > 1) return qmp_spmm(N, w, g.ptr, g.nbr, g.vidx, g.val, dT[1], ld, 1.f, 1.f, dT[0], ld, out, w, stream);
    cc_rows(0, N, w, dT[0], ld, nullptr, 0, out, w, st);
    return 0;
}

struct CcLayout {
    long long Tx, Th, in[4], T[4][8], out, total;
    CcLayout(int N, int F, int C, int K, int S, bool cheb) {
        const long long w8 = 8LL * C;
        Tx = 0; Th = (long long)N * K * F;
        long long pos = (long long)N * K * (F + C);
        const int nb = cheb ? K - 1 : 1;
        for (int l = 1; l < S; ++l) {
            in[l] = pos; pos += N * w8;
            for (int k = 0; k < nb; ++k) { T[l][k + (cheb ? 1 : 0)] = pos; pos += N * w8; }
        }
        out = pos;
        if (S > 1) pos += N * w8;
        total = pos;
    }
};

}  // namespace qmp
using namespace qmp;

#define CC_CHECK(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)

// P [N, 4C] from X [N, F], H [N, C]; ws: forward workspace (cheb_cell.py:layout), kept for the backward pass.
QMP_API int qmp_cheb_cell_fwd(int N, int F, int C, int K, int S, int cheb, const int* in_ptr, const int* in_src, const float* val,
                              const float* X, const float* H, const float* pack0, const float* pack1, const float* pack2,
                              const float* pack3, const float* bias0, const float* bias1, const float* bias2, const float* bias3,
                              float* ws, float* P, void* stream) {
    QMP_REQUIRE(S >= 1 && S <= 3 && K >= 1 && K <= 8 && (cheb || K == 1), "qmp_cheb_cell_fwd: 1 <= S <= 3, 1 <= K <= 8, GCN has K = 1");
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float* packs[4] = {pack0, pack1, pack2, pack3};
    const float* biases[4] = {bias0, bias1, bias2, bias3};
    const CcLayout lay(N, F, C, K, S, cheb != 0);
    const int w8 = 8 * C;
    const bool last = S == 1;
    float* cur = last ? P : ws + lay.in[1];
    const int ldc = last ? 4 * C : w8;
    for (int which = 0; which < 2; ++which) {
        const float* inp = which ? H : X;
        const int w = which ? C : F, ld = K * w;
        float* T = ws + (which ? lay.Th : lay.Tx);
        if (cheb) {
            cc_rows(0, N, w, inp, w, nullptr, 0, T, ld, st);
            if (K > 1) CC_CHECK(qmp_spmm(N, w, in_ptr, in_src, nullptr, val, T, ld, 1.f, 0.f, nullptr, w, T + w, ld, stream));
            for (int k = 2; k < K; ++k)
                CC_CHECK(qmp_spmm(N, w, in_ptr, in_src, nullptr, val, T + (k - 1) * w, ld, 2.f, -1.f, T + (k - 2) * w, ld, T + k * w, ld, stream));
        } else {
            CC_CHECK(qmp_spmm(N, w, in_ptr, in_src, nullptr, val, inp, w, 1.f, 0.f, nullptr, w, T, ld, stream));
        }
        float* out = last ? cur : cur + which * 4 * C;
        CC_CHECK(qmp_gemm(T, packs[which], biases[which], out, N, C, ld, ld, ld + 1, ldc, 0, (long long)C * (ld + 1), C, C, 4, 0,
                          (last && which == 1) ? 1 : 0, 0, stream));
    }
    for (int l = 1; l < S; ++l) {
        float* inp = ws + lay.in[l];
        const float* Ts[8];
        if (cheb) {
            Ts[0] = inp;
            for (int k = 1; k < K; ++k) Ts[k] = ws + lay.T[l][k];
            if (K > 1) CC_CHECK(qmp_spmm(N, w8, in_ptr, in_src, nullptr, val, inp, w8, 1.f, 0.f, nullptr, w8, ws + lay.T[l][1], w8, stream));
            for (int k = 2; k < K; ++k)
                CC_CHECK(qmp_spmm(N, w8, in_ptr, in_src, nullptr, val, Ts[k - 1], w8, 2.f, -1.f, Ts[k - 2], w8, ws + lay.T[l][k], w8, stream));
        } else {
            Ts[0] = ws + lay.T[l][0];
            CC_CHECK(qmp_spmm(N, w8, in_ptr, in_src, nullptr, val, inp, w8, 1.f, 0.f, nullptr, w8, ws + lay.T[l][0], w8, stream));
        }
        float* nxt = (l + 1 < S) ? ws + lay.in[l + 1] : ws + lay.out;
        for (int k = 0; k < K; ++k)
            CC_CHECK(qmp_gemm(Ts[k], packs[l + 1] + k * C, k == 0 ? biases[l + 1] : nullptr, nxt, N, C, C, w8, K * C + 1, w8, C,
                              (long long)C * (K * C + 1), C, C, 8, 0, k ? 1 : 0, 0, stream));
    }
    if (!last) cc_rows(2, N, 4 * C, ws + lay.out, w8, ws + lay.out + 4 * C, w8, P, 4 * C, st);
    QMP_LAUNCH_CHECK("qmp_cheb_cell_fwd");
    return 0;
}

// Gradients: accN [as packN] += weight / bias gradients (in place); dX [N, F], dH [N, C] when asked for.  ws: the forward
// workspace; ws2: scratch of cheb_cell.py:scratch_size floats.
QMP_API int qmp_cheb_cell_bwd(int N, int F, int C, int K, int S, int cheb, const int* out_ptr, const int* out_dst, const int* out_kin,
                              const float* val, const float* dP, const float* pack0, const float* pack1, const float* pack2,
                              const float* pack3, float* acc0, float* acc1, float* acc2, float* acc3, const float* ws, float* ws2,
                              int need_dx, int need_dh, float* dX, float* dH, void* stream) {
    QMP_REQUIRE(S >= 1 && S <= 3 && K >= 1 && K <= 8 && (cheb || K == 1), "qmp_cheb_cell_bwd: 1 <= S <= 3, 1 <= K <= 8, GCN has K = 1");
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float* packs[4] = {pack0, pack1, pack2, pack3};
    float* accs[4] = {acc0, acc1, acc2, acc3};
    const CcLayout lay(N, F, C, K, S, cheb != 0);
    const CcGraph g{out_ptr, out_dst, out_kin, val};
    const long long w8 = 8LL * C;
    float* buf[10];
    for (int i = 0; i < 2 + K; ++i) buf[i] = ws2 + i * N * w8;
    float* dTx = ws2 + (2 + K) * N * w8;
    float* dTh = dTx + (long long)N * K * F;
    const float* dOut = dP;
    int ldo = 4 * C;
    if (S > 1) {
        cc_rows(0, N, 4 * C, dP, 4 * C, nullptr, 0, buf[0], (int)w8, st);
        cc_rows(0, N, 4 * C, dP, 4 * C, nullptr, 0, buf[0] + 4 * C, (int)w8, st);
        dOut = buf[0];
        ldo = (int)w8;
    }
    int nxt = 1;
    for (int l = S - 1; l >= 1; --l) {
        const float* Ts[8];
        if (cheb) {
            Ts[0] = ws + lay.in[l];
            for (int k = 1; k < K; ++k) Ts[k] = ws + lay.T[l][k];
        } else {
            Ts[0] = ws + lay.T[l][0];
        }
        float* dT[8];
        for (int k = 0; k < K; ++k) {
            dT[k] = buf[2 + k];
            CC_CHECK(qmp_gemm(dOut, packs[l + 1] + k * C, nullptr, dT[k], N, C, C, (int)w8, K * C + 1, (int)w8, C, (long long)C * (K * C + 1),
                              C, 0, 8, 1, 0, 0, stream));
            const int ones = (k == K - 1) ? 1 : 0;
            CC_CHECK(qmp_gemm_tn_acc(dOut, Ts[k], accs[l + 1] + k * C, N, C, C + ones, (int)w8, (int)w8, K * C + 1, C, C,
                                     (long long)C * (K * C + 1), 8, ones, stream));
        }
        CC_CHECK(cc_basis_bwd(g, cheb != 0, K, N, (int)w8, dT, (int)w8, buf[nxt], stream));
        dOut = buf[nxt];
        nxt = 1 - nxt;
    }
    for (int which = 0; which < 2; ++which) {
        const int w = which ? C : F, ld = K * w;
        const float* T = ws + (which ? lay.Th : lay.Tx);
        const float* dout = (S == 1) ? dOut : dOut + which * 4 * C;
        CC_CHECK(qmp_gemm_tn_acc(dout, T, accs[which], N, C, ld + 1, ldo, ld, ld + 1, C, 0, (long long)C * (ld + 1), 4, 1, stream));
        if (!(which ? need_dh : need_dx)) continue;
        float* dT0 = which ? dTh : dTx;
        CC_CHECK(qmp_gemm(dout, packs[which], nullptr, dT0, N, ld, 4 * C, ldo, ld + 1, ld, 0, 0, 0, 0, 1, 1, 0, 0, stream));
        float* dT[8];
        for (int k = 0; k < K; ++k) dT[k] = dT0 + k * w;
        CC_CHECK(cc_basis_bwd(g, cheb != 0, K, N, w, dT, ld, which ? dH : dX, stream));
    }
    QMP_LAUNCH_CHECK("qmp_cheb_cell_bwd");
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------------
// A chain of up to three ChebConv / GCNConv layers on one input, optional relu after each: the decoder head of a ChebConv / GCNConv
// model (fc_out2(relu(fc_out1(.))), model/seq2seq.py:182-187) and GraphConv stacks (model/model.py:60-97).  Same building blocks
// and pack format as the cell (G = 1: pack_l [1, M_l, K w_l + 1], w_0 = input width, w_(l+1) = M_l).
// ws (saved for the backward pass): per layer T_l [N, K w_l], then out_l [N, M_l] for every layer but the last.
QMP_API int qmp_relu_mask_to(const float* y, const float* g, float* out, long long n, void* stream);

namespace qmp {
struct CsLayout {
    long long T[3], out[3], total;
    int w[3], M[3];
    CsLayout(int N, int K, int L, int w0, int M0, int M1, int M2) {
        M[0] = M0; M[1] = M1; M[2] = M2;
        long long pos = 0;
        int wl = w0;
        for (int l = 0; l < L; ++l) {
            w[l] = wl;
            T[l] = pos; pos += (long long)N * K * wl;
            out[l] = pos;
            if (l + 1 < L) pos += (long long)N * M[l];
            wl = M[l];
        }
        total = pos;
    }
};
}  // namespace qmp

QMP_API int qmp_cheb_stack_fwd(int N, int K, int cheb, int L, int w0, int M0, int M1, int M2, int relu0, int relu1, int relu2,
                               const int* in_ptr, const int* in_src, const float* val, const float* X, const float* pack0,
                               const float* pack1, const float* pack2, const float* bias0, const float* bias1, const float* bias2,
                               float* ws, float* out, void* stream) {
    QMP_REQUIRE(L >= 1 && L <= 3 && K >= 1 && K <= 8 && (cheb || K == 1), "qmp_cheb_stack_fwd: 1 <= L <= 3, 1 <= K <= 8, GCN has K = 1");
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float* packs[3] = {pack0, pack1, pack2};
    const float* biases[3] = {bias0, bias1, bias2};
    const int relus[3] = {relu0, relu1, relu2};
    const CsLayout lay(N, K, L, w0, M0, M1, M2);
    const float* inp = X;
    for (int l = 0; l < L; ++l) {
        const int w = lay.w[l], ld = K * w, M = lay.M[l];
        float* T = ws + lay.T[l];
        if (cheb) {
            cc_rows(0, N, w, inp, w, nullptr, 0, T, ld, st);
            if (K > 1) CC_CHECK(qmp_spmm(N, w, in_ptr, in_src, nullptr, val, T, ld, 1.f, 0.f, nullptr, w, T + w, ld, stream));
            for (int k = 2; k < K; ++k)
                CC_CHECK(qmp_spmm(N, w, in_ptr, in_src, nullptr, val, T + (k - 1) * w, ld, 2.f, -1.f, T + (k - 2) * w, ld, T + k * w, ld, stream));
        } else {
            CC_CHECK(qmp_spmm(N, w, in_ptr, in_src, nullptr, val, inp, w, 1.f, 0.f, nullptr, w, T, ld, stream));
        }
        float* o = (l + 1 < L) ? ws + lay.out[l] : out;
        CC_CHECK(qmp_gemm(T, packs[l], biases[l], o, N, M, ld, ld, ld + 1, M, 0, 0, 0, 0, 1, 0, 0, relus[l], stream));
        inp = o;
    }
    QMP_LAUNCH_CHECK("qmp_cheb_stack_fwd");
    return 0;
}

// accN += weight | bias gradients; dX [N, w0] when asked for.  out_last: the forward result (relu mask of the last layer).
// ws2: scratch of 2 N max(M) + N K max(w) floats.
QMP_API int qmp_cheb_stack_bwd(int N, int K, int cheb, int L, int w0, int M0, int M1, int M2, int relu0, int relu1, int relu2,
                               const int* out_ptr, const int* out_dst, const int* out_kin, const float* val, const float* dOut,
                               const float* out_last, const float* pack0, const float* pack1, const float* pack2, float* acc0,
                               float* acc1, float* acc2, const float* ws, float* ws2, int need_dx, float* dX, void* stream) {
    QMP_REQUIRE(L >= 1 && L <= 3 && K >= 1 && K <= 8 && (cheb || K == 1), "qmp_cheb_stack_bwd: 1 <= L <= 3, 1 <= K <= 8, GCN has K = 1");
    if (N <= 0) return 0;
    const float* packs[3] = {pack0, pack1, pack2};
    float* accs[3] = {acc0, acc1, acc2};
    const int relus[3] = {relu0, relu1, relu2};
    const CsLayout lay(N, K, L, w0, M0, M1, M2);
    const CcGraph g{out_ptr, out_dst, out_kin, val};
    int maxM = 0, maxw = 0;
    for (int l = 0; l < L; ++l) { maxM = lay.M[l] > maxM ? lay.M[l] : maxM; maxw = lay.w[l] > maxw ? lay.w[l] : maxw; }
    maxM = maxM > maxw ? maxM : maxw;
    float* gbuf[2] = {ws2, ws2 + (long long)N * maxM};
    float* dT0 = ws2 + 2LL * N * maxM;
    const float* gcur = dOut;
    int flip = 0;
    for (int l = L - 1; l >= 0; --l) {
        const int w = lay.w[l], ld = K * w, M = lay.M[l];
        const float* T = ws + lay.T[l];
        if (relus[l]) {
            const float* y = (l + 1 < L) ? ws + lay.out[l] : out_last;
            CC_CHECK(qmp_relu_mask_to(y, gcur, gbuf[flip], (long long)N * M, stream));
            gcur = gbuf[flip];
            flip ^= 1;
        }
        CC_CHECK(qmp_gemm_tn_acc(gcur, T, accs[l], N, M, ld + 1, M, ld, ld + 1, 0, 0, 0, 1, 1, stream));
        if (l == 0 && !need_dx) break;
        CC_CHECK(qmp_gemm(gcur, packs[l], nullptr, dT0, N, ld, M, M, ld + 1, ld, 0, 0, 0, 0, 1, 1, 0, 0, stream));
        float* dT[8];
        for (int k = 0; k < K; ++k) dT[k] = dT0 + k * w;
        float* res = (l == 0) ? dX : gbuf[flip];
        CC_CHECK(cc_basis_bwd(g, cheb != 0, K, N, w, dT, ld, res, stream));
        gcur = res;
        flip ^= 1;
    }
    QMP_LAUNCH_CHECK("qmp_cheb_stack_bwd");
    return 0;
}
