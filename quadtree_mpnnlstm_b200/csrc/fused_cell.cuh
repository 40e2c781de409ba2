// Weight image of the decoder-cell kernel (fused_cell_fwd.cu): the eight convs of one GConvLSTM step -- four on the
// 4-wide input X, four on the 32-wide hidden state H, one pair per gate (model/model.py:394-463) -- packed so that the
// dense contractions of ALL gates run as a few wide tcgen05 instructions instead of one narrow chain per conv:
//
//   W1   [160 x 32]  logit projections of the four H convs side by side: block g = rows 40g .. 40g+39 =
//                     u rows (32) | edge-attribute rows w (2) | zero (6)                       -> U  = h W1^T    (N = 160)
//   W3   [128 x 40]  skip projections of all gates, columns h (32) | x (4) | zero (4)          -> P  = [h|x] W3^T (N = 128)
//   W2_g [ 32 x 48]  value projections of gate g, columns z_h (32) | ze0 ze1 zs (3) | zero (5) | z_x (4) | ze0 ze1 zs 0
//                                                                                             -> P_g += [z_h|..|z_x|..] W2_g^T
// every matrix as a K-major B operand (no swizzle, fused_tc.cuh img_off), split into TF32 hi / lo parts, followed by the
// fp32 vectors the SIMT phases read.  Built by qmp_fused_pack_cell from the two padded packs (fused.cuh layout).
#pragma once
#include "fused_tc.cuh"

namespace qmp {

struct CellLayout {
    static constexpr int UB = 40;                         // columns per H conv in U
    static constexpr int KU = 32, NU = 4 * UB;            // U  contraction
    static constexpr int KS = 40, NS = 4 * FC;            // skip contraction
    static constexpr int KZ = 48, NZ = FC;                // value contraction of one gate
    static constexpr int W1H = 0, W1L = W1H + NU * KU * 4;
    static constexpr int W3H = W1L + NU * KU * 4, W3L = W3H + NS * KS * 4;
    static constexpr int W2H = W3L + NS * KS * 4;         // gate g: hi at W2H + g * W2G, lo NZ * KZ * 4 bytes further
    static constexpr int W2G = 2 * NZ * KZ * 4;
    static constexpr int MMA_BYTES = W2H + 4 * W2G;
    static constexpr int B1H = MMA_BYTES;                 // [4][UB]  logit biases of the H convs (u | w | 0)
    static constexpr int W1X = B1H + 4 * UB * 4;          // [4][6][4] logit weights of the X convs (u rows 0..3, w rows 4, 5)
    static constexpr int B1X = W1X + 4 * 24 * 4;          // [4][8]
    static constexpr int B3S = B1X + 4 * 8 * 4;           // [4][32]  skip biases, X conv + H conv of each gate
    static constexpr int BYTES = B3S + 4 * FC * 4;
};
static_assert(CellLayout::BYTES % 16 == 0, "bulk copies move 16-byte units");

}  // namespace qmp
