// Internal interface of the fused kernel-matrix construction (Family 1).
#pragma once
#include "common.cuh"

namespace gpxb {

enum KernelKind { KIND_SE = 0, KIND_M12 = 1, KIND_M32 = 2, KIND_M52 = 3 };

// Layout of the pre-scaled operand Z = x / ell: row stride S doubles, columns [0, Dp)
// hold z (zero padded to a multiple of 4), column Dp holds ||z||^2, the rest are zero.
// S % 8 == 4 makes the DMMA fragment reads bank-conflict free, and a whole 128-row
// tile is one contiguous 128*S*8-byte block (a single 1-D bulk TMA copy).
struct ZLayout {
  int D, Dp, S;
};
inline ZLayout z_layout(int D) {
  ZLayout z;
  z.D = D;
  z.Dp = (D + 3) / 4 * 4;
  int s = z.Dp + 1;
  while (s % 8 != 4) ++s;
  z.S = s;
  return z;
}
constexpr int GRAM_MAX_S = 108;  // 2 tiles * 128 rows * S * 8 B fit in 227 KB of smem; beyond: the K-tiled variant

int prep_z(Ctx* ctx, const double* x, long long ldx, int n, ZLayout zl, double ell, double* Z,
           cudaStream_t s);

struct GramArgs {
  int kind = KIND_SE;
  const double* Z1 = nullptr;
  int n1 = 0;
  const double* Z2 = nullptr;
  int n2 = 0;
  ZLayout zl{};
  double ell = 1.0;
  double diag_add = 0.0;  // added where global row == col
  int sym = 0;            // Z1 and Z2 are the same points: exact d2 = 1e-12 on the diagonal
  long long diag_off = 0;  // store mode: the diagonal is where row + diag_off == col (Z1 is a row range of Z2)
  int lower_only = 0;
  // store mode
  double* out = nullptr;
  long long ldo = 0;
  // contraction mode (out == nullptr): sum_ij (sum_t a_t[i] b_t[j] - Y[i][j]) dK_ij/d ell.  With
  // sym && lower_only the sum over the full symmetric matrix is evaluated on the lower triangle
  // (Y = lower-stored A^-1, b = a); otherwise every (i, j) of the n1 x n2 block counts once.
  const double* Ainv = nullptr;  // Y
  long long ldainv = 0;
  const double* alpha = nullptr;   // a: [t][n1]
  const double* alpha2 = nullptr;  // b: [t][n2] (nullptr: same as a)
  int t = 0;
  double* partials = nullptr;  // one double per CTA
};
long long gram_num_ctas(const GramArgs& a);
int gram_launch(Ctx* ctx, const GramArgs& a, cudaStream_t s);

}  // namespace gpxb
