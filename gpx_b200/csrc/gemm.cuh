// Internal FP64 DMMA GEMM interface (row-major everywhere).
#pragma once
#include "common.cuh"

namespace gpxb {

// C[M x N] = alpha * opA * opB + beta * C, all row-major with leading dimensions.
//   a_kmajor: A is stored [M][K] (element (i,k) at A[i*lda + k]); else stored [K][M].
//   b_kmajor: B is stored [N][K] (element (k,j) at B[j*ldb + k]); else stored [K][N].
// lower_only: only 128x128 tiles with tile_row >= tile_col are computed and only
//   entries with global row >= col are written (a lower trapezoid; M >= N).
// klo_mode: 0 -> k starts at 0; 1 -> at the tile's first column (B is lower
//   triangular, B[k][j] = 0 for k < j); 2 -> at the tile's first row.
// khi_mode: 0 -> k ends at K; 1 -> at the tile's last row + 1 (A is lower
//   triangular, A[i][k] = 0 for k > i).
// The skipped parts of triangular operands must hold real zeros (partial tiles are read).
struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const double* A = nullptr;
  long long lda = 0;
  bool a_kmajor = true;
  const double* B = nullptr;
  long long ldb = 0;
  bool b_kmajor = true;
  double* C = nullptr;
  long long ldc = 0;
  double alpha = 1.0, beta = 0.0;
  int lower_only = 0;
  int klo_mode = 0;
  int khi_mode = 0;
  int skip_tile0 = 0;  // lower_only: the leading 128 x 128 block is left alone (another kernel owns it)
  // 64 x 64 tiles are used when the launch has at most this many 128 x 128 tiles (0: the library default).  The
  // panel factorisation passes a small number: its GEMMs run beside the one-CTA leaf chain, which needs whole SMs
  // to stay free, and 4x as many CTAs at two per SM would occupy every SM.
  int small_tile_max = 0;
  // batch > 1: `batch` independent products of identical shape in one launch; operand z starts strideX * z
  // doubles after operand 0 (same leading dimensions)
  int batch = 1;
  long long strideA = 0, strideB = 0, strideC = 0;
};

int gemm_f64(Ctx* ctx, const GemmArgs& g, cudaStream_t stream);

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 128;

}  // namespace gpxb
