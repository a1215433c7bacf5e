// Blocked FP64 Cholesky, triangular solves, log-det and SPD inverse for sm_100a.
//
// Replaces jsp.linalg.cholesky / solve_triangular / log(diag) in the reference
// (gpx/models/_gpr.py:457-458, 759-763, 860-861; _sgpr.py:459-462, 684-691) and supplies
// A^-1 for the analytic LML gradient (what value_and_grad yields,
// gpx/optimizers/scipy_optimize.py:72-78).
//
// Structure: two-level right-looking lower Cholesky on row-major storage; every O(n^3)
// flop is a DMMA GEMM (gemm.cu) -- SYRK trailing updates and TRSMs that multiply
// by explicitly inverted 128x128 diagonal blocks.  The 128x128 leaves are
// factorised AND inverted by one CTA in shared memory.  Solves and the triangular
// inverse recurse on halves so that their flops are large GEMMs too.
#include "chol.cuh"
#include "gemm.cuh"
#include "mma.cuh"

#include <algorithm>

namespace gpxb {

using std::max;
using std::min;

namespace {

constexpr int LB = CHOL_LEAF;  // 128
constexpr int LS = LB + 1;     // padded smem stride (odd -> conflict-free columns)
constexpr int LEAF_NT = 256;
constexpr int PB = 8;          // panel width inside the leaf
constexpr int PTS = LB + 8;    // row stride of the transposed panel buffer
constexpr int LEAF_SMEM = (LB * LS + PB * PTS + 128) * (int)sizeof(double);

// One CTA: factor the n x n (n <= 128) lower block in place, then write its inverse
// (full 128x128 block, zero upper / zero padding) to inv.
//
// Blocked inside shared memory with 8-column panels: the panel is factorised with one thread per
// row holding its 8 entries in registers (two barriers per column, almost no work between them);
// the rank-8 trailing update and the inverse's block column use 4x4 / 1x8 register tiles.
__global__ void __launch_bounds__(LEAF_NT, 1)
potrf_leaf_kernel(double* __restrict__ A, long long lda, int n, double* __restrict__ inv) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                 // [LB][LS]
  double* PT = sm + LB * LS;      // [PB][PTS]  transposed panel: PT[c][row]
  double* lsh = PT + PB * PTS;    // [8] multipliers of the diagonal-block rows, [8] pivot
  double* xbb = lsh + 16;         // [8][8] inverse of a diagonal 8x8 block
  double* rdg = S + LB;           // rdg[j * LS] = 1 / L[j][j], kept in the padding column of S
  const int tid = threadIdx.x;

  for (int idx = tid; idx < LB * LB; idx += LEAF_NT) {
    const int i = idx >> 7, k = idx & (LB - 1);
    double v = 0.0;
    if (i < n && k <= i) v = A[(long long)i * lda + k];
    S[i * LS + k] = v;
  }
  __syncthreads();

  // ---------------- factorisation ----------------
  for (int jb = 0; jb < n; jb += PB) {
    const int bw = (n - jb < PB) ? (n - jb) : PB;
    const int i = jb + tid;  // this thread's row in the panel
    const bool act = (tid < LB) && (i < n);
    double a[PB];
#pragma unroll
    for (int c = 0; c < PB; ++c) a[c] = (act && c < bw) ? S[i * LS + jb + c] : 0.0;
#pragma unroll
    for (int c = 0; c < PB; ++c) {
      if (c < bw) {  // uniform
        if (act && i == jb + c) lsh[8] = a[c];
        __syncthreads();
        const double piv = lsh[8];
        const double rd = rsqrt(piv);  // NaN for a non-PD pivot: propagates like JAX
        if (act) {
          if (i == jb + c) {
            a[c] = piv * rd;
            rdg[i * LS] = rd;
          } else if (i > jb + c) {
            a[c] = a[c] * rd;
          }
          if (i - jb < PB) lsh[i - jb] = a[c];
        }
        __syncthreads();
        if (act) {
#pragma unroll
          for (int c2 = c + 1; c2 < PB; ++c2)
            if (c2 < bw && i >= jb + c2) a[c2] -= a[c] * lsh[c2];
        }
      }
    }
    if (act) {
#pragma unroll
      for (int c = 0; c < PB; ++c) {
        if (c < bw && jb + c <= i) S[i * LS + jb + c] = a[c];
        if (i >= jb + PB) PT[c * PTS + i] = (c < bw) ? a[c] : 0.0;
      }
    }
    __syncthreads();
    // rank-8 trailing update on 4x4 register tiles of the lower triangle
    const int t0 = jb + PB;
    const int t = n - t0;
    if (t > 0) {
      const int nt = (t + 3) >> 2;
      const int ntiles = nt * (nt + 1) / 2;
      for (int idx = tid; idx < ntiles; idx += LEAF_NT) {
        int ti = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while (ti * (ti + 1) / 2 > idx) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= idx) ++ti;
        const int tk = idx - ti * (ti + 1) / 2;
        const int i0 = t0 + 4 * ti, k0 = t0 + 4 * tk;
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
#pragma unroll
        for (int c = 0; c < PB; ++c) {
          double pi[4], pk[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            pi[x] = (i0 + x < n) ? PT[c * PTS + i0 + x] : 0.0;
            pk[x] = (k0 + x < n) ? PT[c * PTS + k0 + x] : 0.0;
          }
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] += pi[x] * pk[y];
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            const int r = i0 + x, k = k0 + y;
            if (r < n && k <= r) S[r * LS + k] -= acc[x][y];
          }
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < n * n; idx += LEAF_NT) {
    const int i = idx / n, k = idx - i * n;
    if (k <= i) A[(long long)i * lda + k] = S[i * LS + k];
  }
  if (inv == nullptr) return;
  __syncthreads();

  // ---------------- in-place inverse, block columns right to left ----------------
  // X11 = inv(L11);  X21 = -X22 L21 X11  with X22 (rows/cols > block) already inverted in place.
  const int jlast = ((n - 1) / PB) * PB;
  for (int jb = jlast; jb >= 0; jb -= PB) {
    const int bw = (n - jb < PB) ? (n - jb) : PB;
    // inverse of the 8x8 diagonal block: lane c of the LAST warp (idle otherwise) solves column c
    // by forward substitution, multiplying by the reciprocals saved during the factorisation
    if (tid >= LEAF_NT - 32 && tid < LEAF_NT - 32 + PB) {
      const int c = tid - (LEAF_NT - 32);
      double x[PB];
#pragma unroll
      for (int r = 0; r < PB; ++r) {
        double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < PB; ++k)
          if (k < r && k >= c && r < bw) v -= S[(jb + r) * LS + jb + k] * x[k];
        x[r] = (r < bw && r >= c && c < bw) ? v * rdg[(jb + r) * LS] : 0.0;
      }
#pragma unroll
      for (int r = 0; r < PB; ++r) xbb[r * PB + c] = x[r];
    }
    const int i = jb + PB + tid;  // one thread per row below the block
    const bool act = (tid < LB) && (i < n);
    double r8[PB];
#pragma unroll
    for (int c = 0; c < PB; ++c) r8[c] = 0.0;
    if (act) {
#pragma unroll 4
      for (int k = jb + PB; k <= i; ++k) {
        const double xik = S[i * LS + k];
        const double* lk = S + k * LS + jb;
#pragma unroll
        for (int c = 0; c < PB; ++c) r8[c] += xik * lk[c];
      }
    }
    __syncthreads();  // xbb ready; every read of the L panel (columns jb..jb+7) is done
    if (act) {
#pragma unroll
      for (int c = 0; c < PB; ++c) {
        double v = 0.0;
#pragma unroll
        for (int c2 = 0; c2 < PB; ++c2)
          if (c2 >= c) v += r8[c2] * xbb[c2 * PB + c];
        if (c < bw) S[i * LS + jb + c] = -v;
      }
    }
    if (tid < PB * PB) {
      const int r = tid / PB, c = tid % PB;
      if (r < bw && c <= r) S[(jb + r) * LS + jb + c] = xbb[r * PB + c];
    }
    __syncthreads();
  }
  for (int idx = tid; idx < LB * LB; idx += LEAF_NT) {
    const int i = idx >> 7, k = idx & (LB - 1);
    inv[idx] = (i < n && k <= i) ? S[i * LS + k] : 0.0;
  }
}

__global__ void logdet_kernel(const double* __restrict__ L, long long ld, int n,
                              double* __restrict__ out, double scale) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += log(L[(long long)i * ld + i]);
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = scale * s;
}

__global__ void copy_block_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                  long long ldd, int n) {
  // src: packed 128x128 leaf inverse; dst: n x n lower block
  for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < n * n; idx += blockDim.x * gridDim.x) {
    const int i = idx / n, k = idx - i * n;
    if (k <= i) dst[(long long)i * ldd + k] = src[i * LB + k];
  }
}

int leaf(Ctx* ctx, double* A, long long lda, int n, double* inv, cudaStream_t s) {
  GPXB_CUDA_CHECK(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       LEAF_SMEM));
  potrf_leaf_kernel<<<1, LEAF_NT, LEAF_SMEM, s>>>(A, lda, n, inv);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

inline int split_point(int n) {
  const int nb = (n + LB - 1) / LB;
  return (nb / 2) * LB;
}

}  // namespace

namespace {
constexpr int TRSV_MAXM = 4;

// x_k = inv_kk b_k (forward) or inv_kk^T b_k (backward) for m right-hand sides stored as rows of B
// (ld ldb); one CTA of 256 threads, in place.  Both directions read the 128x128 inverse coalesced:
// forward: one warp per output row (lanes over the columns, warp reduction);
// backward: one thread per output index (lanes over consecutive columns of inv).
__global__ void __launch_bounds__(256) trsv_diag_kernel(const double* __restrict__ inv, double* __restrict__ B,
                                                        long long ldb, int m, int nb, int backward) {
  __shared__ double sb[TRSV_MAXM][LB];
  __shared__ double so[TRSV_MAXM][LB];
  const int tid = threadIdx.x;
  for (int i = tid; i < TRSV_MAXM * LB; i += 256) {
    const int r = i / LB, c = i % LB;
    sb[r][c] = (r < m && c < nb) ? B[(long long)r * ldb + c] : 0.0;
  }
  __syncthreads();
  if (!backward) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int t = warp; t < nb; t += 8) {  // x[t] = sum_{c <= t} inv[t][c] b[c]
      double acc[TRSV_MAXM];
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
      for (int c = lane; c <= t; c += 32) {
        const double v = inv[t * LB + c];
#pragma unroll
        for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sb[r][c];
      }
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) {
        const double sum = warp_sum(acc[r]);
        if (lane == 0) so[r][t] = sum;
      }
    }
  } else if (tid < nb) {
    const int t = tid;  // x[t] = sum_{c >= t} inv[c][t] b[c]
    double acc[TRSV_MAXM];
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
    for (int c = t; c < nb; ++c) {
      const double v = inv[c * LB + t];
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sb[r][c];
    }
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) so[r][t] = acc[r];
  }
  __syncthreads();
  for (int i = tid; i < m * nb; i += 256) {
    const int r = i / nb, c = i % nb;
    B[(long long)r * ldb + c] = so[r][c];
  }
}

// forward: B[:, j] -= sum_c L[j][c] x[c] for the rows j below the block (one warp per row j)
__global__ void __launch_bounds__(256) trsv_fwd_update_kernel(const double* __restrict__ Lblk, long long ldl,
                                                              const double* __restrict__ X, double* __restrict__ B,
                                                              long long ldb, int m, int nb, int nrows) {
  __shared__ double sx[TRSV_MAXM][LB];
  for (int i = threadIdx.x; i < TRSV_MAXM * LB; i += 256) {
    const int r = i / LB, c = i % LB;
    sx[r][c] = (r < m && c < nb) ? X[(long long)r * ldb + c] : 0.0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= nrows) return;
  const double* lrow = Lblk + (long long)j * ldl;
  double acc[TRSV_MAXM];
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
  for (int c = lane; c < nb; c += 32) {
    const double v = lrow[c];
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sx[r][c];
  }
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r) {
    const double sum = warp_sum(acc[r]);
    if (lane == 0 && r < m) B[(long long)r * ldb + j] -= sum;
  }
}

// backward: B[:, j] -= sum_c L[c][j] x[c] for the columns j left of the block (one thread per j)
__global__ void __launch_bounds__(256) trsv_bwd_update_kernel(const double* __restrict__ Lrows, long long ldl,
                                                              const double* __restrict__ X, double* __restrict__ B,
                                                              long long ldb, int m, int nb, int ncols) {
  __shared__ double sx[TRSV_MAXM][LB];
  for (int i = threadIdx.x; i < TRSV_MAXM * LB; i += 256) {
    const int r = i / LB, c = i % LB;
    sx[r][c] = (r < m && c < nb) ? X[(long long)r * ldb + c] : 0.0;
  }
  __syncthreads();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= ncols) return;
  double acc[TRSV_MAXM];
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
#pragma unroll 4
  for (int c = 0; c < nb; ++c) {
    const double v = Lrows[(long long)c * ldl + j];
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sx[r][c];
  }
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r)
    if (r < m) B[(long long)r * ldb + j] -= acc[r];
}

// backward update on the block-column packed layout: the row strip L[c0:c0+nb, 0:ncols] crosses panels
__global__ void __launch_bounds__(256) trsv_bwd_update_bcp_kernel(const double* __restrict__ Lp, BcpLayout lay, int c0,
                                                                  const double* __restrict__ X, double* __restrict__ B,
                                                                  long long ldb, int m, int nb, int ncols) {
  __shared__ double sx[TRSV_MAXM][LB];
  for (int i = threadIdx.x; i < TRSV_MAXM * LB; i += 256) {
    const int r = i / LB, c = i % LB;
    sx[r][c] = (r < m && c < nb) ? X[(long long)r * ldb + c] : 0.0;
  }
  __syncthreads();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= ncols) return;
  const double* Lrows = Lp + lay.index(c0, j);
  double acc[TRSV_MAXM];
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
#pragma unroll 4
  for (int c = 0; c < nb; ++c) {
    const double v = Lrows[(long long)c * BcpLayout::NB];
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sx[r][c];
  }
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r)
    if (r < m) B[(long long)r * ldb + j] -= acc[r];
}

__global__ void logdet_bcp_kernel(const double* __restrict__ Lp, BcpLayout lay, double* __restrict__ out,
                                  double scale) {
  __shared__ double sh[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < lay.n; i += 1024) s += log(Lp[lay.index(i, i)]);
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = scale * s;
}

// few right-hand sides (m <= 4): blocked substitution with GEMV updates -- HBM bound (reads L once)
int trsv_few(Ctx* ctx, const double* L, long long ldl, const double* inv, double* B, long long ldb, int m, int n,
             bool transposed_solve, cudaStream_t s) {
  const int nblk = (n + LB - 1) / LB;
  if (transposed_solve) {
    // X L^T = B  <=>  L x = b: forward over blocks
    for (int k = 0; k < nblk; ++k) {
      const int c0 = k * LB, nb = min(LB, n - c0);
      trsv_diag_kernel<<<1, 256, 0, s>>>(inv + (long long)k * LB * LB, B + c0, ldb, m, nb, 0);
      const int below = n - c0 - nb;
      if (below > 0)
        trsv_fwd_update_kernel<<<(below + 7) / 8, 256, 0, s>>>(L + (long long)(c0 + nb) * ldl + c0, ldl, B + c0,
                                                                B + c0 + nb, ldb, m, nb, below);
      ctx->launches += below > 0 ? 2 : 1;
    }
  } else {
    // X L = B  <=>  L^T x = b: backward over blocks
    for (int k = nblk - 1; k >= 0; --k) {
      const int c0 = k * LB, nb = min(LB, n - c0);
      trsv_diag_kernel<<<1, 256, 0, s>>>(inv + (long long)k * LB * LB, B + c0, ldb, m, nb, 1);
      if (c0 > 0)
        trsv_bwd_update_kernel<<<(c0 + 255) / 256, 256, 0, s>>>(L + (long long)c0 * ldl, ldl, B + c0, B, ldb, m, nb,
                                                                 c0);
      ctx->launches += c0 > 0 ? 2 : 1;
    }
  }
  GPXB_LAUNCH_CHECK();
  return 0;
}
}  // namespace

// X * L^T = B  (X = B L^-T), B is m x n row-major, overwritten.  inv: leaf inverses of L's
// diagonal blocks (128x128 packed each), L n x n lower.
int trsm_right_lt(Ctx* ctx, const double* L, long long ldl, const double* inv, double* B,
                  long long ldb, int m, int n, cudaStream_t s) {
  if (m <= 0 || n <= 0) return 0;
  if (m <= TRSV_MAXM && n > LB) return trsv_few(ctx, L, ldl, inv, B, ldb, m, n, true, s);
  if (n <= LB) {
    GemmArgs g;
    g.M = m; g.N = n; g.K = n;
    g.A = B; g.lda = ldb; g.a_kmajor = true;
    g.B = inv; g.ldb = LB; g.b_kmajor = true;  // (B inv^T)[i][j] = sum_k B[i][k] inv[j][k]
    g.C = B; g.ldc = ldb;                       // in place: single column tile, rows private per CTA
    return gemm_f64(ctx, g, s);
  }
  const int n1 = split_point(n), n2 = n - n1;
  GPXB_TRY(trsm_right_lt(ctx, L, ldl, inv, B, ldb, m, n1, s));
  {
    GemmArgs g;  // B2 -= X1 * L21^T
    g.M = m; g.N = n2; g.K = n1;
    g.A = B; g.lda = ldb; g.a_kmajor = true;
    g.B = L + (long long)n1 * ldl; g.ldb = ldl; g.b_kmajor = true;
    g.C = B + n1; g.ldc = ldb;
    g.alpha = -1.0; g.beta = 1.0;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  return trsm_right_lt(ctx, L + (long long)n1 * ldl + n1, ldl, inv + (long long)(n1 / LB) * LB * LB,
                       B + n1, ldb, m, n2, s);
}

// X * L = B  (X = B L^-1), B is m x n row-major, overwritten.
int trsm_right_l(Ctx* ctx, const double* L, long long ldl, const double* inv, double* B,
                 long long ldb, int m, int n, cudaStream_t s) {
  if (m <= 0 || n <= 0) return 0;
  if (m <= TRSV_MAXM && n > LB) return trsv_few(ctx, L, ldl, inv, B, ldb, m, n, false, s);
  if (n <= LB) {
    GemmArgs g;
    g.M = m; g.N = n; g.K = n;
    g.A = B; g.lda = ldb; g.a_kmajor = true;
    g.B = inv; g.ldb = LB; g.b_kmajor = false;  // (B inv)[i][j] = sum_k B[i][k] inv[k][j]
    g.C = B; g.ldc = ldb;
    return gemm_f64(ctx, g, s);
  }
  const int n1 = split_point(n), n2 = n - n1;
  GPXB_TRY(trsm_right_l(ctx, L + (long long)n1 * ldl + n1, ldl, inv + (long long)(n1 / LB) * LB * LB,
                        B + n1, ldb, m, n2, s));
  {
    GemmArgs g;  // B1 -= X2 * L21
    g.M = m; g.N = n1; g.K = n2;
    g.A = B + n1; g.lda = ldb; g.a_kmajor = true;
    g.B = L + (long long)n1 * ldl; g.ldb = ldl; g.b_kmajor = false;
    g.C = B; g.ldc = ldb;
    g.alpha = -1.0; g.beta = 1.0;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  return trsm_right_l(ctx, L, ldl, inv, B, ldb, m, n1, s);
}

// In-place lower Cholesky of the n x n row-major matrix A (only the lower triangle is
// read or written).  inv receives ceil(n/128) packed 128x128 inverses of L's diagonal blocks.
//
// Two-level right-looking: 1024-wide panels; inside a panel each 128-column step is
// leaf (one CTA) -> tall TRSM (multiply by the inverted leaf) -> rank-128 update of the rest
// of the panel; then one rank-1024 SYRK on the trailing matrix.  Every GEMM is tall, so only
// the leaves run on a single SM.
namespace {
// Factor the tall panel A[k0:n, k0:k0+w] (its w x w head is the diagonal block).
int factor_panel(Ctx* ctx, double* A, long long lda, int n, int k0, int w, double* inv,
                 cudaStream_t s) {
  for (int j = 0; j < w; j += LB) {
    const int jb = min(LB, w - j);
    const int c0 = k0 + j;  // global first column of this step
    double* Ajj = A + (long long)c0 * lda + c0;
    double* invj = inv + (long long)(c0 / LB) * LB * LB;
    GPXB_TRY(leaf(ctx, Ajj, lda, jb, invj, s));
    const int below = n - c0 - jb;
    if (below <= 0) continue;
    double* B = A + (long long)(c0 + jb) * lda + c0;  // below x jb
    {
      GemmArgs g;  // B <- B inv^T (in place: one column tile, rows private per CTA)
      g.M = below; g.N = jb; g.K = jb;
      g.A = B; g.lda = lda; g.a_kmajor = true;
      g.B = invj; g.ldb = LB; g.b_kmajor = true;
      g.C = B; g.ldc = lda;
      GPXB_TRY(gemm_f64(ctx, g, s));
    }
    const int wr = w - j - jb;  // panel columns still to update
    if (wr > 0) {
      GemmArgs g;  // C -= B B[0:wr]^T on the lower trapezoid
      g.M = below; g.N = wr; g.K = jb;
      g.A = B; g.lda = lda; g.a_kmajor = true;
      g.B = B; g.ldb = lda; g.b_kmajor = true;
      g.C = B + jb; g.ldc = lda;
      g.alpha = -1.0; g.beta = 1.0;
      g.lower_only = 1;
      GPXB_TRY(gemm_f64(ctx, g, s));
    }
  }
  return 0;
}

// C[r0:n, c0:c0+nc] -= P[r0.., :] P[c0.., :]^T on the lower trapezoid, P = A[:, k0:k0+w].
int panel_update(Ctx* ctx, double* A, long long lda, int n, int k0, int w, int c0, int nc,
                 cudaStream_t s) {
  const int m = n - c0;
  if (m <= 0 || nc <= 0) return 0;
  double* P = A + (long long)c0 * lda + k0;
  GemmArgs g;
  g.M = m; g.N = nc; g.K = w;
  g.A = P; g.lda = lda; g.a_kmajor = true;
  g.B = P; g.ldb = lda; g.b_kmajor = true;
  g.C = A + (long long)c0 * lda + c0; g.ldc = lda;
  g.alpha = -1.0; g.beta = 1.0;
  g.lower_only = 1;
  return gemm_f64(ctx, g, s);
}
}  // namespace

int potrf_lower(Ctx* ctx, double* A, long long lda, int n, double* inv, cudaStream_t s) {
  if (n <= 0) return 0;
  constexpr int NB = 1024;
  if (n <= NB) return factor_panel(ctx, A, lda, n, 0, n, inv, s);
  // Depth-1 look-ahead: panel p+1 is factorised on the high-priority stream hp while the
  // caller's stream s applies panel p's rank-NB update to the columns right of panel p+1.
  cudaStream_t hp = ctx->side;
  cudaEvent_t ev, ev_rest_prev = nullptr;
  GPXB_TRY(ctx->next_event(&ev));
  GPXB_CUDA_CHECK(cudaEventRecord(ev, s));
  GPXB_CUDA_CHECK(cudaStreamWaitEvent(hp, ev, 0));
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int w = min(NB, n - k0);
    GPXB_TRY(factor_panel(ctx, A, lda, n, k0, w, inv, hp));
    const int c1 = k0 + w;  // first column of the next panel
    if (c1 >= n) break;
    const int w1 = min(NB, n - c1);
    cudaEvent_t ev_factor;
    GPXB_TRY(ctx->next_event(&ev_factor));
    GPXB_CUDA_CHECK(cudaEventRecord(ev_factor, hp));
    // look-ahead part (next panel's columns) on hp, after the previous rest-update wrote them
    if (ev_rest_prev) GPXB_CUDA_CHECK(cudaStreamWaitEvent(hp, ev_rest_prev, 0));
    GPXB_TRY(panel_update(ctx, A, lda, n, k0, w, c1, w1, hp));
    // the rest of the trailing matrix on s
    const int c2 = c1 + w1;
    if (c2 < n) {
      GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_factor, 0));
      GPXB_TRY(panel_update(ctx, A, lda, n, k0, w, c2, n - c2, s));
      GPXB_TRY(ctx->next_event(&ev_rest_prev));
      GPXB_CUDA_CHECK(cudaEventRecord(ev_rest_prev, s));
    } else {
      ev_rest_prev = nullptr;
    }
  }
  GPXB_TRY(ctx->next_event(&ev));
  GPXB_CUDA_CHECK(cudaEventRecord(ev, hp));
  GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
  return 0;
}

// Same algorithm on the block-column packed layout (chol.cuh: BcpLayout): the trailing update is one
// trapezoid GEMM per block column instead of one for the whole trailing matrix.
int potrf_lower_bcp(Ctx* ctx, double* Ap, int n, double* inv, cudaStream_t s) {
  if (n <= 0) return 0;
  constexpr int NB = BcpLayout::NB;
  const BcpLayout lay{n};
  if (n <= NB) return factor_panel(ctx, Ap, NB, n, 0, n, inv, s);
  const int np = lay.panels();
  auto update = [&](int p, int q, cudaStream_t st) -> int {  // block column q -= L[q.., p] L[q, p]^T
    const long long k0 = (long long)p * NB, c0 = (long long)q * NB;
    const double* P = lay.vbase(Ap, p) + c0 * NB + k0;
    GemmArgs g;
    g.M = (int)(n - c0); g.N = (int)std::min<long long>(NB, n - c0); g.K = NB;
    g.A = P; g.lda = NB; g.a_kmajor = true;
    g.B = P; g.ldb = NB; g.b_kmajor = true;
    g.C = lay.vbase(Ap, q) + c0 * NB + c0; g.ldc = NB;
    g.alpha = -1.0; g.beta = 1.0;
    g.lower_only = 1;
    return gemm_f64(ctx, g, st);
  };
  cudaStream_t hp = ctx->side;
  cudaEvent_t ev, ev_next_prev = nullptr;
  GPXB_TRY(ctx->next_event(&ev));
  GPXB_CUDA_CHECK(cudaEventRecord(ev, s));
  GPXB_CUDA_CHECK(cudaStreamWaitEvent(hp, ev, 0));
  for (int p = 0; p < np; ++p) {
    const int k0 = p * NB, w = min(NB, n - k0);
    GPXB_TRY(factor_panel(ctx, lay.vbase(Ap, p), NB, n, k0, w, inv, hp));
    if (p + 1 >= np) break;
    cudaEvent_t ev_factor;
    GPXB_TRY(ctx->next_event(&ev_factor));
    GPXB_CUDA_CHECK(cudaEventRecord(ev_factor, hp));
    // look-ahead: block column p+1 on hp, once the earlier panels' updates of it (on s) are done
    if (ev_next_prev) GPXB_CUDA_CHECK(cudaStreamWaitEvent(hp, ev_next_prev, 0));
    GPXB_TRY(update(p, p + 1, hp));
    ev_next_prev = nullptr;
    if (p + 2 < np) {
      GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_factor, 0));
      for (int q = p + 2; q < np; ++q) {
        GPXB_TRY(update(p, q, s));
        if (q == p + 2) {  // all the next look-ahead needs
          GPXB_TRY(ctx->next_event(&ev_next_prev));
          GPXB_CUDA_CHECK(cudaEventRecord(ev_next_prev, s));
        }
      }
    }
  }
  GPXB_TRY(ctx->next_event(&ev));
  GPXB_CUDA_CHECK(cudaEventRecord(ev, hp));
  GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
  return 0;
}

int trsv_bcp(Ctx* ctx, const double* Lp, int n, const double* inv, double* B, long long ldb, int m, bool forward,
             cudaStream_t s) {
  GPXB_ARG_CHECK(m >= 1 && m <= TRSV_MAXM, -30);
  constexpr int NB = BcpLayout::NB;
  const BcpLayout lay{n};
  const int nblk = (n + LB - 1) / LB;
  if (forward) {
    for (int k = 0; k < nblk; ++k) {
      const int c0 = k * LB, nb = min(LB, n - c0);
      trsv_diag_kernel<<<1, 256, 0, s>>>(inv + (long long)k * LB * LB, B + c0, ldb, m, nb, 0);
      const int below = n - c0 - nb;
      if (below > 0)
        trsv_fwd_update_kernel<<<(below + 7) / 8, 256, 0, s>>>(
            lay.vbase(Lp, c0 / NB) + (long long)(c0 + nb) * NB + c0, NB, B + c0, B + c0 + nb, ldb, m, nb, below);
      ctx->launches += below > 0 ? 2 : 1;
    }
  } else {
    for (int k = nblk - 1; k >= 0; --k) {
      const int c0 = k * LB, nb = min(LB, n - c0);
      trsv_diag_kernel<<<1, 256, 0, s>>>(inv + (long long)k * LB * LB, B + c0, ldb, m, nb, 1);
      if (c0 > 0)
        trsv_bwd_update_bcp_kernel<<<(c0 + 255) / 256, 256, 0, s>>>(Lp, lay, c0, B + c0, B, ldb, m, nb, c0);
      ctx->launches += c0 > 0 ? 2 : 1;
    }
  }
  GPXB_LAUNCH_CHECK();
  return 0;
}

int logdet_bcp(Ctx* ctx, const double* Lp, int n, double* out, double scale, cudaStream_t s) {
  logdet_bcp_kernel<<<1, 1024, 0, s>>>(Lp, BcpLayout{n}, out, scale);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

namespace {
size_t trtri_scratch(int n) {
  if (n <= LB) return 0;
  const int n1 = split_point(n), n2 = n - n1;
  return (size_t)n1 * n2 + trtri_scratch(n1) + trtri_scratch(n2);
}

// W (n x n, zero-initialised, lower) <- L^-1.  T: trtri_scratch(n) doubles.  The two half-size
// inversions are independent; down to depth 2 they run on separate streams (4-way concurrency)
// so that the small GEMMs deep in the recursion still fill the GPU.  st[0] is the caller's stream.
int trtri_rec(Ctx* ctx, const double* L, long long ldl, const double* inv, double* W, long long ldw,
              int n, double* T, cudaStream_t* st, int depth, int branch) {
  cudaStream_t s = st[depth <= 2 ? branch * (4 >> depth) : 0];
  if (depth > 2) s = st[branch];  // below the fork levels 'branch' carries the stream index
  if (n <= LB) {
    copy_block_kernel<<<32, 256, 0, s>>>(inv, W, ldw, n);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
    return 0;
  }
  const int n1 = split_point(n), n2 = n - n1;
  const double* inv2 = inv + (long long)(n1 / LB) * LB * LB;
  double* Tc1 = T + (size_t)n1 * n2;
  double* Tc2 = Tc1 + trtri_scratch(n1);
  const double* L22 = L + (long long)n1 * ldl + n1;
  double* W22 = W + (long long)n1 * ldw + n1;
  if (depth < 2 && n >= 2048) {
    const int b1 = 2 * branch, b2 = 2 * branch + 1;
    cudaStream_t s2 = st[b2 * (4 >> (depth + 1))];
    cudaEvent_t e0, e1;
    GPXB_TRY(ctx->next_event(&e0));
    GPXB_CUDA_CHECK(cudaEventRecord(e0, s));
    GPXB_CUDA_CHECK(cudaStreamWaitEvent(s2, e0, 0));
    GPXB_TRY(trtri_rec(ctx, L, ldl, inv, W, ldw, n1, Tc1, st, depth + 1, b1));
    GPXB_TRY(trtri_rec(ctx, L22, ldl, inv2, W22, ldw, n2, Tc2, st, depth + 1, b2));
    GPXB_TRY(ctx->next_event(&e1));
    GPXB_CUDA_CHECK(cudaEventRecord(e1, s2));
    GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, e1, 0));
  } else {
    // sequential on this node's stream
    const int sidx = (depth <= 2) ? branch * (4 >> depth) : branch;
    GPXB_TRY(trtri_rec(ctx, L, ldl, inv, W, ldw, n1, Tc1, st, 3, sidx));
    GPXB_TRY(trtri_rec(ctx, L22, ldl, inv2, W22, ldw, n2, Tc2, st, 3, sidx));
  }
  {
    GemmArgs g;  // T = L21 * W11   (W11 lower: k starts at the tile's first column)
    g.M = n2; g.N = n1; g.K = n1;
    g.A = L + (long long)n1 * ldl; g.lda = ldl; g.a_kmajor = true;
    g.B = W; g.ldb = ldw; g.b_kmajor = false;
    g.C = T; g.ldc = n1;
    g.klo_mode = 1;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  {
    GemmArgs g;  // W21 = -W22 * T   (W22 lower: k ends at the tile's last row)
    g.M = n2; g.N = n1; g.K = n2;
    g.A = W22; g.lda = ldw; g.a_kmajor = true;
    g.B = T; g.ldb = n1; g.b_kmajor = false;
    g.C = W + (long long)n1 * ldw; g.ldc = ldw;
    g.alpha = -1.0;
    g.khi_mode = 1;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  return 0;
}
}  // namespace

// Ainv (lower triangle, n x n, ld ldo) <- (L L^T)^-1 given the factor L and its leaf inverses.
// W: n x n workspace (ld n), T: potri_scratch_doubles(n) workspace.  out may alias L's storage.
int potri_lower(Ctx* ctx, const double* L, long long ldl, const double* inv, int n, double* W,
                double* T, double* out, long long ldo, cudaStream_t s) {
  if (n <= 0) return 0;
  GPXB_CUDA_CHECK(cudaMemsetAsync(W, 0, sizeof(double) * (size_t)n * (size_t)n, s));
  cudaStream_t st[4] = {s, ctx->aux[0], ctx->aux[1], ctx->aux[2]};
  GPXB_TRY(trtri_rec(ctx, L, ldl, inv, W, n, n, T, st, 0, 0));
  GemmArgs g;  // out = W^T W (lower); W[k][i] = 0 for k < i -> k starts at the tile's first row
  g.M = n; g.N = n; g.K = n;
  g.A = W; g.lda = n; g.a_kmajor = false;
  g.B = W; g.ldb = n; g.b_kmajor = false;
  g.C = out; g.ldc = ldo;
  g.lower_only = 1;
  g.klo_mode = 2;
  return gemm_f64(ctx, g, s);
}

int logdet_lower(Ctx* ctx, const double* L, long long ld, int n, double* out, double scale,
                 cudaStream_t s) {
  logdet_kernel<<<1, 1024, 0, s>>>(L, ld, n, out, scale);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

size_t potri_scratch_doubles(int n) { return trtri_scratch(n) + 64; }

}  // namespace gpxb
