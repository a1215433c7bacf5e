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
#include <cstdlib>
#include <vector>

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

// ---------------------------------------------------------------------------------------------------
// Leaf, second generation (default): the same contract as potrf_leaf_kernel, organised so that only the
// four 32x32 diagonal blocks are eliminated column by column -- by ONE warp each, rows in registers, pivots
// and multipliers exchanged by warp shuffles (no block barrier inside the 32 column steps) -- while every
// off-diagonal operation (panel solve against the inverted diagonal block, rank-32 trailing update, the two
// merge levels of the inverse) runs on the FP64 tensor pipe from shared memory.  12 block barriers in the
// factorisation instead of 256.
#ifdef GPXB_LEAF_STAMPS  // developer build (tools/leaf_timing.cu): merge >> 8 names a phase after which the kernel returns
#define LEAF_STAMP(k) do { if ((merge >> 8) == (k) + 1) return; } while (0)
// timeline of the panel's chain kernels: (kernel id, begin, end) in ns of %globaltimer, block 0 / thread 0
__device__ unsigned long long g_chain_log[3 * 4096];
__device__ unsigned int g_chain_n;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CHAIN_BEGIN(id)                                                              \
  unsigned int chain_slot = 0;                                                       \
  if (threadIdx.x == 0 && blockIdx.x == 0) {                                         \
    chain_slot = atomicAdd(&g_chain_n, 1u) & 4095u;                                  \
    g_chain_log[3 * chain_slot] = (id);                                              \
    g_chain_log[3 * chain_slot + 1] = gtimer();                                      \
  }
#define CHAIN_END()                                                                  \
  do {                                                                               \
    if (threadIdx.x == 0 && blockIdx.x == 0) g_chain_log[3 * chain_slot + 2] = gtimer(); \
  } while (0)
#else
#define LEAF_STAMP(k) do {} while (0)
#define CHAIN_BEGIN(id)
#define CHAIN_END() do {} while (0)
#endif
constexpr int L2S = 132;   // row stride of the 128x128 matrix in smem (== 4 mod 16: conflict-free fragment reads)
constexpr int DIS = 36;    // row stride of a 32x32 diagonal-block inverse
constexpr int T2S = 68;    // row stride of the 64x64 scratch
constexpr int LEAF2_SMEM = (LB * L2S + 4 * 32 * DIS + 64 * T2S + LB) * (int)sizeof(double);

// C[M x N] (ldc) = alpha * A[M x K] (row-major, lda) * B[K x N] (row-major, ldb); 8x8 tiles over the CTA's 8 warps
__device__ __forceinline__ void cta_mm(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M,
                                       int N, int K, double alpha, int warp, int lane) {
  const int lr = lane >> 2, lq = lane & 3;
  const int tn = N >> 3, nt = (M >> 3) * tn;
  for (int t = warp; t < nt; t += 8) {
    const int ti = t / tn, tj = t - ti * tn;
    double c0 = 0.0, c1 = 0.0;
    const double* ap = A + (ti * 8 + lr) * lda + lq;
    const double* bp = B + lq * ldb + tj * 8 + lr;
    for (int k = 0; k < K; k += 4) dmma884(c0, c1, ap[k], bp[k * ldb]);
    double* cp = C + (ti * 8 + lr) * ldc + tj * 8 + 2 * lq;
    cp[0] = alpha * c0;
    cp[1] = alpha * c1;
  }
}

// C = alpha * A B on (8 NT)^2 operands in shared memory with ONE lower-triangular factor.  A_LOWER: A is lower
// triangular (k < 8 (w + 1)) and the warp owns tile row w; otherwise B is (k >= 8 w) and the warp owns tile column w.
template <int NT, bool A_LOWER>
__device__ __forceinline__ void tri_mm(double* C, int ldc, const double* A, int lda, const double* B, int ldb,
                                       double alpha, int w, int lane) {
  const int lr = lane >> 2, lq = lane & 3;
#pragma unroll
  for (int g = 0; g < NT; g += 4) {
    double c0[4] = {0.0, 0.0, 0.0, 0.0}, c1[4] = {0.0, 0.0, 0.0, 0.0};
    if (A_LOWER) {
      const double* ap = A + (w * 8 + lr) * lda + lq;
      const double* bp = B + lq * ldb + g * 8 + lr;
      for (int k = 0; k < (w + 1) * 8; k += 4) {
        const double a = ap[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(c0[q], c1[q], a, bp[k * ldb + q * 8]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double* cp = C + (w * 8 + lr) * ldc + (g + q) * 8 + 2 * lq;
        cp[0] = alpha * c0[q];
        cp[1] = alpha * c1[q];
      }
    } else {
      const double* ap = A + (g * 8 + lr) * lda + lq;
      const double* bp = B + lq * ldb + w * 8 + lr;
      for (int k = w * 8; k < NT * 8; k += 4) {
        const double b = bp[k * ldb];
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(c0[q], c1[q], ap[q * 8 * lda + k], b);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double* cp = C + ((g + q) * 8 + lr) * ldc + w * 8 + 2 * lq;
        cp[0] = alpha * c0[q];
        cp[1] = alpha * c1[q];
      }
    }
  }
}

// The merge levels of the packed inverse: S holds L with its four 32 x 32 diagonal blocks already replaced by their
// inverses; X21 = -X22 (L21 X11) for the 32 -> 64 and 64 -> 128 levels on DMMA, then the 128 x 128 block goes to inv.
template <bool BULK_OUT>
__device__ __forceinline__ void leaf_merge_and_store(double* S, double* T, int n, double* __restrict__ inv, int tid,
                                                     int warp, int lane, bool skip_diag_blocks) {
  // every product has one triangular factor: the k range is cut to its non-zero part (exact zeros skipped: the
  // sums are unchanged) and each warp runs four independent accumulator chains that share one fragment load
  {
    const int p = (warp >> 2) * 64, q = warp & 3;
    tri_mm<4, false>(T + (p >> 1) * T2S, T2S, S + (p + 32) * L2S + p, L2S, S + p * L2S + p, L2S, 1.0, q, lane);
    __syncthreads();  // T_p = L21 X11   (32 x 32 x 32, X11 lower)
    tri_mm<4, true>(S + (p + 32) * L2S + p, L2S, S + (p + 32) * L2S + p + 32, L2S, T + (p >> 1) * T2S, T2S, -1.0, q,
                    lane);
    __syncthreads();  // X21 = -X22 T_p  (X22 lower)
  }
  tri_mm<8, false>(T, T2S, S + 64 * L2S, L2S, S, L2S, 1.0, warp, lane);
  __syncthreads();
  tri_mm<8, true>(S + 64 * L2S, L2S, S + 64 * L2S + 64, L2S, T, T2S, -1.0, warp, lane);
  if (BULK_OUT) {  // S is the packed inverse exactly (zero upper triangle): one 1 KB bulk store per row
    fence_proxy_async();
    __syncthreads();
    if (tid < LB) {
      tma_bulk_s2g(inv + tid * LB, S + tid * L2S, LB * (int)sizeof(double));
      tma_bulk_commit_wait_read();
    }
    return;
  }
  __syncthreads();
#pragma unroll 8
  for (int idx = tid; idx < LB * LB; idx += LEAF_NT) {
    const int i = idx >> 7, k = idx & (LB - 1);
    if (skip_diag_blocks && (i >> 5) == (k >> 5)) continue;  // final already, and being read by the head solve
    inv[idx] = (i < n && k <= i) ? S[i * L2S + k] : 0.0;
  }
}

// merge == 0: the kernel stops after the factor and the four diagonal-block inverses (written into the diagonal
// 32 x 32 blocks of inv): the merge levels and the 128 KB store of the packed inverse leave the panel's dependency
// chain -- the head rows are solved by blocked substitution on L and those four blocks (panel_trsm32s_kernel), and
// potrf_leaf_merge_kernel completes inv beside the chain, before the tall solve needs it.
// Inverse of the 32 x 32 lower-triangular block at Lb (row stride L2S; rd: reciprocals of its diagonal) into Dinv (row
// stride DIS, zeros above the diagonal), by ONE warp: the four 8 x 8 diagonal blocks by forward substitution (lane =
// block x column, 8 dependent steps), then the levels 8 -> 16 and 16 -> 32 as  X21 = -X22 (L21 X11)  on DMMA, the
// intermediate T going through `sc` (>= 288 doubles of scratch) to change from the accumulator to the operand layout.
// Replaces a 32-step substitution with lane = column (496 dependent FMA + broadcast LDS per lane, 4.2 us): the last
// block's inverse has nothing to hide behind.
__device__ __forceinline__ void inv32_warp(const double* __restrict__ Lb, const double* __restrict__ rd,
                                           double* __restrict__ Dinv, double* __restrict__ sc, int lane) {
  const int lr = lane >> 2, lq = lane & 3;
  constexpr int TS = 10, TS2 = 18;
#pragma unroll
  for (int e = lane; e < 6 * 64; e += 32) {  // the six 8 x 8 blocks above the block diagonal
    const int blk = e >> 6, r = (e >> 3) & 7, c = e & 7;
    const int bi = blk < 3 ? 0 : (blk < 5 ? 1 : 2);
    const int bj = blk < 3 ? blk + 1 : (blk < 5 ? blk - 1 : 3);
    Dinv[(bi * 8 + r) * DIS + bj * 8 + c] = 0.0;
  }
  {
    const int b = lane >> 3, j = lane & 7;
    const double* Lbb = Lb + (b * 8) * L2S + b * 8;
    double x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      double v = (r == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < r) v -= Lbb[r * L2S + k] * x[k];
      x[r] = (r >= j) ? v * rd[b * 8 + r] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) Dinv[(b * 8 + r) * DIS + b * 8 + j] = x[r];
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // 8 -> 16: T = L21 X11 for the pairs at p = 0 and p = 16
    const int p = h * 16;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
      dmma884(c0, c1, Lb[(p + 8 + lr) * L2S + p + lq + 4 * s2], Dinv[(p + lq + 4 * s2) * DIS + p + lr]);
    sc[h * 8 * TS + lr * TS + 2 * lq] = c0;
    sc[h * 8 * TS + lr * TS + 2 * lq + 1] = c1;
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // X21 = -X22 T
    const int p = h * 16;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
      dmma884(c0, c1, Dinv[(p + 8 + lr) * DIS + p + 8 + lq + 4 * s2], sc[h * 8 * TS + (lq + 4 * s2) * TS + lr]);
    Dinv[(p + 8 + lr) * DIS + p + 2 * lq] = -c0;
    Dinv[(p + 8 + lr) * DIS + p + 2 * lq + 1] = -c1;
  }
  __syncwarp();
  double t[2][2][2];
#pragma unroll
  for (int ti = 0; ti < 2; ++ti)  // 16 -> 32: T = L21 X11 (16 x 16 x 16)
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) {
      t[ti][tj][0] = t[ti][tj][1] = 0.0;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4)
        dmma884(t[ti][tj][0], t[ti][tj][1], Lb[(16 + ti * 8 + lr) * L2S + lq + 4 * s4],
                Dinv[(lq + 4 * s4) * DIS + tj * 8 + lr]);
      sc[(ti * 8 + lr) * TS2 + tj * 8 + 2 * lq] = t[ti][tj][0];
      sc[(ti * 8 + lr) * TS2 + tj * 8 + 2 * lq + 1] = t[ti][tj][1];
    }
  __syncwarp();
#pragma unroll
  for (int ti = 0; ti < 2; ++ti)  // X21 = -X22 T
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4)
        dmma884(c0, c1, Dinv[(16 + ti * 8 + lr) * DIS + 16 + lq + 4 * s4], sc[(lq + 4 * s4) * TS2 + tj * 8 + lr]);
      Dinv[(16 + ti * 8 + lr) * DIS + tj * 8 + 2 * lq] = -c0;
      Dinv[(16 + ti * 8 + lr) * DIS + tj * 8 + 2 * lq + 1] = -c1;
    }
  __syncwarp();
}

// FAST (chosen by the host: a full 128 x 128 leaf, 16-byte aligned rows, packed inverse wanted and merged here):
// bulk-TMA load and store only -- the element-wise loops and the merge-less exit are not even compiled in; the leaf
// is straight-line unrolled code executed once per 32-column block and its instruction footprint is part of its cost
// (carrying a second eliminator in the kernel slowed the first one by 6 us).
template <bool FAST>
__global__ void __launch_bounds__(LEAF_NT, 1)
potrf_leaf_dmma_kernel(double* __restrict__ A, long long lda, int n, double* __restrict__ inv, int merge) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                       // [128][L2S]  A -> L -> inverse, in place
  double* DI = S + LB * L2S;            // [4][32][DIS] inverses of the diagonal blocks
  double* T = DI + 4 * 32 * DIS;        // [64][T2S]
  double* rds = T + 64 * T2S;           // [128] reciprocals of the pivots, per diagonal block
  // the warp index through a lane-0 shuffle: provably warp-uniform for the compiler, so the role branches below are
  // uniform branches and the eliminator's shuffles compile to bare SHFL (with `tid >> 5` every __shfl_sync inside
  // `if (warp == 0)` was wrapped in WARPSYNC.COLLECTIVE ... ENDCOLLECTIVE plus register moves: 14 400 instructions)
  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int lr = lane >> 2, lq = lane & 3;

  LEAF_STAMP(0);
  CHAIN_BEGIN(0)
  // A full, aligned leaf comes in by bulk TMA: row i up to the end of its diagonal 32 x 32 block (one 256..1024-byte
  // copy per row, all 128 in flight at once; the element-wise loop below costs ~20 us of exposed L2 latency on the
  // panel's critical chain), zeros to the right of it.  What lands above the diagonal inside the diagonal blocks is
  // never read.
  __shared__ __align__(8) uint64_t lbar;
  constexpr bool bulk = FAST;
  if (bulk && tid == 0) {
    mbar_init(&lbar, 1);
    fence_mbar_init();
  }
  // no-ops unless launched as a programmatic dependent: wait for the head update, then let the head solve that
  // follows become resident (its four CTAs wait for this grid themselves)
  GPXB_PDL_WAIT_THEN_RELEASE();
  if (bulk) {
    __syncthreads();
    if (tid == 0) mbar_arrive_expect_tx(&lbar, 32 * (256 + 512 + 768 + 1024));
    if (tid < LB) tma_bulk_g2s(S + tid * L2S, A + (long long)tid * lda, ((tid >> 5) + 1) * 256, &lbar);
    for (int idx = tid; idx < 96 * 96; idx += LEAF_NT) {  // rows 0..95, columns beyond the row's diagonal block
      const int i = idx / 96, k = ((i >> 5) + 1) * 32 + (idx - i * 96);
      if (k < LB) S[i * L2S + k] = 0.0;
    }
    mbar_wait(&lbar, 0);
  } else {
    // lower triangle in, identity padding beyond n (keeps every block operation well defined); eight independent
    // loads per thread in flight
#pragma unroll 8
    for (int idx = tid; idx < LB * LB; idx += LEAF_NT) {
      const int i = idx >> 7, k = idx & (LB - 1);
      double v = 0.0;
      if (i < n) { if (k <= i) v = A[(long long)i * lda + k]; }
      else if (k == i) v = 1.0;
      S[i * L2S + k] = v;
    }
  }
  __syncthreads();
  LEAF_STAMP(1);

  // Roles inside the CTA.  The column-by-column elimination of a 32 x 32 diagonal block is a latency chain that one
  // warp must walk alone (warp 0, the eliminator); everything else is arranged so that the chain waits for as little
  // as possible:
  //   warp 7 (inverter): the inverse of the block just eliminated -- needed only for the packed inverse at the end;
  //   warps 1..6 (workers): the panel below it (forward substitution, one thread per row), then the trailing update
  //     on DMMA -- FIRST the ten 8 x 8 tiles of the NEXT diagonal block, after which the eliminator is released
  //     (barrier 2) while the rest of the trailing update continues.
  // Barrier 0 (__syncthreads): the eliminated block and its pivot reciprocals are visible, and all work of the
  // previous block has ended.  Barrier 1 (workers, 192 threads): panel complete.  Barrier 2 (eliminator + workers,
  // 224 threads): next diagonal block up to date.  The pivot reciprocals are kept per block: the inverter may still
  // read block jb's while the eliminator writes block jb + 32's.
  for (int jb = 0; jb < LB; jb += 32) {
    double* Dinv = DI + (jb >> 5) * 32 * DIS;
    double* rd_blk = rds + jb;
    if (warp == 0) {
      // ---- 32x32 diagonal block: lane = row; right-looking elimination through shuffles ----
      double a[32];
      const double* row = S + (jb + lane) * L2S + jb;
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? row[c] : 0.0;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double piv = __shfl_sync(0xffffffffu, a[c], c);
        const double rd = rsqrt(piv);  // NaN for a non-PD pivot: propagates like JAX
        if (lane == c) {
          a[c] = piv * rd;
          rd_blk[c] = rd;
        } else if (lane > c) {
          a[c] *= rd;
        }
        // multipliers by shuffle: a shared-memory broadcast (STS + __syncwarp + LDS) was measured 12 % SLOWER on the
        // whole factorisation -- its round trip sits on the column-to-column dependency chain
#pragma unroll
        for (int c2 = 0; c2 < 32; ++c2) {  // rectangular bounds: the triangular condition folds after unrolling
          if (c2 > c) {
            const double l2 = __shfl_sync(0xffffffffu, a[c], c2);  // L[c2][c]
            if (lane >= c2) a[c2] -= a[c] * l2;
          }
        }
      }
      double* rowo = S + (jb + lane) * L2S + jb;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c <= lane) rowo[c] = a[c];
    }
    __syncthreads();  // barrier 0
    LEAF_STAMP(4 + (jb >> 5) * 6);
    const int r0 = jb + 32, R = LB - r0;  // rows below the block
    if (warp == 7) {
      // ---- inverse of the diagonal block (needed only for the packed inverse at the end) ----
      inv32_warp(S + jb * L2S + jb, rd_blk, Dinv, T, lane);
    } else if (R > 0) {
      if (warp != 0) {
        // ---- panel: L21 = A21 L11^-T by forward substitution, one thread per row (warps 1..3) ----
        const int pr = tid - 32;
        if (pr < R) {
          double* arow = S + (r0 + pr) * L2S + jb;
          const double* Lb = S + jb * L2S + jb;
          double xr[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            double v[4] = {arow[c], 0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < c) v[k & 3] -= xr[k] * Lb[c * L2S + k];
            xr[c] = ((v[0] + v[1]) + (v[2] + v[3])) * rd_blk[c];
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) arow[c] = xr[c];
        }
        asm volatile("bar.sync 1, 192;" ::: "memory");  // workers: the panel is complete
      }
      // ---- trailing update A22 -= L21 L21^T on the lower 8 x 8 tiles: tiles 0..9 are the next diagonal block ----
      const int nt = R >> 3, ntiles = nt * (nt + 1) / 2;
      auto tile = [&](int t) {
        int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while (ti * (ti + 1) / 2 > t) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        const int tj = t - ti * (ti + 1) / 2;
        const double* ap = S + (r0 + ti * 8 + lr) * L2S + jb + lq;
        const double* bp = S + (r0 + tj * 8 + lr) * L2S + jb + lq;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) dmma884(c0, c1, ap[ks * 4], bp[ks * 4]);
        double* cp = S + (r0 + ti * 8 + lr) * L2S + r0 + tj * 8 + 2 * lq;
        cp[0] -= c0;
        cp[1] -= c1;
      };
      if (warp != 0)
        for (int t = warp - 1; t < 10; t += 6) tile(t);
      asm volatile("bar.sync 2, 224;" ::: "memory");  // the next diagonal block is up to date: release the eliminator
      if (warp != 0)
        for (int t = 10 + warp - 1; t < ntiles; t += 6) tile(t);
    }
    LEAF_STAMP(7 + (jb >> 5) * 6);
  }
  __syncthreads();  // the last inverse (warp 7) is complete
  LEAF_STAMP(27);
#pragma unroll 8
  for (int idx = tid; idx < LB * LB; idx += LEAF_NT) {
    const int i = idx >> 7, k = idx & (LB - 1);
    if (i < n && k <= i) A[(long long)i * lda + k] = S[i * L2S + k];
  }
  LEAF_STAMP(28);
  if (!FAST && inv == nullptr) return;
  if (!FAST && !(merge & 0xff)) {  // the four diagonal-block inverses only (the rest of the block is written by potrf_leaf_merge_kernel)
    for (int idx = tid; idx < 4 * 32 * 32; idx += LEAF_NT) {
      const int b = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
      inv[(b * 32 + r) * LB + b * 32 + c] = DI[(b * 32 + r) * DIS + c];
    }
    return;
  }
  __syncthreads();
  // ---- inverse in place: diagonal blocks, then the merge levels ----
  for (int idx = tid; idx < 4 * 32 * 32; idx += LEAF_NT) {
    const int b = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
    S[(b * 32 + r) * L2S + b * 32 + c] = DI[(b * 32 + r) * DIS + c];
  }
  __syncthreads();
  LEAF_STAMP(29);
  leaf_merge_and_store<FAST>(S, T, n, inv, tid, warp, lane, false);
  CHAIN_END();
  LEAF_STAMP(31);
}

// Completes the packed inverse of a leaf factorised with merge == 0: L from A (lower triangle, identity padding beyond
// n), the diagonal-block inverses from inv, merge levels, store.  One CTA, beside the panel's dependency chain.
__global__ void __launch_bounds__(LEAF_NT, 1)
potrf_leaf_merge_kernel(const double* __restrict__ A, long long lda, int n, double* __restrict__ inv) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;
  double* T = S + LB * L2S + 4 * 32 * DIS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 8
  for (int idx = tid; idx < LB * LB; idx += LEAF_NT) {
    const int i = idx >> 7, k = idx & (LB - 1);
    double v = 0.0;
    if ((i >> 5) == (k >> 5)) v = inv[idx];                      // diagonal 32 x 32 blocks: their inverses
    else if (i < n && k < i) v = A[(long long)i * lda + k];      // below them: L
    S[i * L2S + k] = v;
  }
  __syncthreads();
  leaf_merge_and_store<false>(S, T, n, inv, tid, warp, lane, true);
}

__global__ void logdet_kernel(const double* __restrict__ L, long long ld, int n,
                              double* __restrict__ out, double scale) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += log(L[(long long)i * ld + i]);
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = scale * s;
}

int leaf_bulk_io() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GPXB_LEAF_IO");  // "ldg": the element-wise load / store loops kept for comparison
    on = (e && e[0] == 'l') ? 0 : 1;
  }
  return on;
}

int leaf_generation() {
  static int gen = -1;
  if (gen < 0) {
    const char* e = getenv("GPXB_LEAF");  // "v1": the register-panel leaf kept for comparison
    gen = (e && e[0] == 'v' && e[1] == '1') ? 1 : 2;
  }
  return gen;
}

// merge == 0 (second-generation leaf only): inv receives the four 32 x 32 diagonal-block inverses; leaf_merge
// completes it later
int leaf(Ctx* ctx, double* A, long long lda, int n, double* inv, cudaStream_t s, int merge = 1, bool pdl = false) {
  const int gen = leaf_generation();
  if (gen == 2) {
    const bool fast = leaf_bulk_io() && n == LB && (lda & 1) == 0 && inv != nullptr && merge == 1 &&
                      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(inv)) & 15) == 0;
    auto kern = fast ? potrf_leaf_dmma_kernel<true> : potrf_leaf_dmma_kernel<false>;
    GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)kern, LEAF2_SMEM));
    if (pdl) GPXB_CUDA_CHECK(launch_pdl(kern, dim3(1), dim3(LEAF_NT), LEAF2_SMEM, s, A, lda, n, inv, merge));
    else kern<<<1, LEAF_NT, LEAF2_SMEM, s>>>(A, lda, n, inv, merge);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
    return 0;
  }
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)(potrf_leaf_kernel), LEAF_SMEM));
  potrf_leaf_kernel<<<1, LEAF_NT, LEAF_SMEM, s>>>(A, lda, n, inv);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

int leaf_merge(Ctx* ctx, const double* A, long long lda, int n, double* inv, cudaStream_t s) {
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)(potrf_leaf_merge_kernel), LEAF2_SMEM));
  potrf_leaf_merge_kernel<<<1, LEAF_NT, LEAF2_SMEM, s>>>(A, lda, n, inv);
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
    const int warp = uniform_warp_id(), lane = tid & 31;
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
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
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

// ---- few right-hand sides, second generation: ONE launch per 128-column step ---------------------------------
// The first generation ran [diagonal solve, 1 CTA] -> [update of the remaining right-hand side, many CTAs] as two
// dependent launches per step, and the diagonal kernel read its 128 x 128 inverse one row at a time (36 us per
// step by ncu, barrier / long-scoreboard bound).  Here CTA 0 of the update launch first updates the NEXT block's
// 128 entries, then applies that block's inverse to them, while the other CTAs update the rest: one launch per
// step on the dependency chain, and every thread issues its loads before it consumes them.

// so[r][t] = sum_{c <= t} inv[t][c] sb[r][c]: one warp per output row, four rows in flight (16 loads per lane)
__device__ __forceinline__ void trsv_diag_fwd_apply(const double* __restrict__ inv, double (*sb)[LB], double (*so)[LB],
                                                    int nb, int tid, int m) {
  const int warp = uniform_warp_id(), lane = tid & 31;
  for (int t0 = warp; t0 < nb; t0 += 32) {
    double v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = t0 + 8 * i, c = lane + 32 * q;
        v[i][q] = (t < nb && c <= t) ? inv[t * LB + c] : 0.0;
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + 8 * i;
      double acc[TRSV_MAXM];
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) {
        acc[r] = 0.0;
        if (r < m) {  // m is uniform: no reduction for absent right-hand sides
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r] += v[i][q] * sb[r][lane + 32 * q];
          acc[r] = warp_sum(acc[r]);
        }
      }
      if (lane == 0 && t < nb) {
#pragma unroll
        for (int r = 0; r < TRSV_MAXM; ++r) so[r][t] = acc[r];
      }
    }
  }
}
// so[r][t] = sum_{c >= t} inv[c][t] sb[r][c]: thread (t, half) sums half of the c range, halves combined through
// shared memory in a fixed order
__device__ __forceinline__ void trsv_diag_bwd_apply(const double* __restrict__ inv, double (*sb)[LB], double (*so)[LB],
                                                    double (*tmp)[LB], int nb, int tid) {
  const int t = tid & (LB - 1), half = tid >> 7;
  double acc[TRSV_MAXM];
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
  const int cb = half * (LB / 2), ce = min(nb, cb + LB / 2);
  if (t < nb) {
#pragma unroll 8
    for (int c = cb; c < ce; ++c) {
      const double v = (c >= t) ? inv[c * LB + t] : 0.0;
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sb[r][c];
    }
  }
  if (half == 1) {
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) tmp[r][t] = acc[r];
  }
  __syncthreads();
  if (half == 0 && t < nb) {
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) so[r][t] = acc[r] + tmp[r][t];
  }
}

// forward step: x = X[:, 0:nb) solved.  CTA 0: next block (rows [0, nb_next) below) updated, then solved against
// inv_next; CTAs >= 1: B[:, j] -= sum_c L[j][c] x[c] for the rows j >= nb_next below (one warp per row).
__global__ void __launch_bounds__(256) trsv_fwd_step_kernel(const double* __restrict__ Lblk, long long ldl,
                                                            const double* __restrict__ X, double* __restrict__ Bb,
                                                            long long ldb, int m, int nb, int nrows,
                                                            const double* __restrict__ inv_next, int nb_next) {
  __shared__ double sx[TRSV_MAXM][LB];
  __shared__ double sb[TRSV_MAXM][LB];
  __shared__ double so[TRSV_MAXM][LB];
  const int tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
  GPXB_PDL_WAIT_THEN_RELEASE();  // the sweep is a chain of short dependent launches (trsv_few_v2)
  for (int i = tid; i < TRSV_MAXM * LB; i += 256) {
    const int r = i / LB, c = i % LB;
    sx[r][c] = (r < m && c < nb) ? X[(long long)r * ldb + c] : 0.0;
  }
  __syncthreads();
  auto row_dot = [&](int j, double* acc) {  // acc[r] = sum_c L[j][c] x[r][c], reduced over the warp
    const double* Lr = Lblk + (long long)j * ldl;
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (lane + 32 * q < nb) ? Lr[lane + 32 * q] : 0.0;
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) {
      acc[r] = 0.0;
      if (r < m) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r] += v[q] * sx[r][lane + 32 * q];
        acc[r] = warp_sum(acc[r]);
      }
    }
  };
  if (blockIdx.x == 0) {
    for (int i = tid; i < TRSV_MAXM * LB; i += 256) {
      const int r = i / LB, c = i % LB;
      sb[r][c] = (r < m && c < nb_next) ? Bb[(long long)r * ldb + c] : 0.0;
    }
    __syncthreads();
    for (int j0 = warp; j0 < nb_next; j0 += 32) {  // four rows per pass: their loads are independent
      double acc[4][TRSV_MAXM];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (j0 + 8 * i < nb_next) row_dot(j0 + 8 * i, acc[i]);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (j0 + 8 * i < nb_next) {
#pragma unroll
            for (int r = 0; r < TRSV_MAXM; ++r) sb[r][j0 + 8 * i] -= acc[i][r];
          }
      }
    }
    __syncthreads();
    trsv_diag_fwd_apply(inv_next, sb, so, nb_next, tid, m);
    __syncthreads();
    for (int i = tid; i < m * nb_next; i += 256) {
      const int r = i / nb_next, c = i % nb_next;
      Bb[(long long)r * ldb + c] = so[r][c];
    }
    return;
  }
  const int j = nb_next + (blockIdx.x - 1) * 8 + warp;
  if (j >= nrows) return;
  double acc[TRSV_MAXM];
  row_dot(j, acc);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r)
      if (r < m) Bb[(long long)r * ldb + j] -= acc[r];
  }
}

// backward step: x = X[:, 0:nb) solved (block at column c0).  CTA 0: the previous block (columns [c0 - 128, c0))
// updated, then solved against inv_prev^T; CTAs >= 1: B[:, j] -= sum_c L[c0 + c][j] x[c] for the columns j < c0 - 128.
__global__ void __launch_bounds__(256) trsv_bwd_step_kernel(const double* __restrict__ Lrows, long long ldl,
                                                            const double* __restrict__ X, double* __restrict__ B,
                                                            long long ldb, int m, int nb, int c0,
                                                            const double* __restrict__ inv_prev) {
  __shared__ double sx[TRSV_MAXM][LB];
  __shared__ double sb[TRSV_MAXM][LB];
  __shared__ double so[TRSV_MAXM][LB];
  __shared__ double tmp[TRSV_MAXM][LB];
  const int tid = threadIdx.x;
  GPXB_PDL_WAIT_THEN_RELEASE();
  for (int i = tid; i < TRSV_MAXM * LB; i += 256) {
    const int r = i / LB, c = i % LB;
    sx[r][c] = (r < m && c < nb) ? X[(long long)r * ldb + c] : 0.0;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    const int t = tid & (LB - 1), half = tid >> 7, j = c0 - LB + t;
    double acc[TRSV_MAXM];
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
    const int cb = half * (LB / 2), ce = min(nb, cb + LB / 2);
#pragma unroll 8
    for (int c = cb; c < ce; ++c) {
      const double v = Lrows[(long long)c * ldl + j];
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sx[r][c];
    }
    if (half == 1) {
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) tmp[r][t] = acc[r];
    }
    __syncthreads();
    if (half == 0) {
#pragma unroll
      for (int r = 0; r < TRSV_MAXM; ++r) sb[r][t] = (r < m) ? B[(long long)r * ldb + j] - (acc[r] + tmp[r][t]) : 0.0;
    }
    __syncthreads();
    trsv_diag_bwd_apply(inv_prev, sb, so, tmp, LB, tid);
    __syncthreads();
    for (int i = tid; i < m * LB; i += 256) {
      const int r = i / LB, c = i % LB;
      B[(long long)r * ldb + c0 - LB + c] = so[r][c];
    }
    return;
  }
  const int j = (blockIdx.x - 1) * 256 + tid;
  if (j >= c0 - LB) return;
  double acc[TRSV_MAXM];
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r) acc[r] = 0.0;
#pragma unroll 8
  for (int c = 0; c < nb; ++c) {
    const double v = Lrows[(long long)c * ldl + j];
#pragma unroll
    for (int r = 0; r < TRSV_MAXM; ++r) acc[r] += v * sx[r][c];
  }
#pragma unroll
  for (int r = 0; r < TRSV_MAXM; ++r)
    if (r < m) B[(long long)r * ldb + j] -= acc[r];
}

// the first block of a sweep (nothing to update yet)
__global__ void __launch_bounds__(256) trsv_first_kernel(const double* __restrict__ inv, double* __restrict__ B,
                                                         long long ldb, int m, int nb, int backward) {
  __shared__ double sb[TRSV_MAXM][LB];
  __shared__ double so[TRSV_MAXM][LB];
  __shared__ double tmp[TRSV_MAXM][LB];
  const int tid = threadIdx.x;
  GPXB_PDL_WAIT_THEN_RELEASE();
  for (int i = tid; i < TRSV_MAXM * LB; i += 256) {
    const int r = i / LB, c = i % LB;
    sb[r][c] = (r < m && c < nb) ? B[(long long)r * ldb + c] : 0.0;
  }
  __syncthreads();
  if (backward) trsv_diag_bwd_apply(inv, sb, so, tmp, nb, tid);
  else trsv_diag_fwd_apply(inv, sb, so, nb, tid, m);
  __syncthreads();
  for (int i = tid; i < m * nb; i += 256) {
    const int r = i / nb, c = i % nb;
    B[(long long)r * ldb + c] = so[r][c];
  }
}

int trsv_few_v2(Ctx* ctx, const double* L, long long ldl, const double* inv, double* B, long long ldb, int m, int n,
                bool transposed_solve, cudaStream_t s) {
  const int nblk = (n + LB - 1) / LB;
  if (transposed_solve) {  // L x = b: forward over blocks
    trsv_first_kernel<<<1, 256, 0, s>>>(inv, B, ldb, m, min(LB, n), 0);
    ctx->launches++;
    for (int k = 0; k + 1 < nblk; ++k) {
      const int c0 = k * LB, below = n - c0 - LB, nb_next = min(LB, below);
      GPXB_CUDA_CHECK(launch_pdl(trsv_fwd_step_kernel, dim3(1 + (below - nb_next + 7) / 8), dim3(256), 0, s,
                                 L + (long long)(c0 + LB) * ldl + c0, ldl, B + c0, B + c0 + LB, ldb, m, LB, below,
                                 inv + (long long)(k + 1) * LB * LB, nb_next));
      ctx->launches++;
    }
  } else {  // L^T x = b: backward over blocks
    const int cl = (nblk - 1) * LB;
    trsv_first_kernel<<<1, 256, 0, s>>>(inv + (long long)(nblk - 1) * LB * LB, B + cl, ldb, m, n - cl, 1);
    ctx->launches++;
    for (int k = nblk - 1; k >= 1; --k) {
      const int c0 = k * LB, nb = min(LB, n - c0);
      GPXB_CUDA_CHECK(launch_pdl(trsv_bwd_step_kernel, dim3(1 + (c0 - LB + 255) / 256), dim3(256), 0, s,
                                 L + (long long)c0 * ldl, ldl, B + c0, B, ldb, m, nb, c0,
                                 inv + (long long)(k - 1) * LB * LB));
      ctx->launches++;
    }
  }
  GPXB_LAUNCH_CHECK();
  return 0;
}

// few right-hand sides (m <= 4): blocked substitution with GEMV updates -- HBM bound (reads L once)
int trsv_few(Ctx* ctx, const double* L, long long ldl, const double* inv, double* B, long long ldb, int m, int n,
             bool transposed_solve, cudaStream_t s) {
  static const bool v1 = [] {
    const char* e = getenv("GPXB_TRSV");  // "v1": two launches per step (kept for comparison)
    return e && e[0] == 'v' && e[1] == '1';
  }();
  if (!v1) return trsv_few_v2(ctx, L, ldl, inv, B, ldb, m, n, transposed_solve, s);
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
// ---- small-tile kernels of the panel's dependency chain -----------------------------------------------------
// Inside a panel the chain  leaf(j) -> rows of block j+1 solved -> diagonal block j+1 updated -> leaf(j+1)  is what
// bounds a mid-size factorisation (n <= 8192): run through the 128 x 128-tile GEMM each link costs ~25 us whatever
// its size (one tile = 4.2 MFLOP on ONE SM at 250 GFLOP/s).  These two kernels cut the links into 32-row pieces
// spread over several SMs; the bulk of the solve / update runs beside them on the other stream.
constexpr int PSL = LB + 4;  // shared-memory row stride (== 4 mod 8: conflict-free DMMA fragment reads)
constexpr int TRSM32_SMEM = (32 + LB) * PSL * (int)sizeof(double);
constexpr int SYRK32_SMEM = 2 * 32 * PSL * (int)sizeof(double);

// B[r0 : r0 + 32, 0 : jb) <- B[same] inv^T in place (C[i][n] = sum_k B[i][k] inv[n][k]); inv: packed 128 x 128 leaf
// inverse (zero above the diagonal and beyond jb).  One CTA per 32 rows, 8 warps x (32 rows x 16 columns).
__global__ void __launch_bounds__(256) panel_trsm32_kernel(double* __restrict__ B, long long ldb, int rows, int jb,
                                                           const double* __restrict__ inv, int bulk_io) {
  extern __shared__ __align__(16) double psm[];
  double* sB = psm;               // [32][PSL]
  double* sI = psm + 32 * PSL;    // [128][PSL]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lq = lane & 3;
  const int r0 = blockIdx.x * 32;
  CHAIN_BEGIN(1)
  // programmatic dependent launch: the head update that follows on the chain stream may be scheduled now (it waits
  // for this grid's completion itself, griddepcontrol.wait) -- hides its launch latency behind this kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the leaf, when this grid is its programmatic dependent
  // Full, aligned pieces come in by bulk TMA -- 160 row copies in flight at once (inv: row i up to the end of its
  // 16-column group, all the product reads); the element-wise loops cost several exposed L2 round trips on the
  // panel's critical chain.
  __shared__ __align__(8) uint64_t bar;
  const bool bulk = bulk_io && jb == LB && r0 + 32 <= rows && (ldb & 1) == 0 &&
                    ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(inv)) & 15) == 0;
  if (bulk) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) mbar_arrive_expect_tx(&bar, 32 * 1024 + 16 * 128 * 36);
    if (tid < LB) tma_bulk_g2s(sI + tid * PSL, inv + tid * LB, ((tid >> 4) + 1) * 128, &bar);
    else if (tid < LB + 32) tma_bulk_g2s(sB + (tid - LB) * PSL, B + (long long)(r0 + tid - LB) * ldb, 1024, &bar);
    mbar_wait(&bar, 0);
  } else {
#pragma unroll 8
    for (int idx = tid; idx < LB * LB / 2; idx += 256) {  // inv: 16-byte loads (the packed block is 16-byte aligned)
      const int i = idx >> 6, k2 = (idx & 63) * 2;
      const double2 v = *reinterpret_cast<const double2*>(inv + i * LB + k2);
      sI[i * PSL + k2] = v.x;
      sI[i * PSL + k2 + 1] = v.y;
    }
#pragma unroll 4
    for (int idx = tid; idx < 32 * LB; idx += 256) {
      const int i = idx >> 7, k = idx & (LB - 1);
      sB[i * PSL + k] = (r0 + i < rows && k < jb) ? B[(long long)(r0 + i) * ldb + k] : 0.0;
    }
    __syncthreads();
  }
  double c[4][2][2];
#pragma unroll
  for (int mf = 0; mf < 4; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) c[mf][nf][0] = c[mf][nf][1] = 0.0;
  const double* ap = sB + lr * PSL + lq;
  const double* bp = sI + (warp * 16 + lr) * PSL + lq;
  const int kend = min((jb + 3) & ~3, warp * 16 + 16);  // inv is lower triangular: inv[n][k] = 0 for k > n
  for (int k = 0; k < kend; k += 4) {
    double a[4], b[2];
#pragma unroll
    for (int mf = 0; mf < 4; ++mf) a[mf] = ap[mf * 8 * PSL + k];
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) b[nf] = bp[nf * 8 * PSL + k];
#pragma unroll
    for (int mf = 0; mf < 4; ++mf)
#pragma unroll
      for (int nf = 0; nf < 2; ++nf) dmma884(c[mf][nf][0], c[mf][nf][1], a[mf], b[nf]);
  }
#pragma unroll
  for (int mf = 0; mf < 4; ++mf) {
    const int row = r0 + mf * 8 + lr;
    if (row >= rows) continue;
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      const int col = warp * 16 + nf * 8 + 2 * lq;
      if (col < jb) B[(long long)row * ldb + col] = c[mf][nf][0];
      if (col + 1 < jb) B[(long long)row * ldb + col + 1] = c[mf][nf][1];
    }
  }
  CHAIN_END();
}

// The same solve from L itself and the four 32 x 32 diagonal-block inverses (the leaf's merge-less mode): blocked
// forward substitution over the four column blocks,  X_q = (B_q - sum_{p<q} X_p L_qp^T) D_q^-T,  both products on
// DMMA from shared memory.  Lj: the factorised diagonal block in A (ld ldl); dinv: inv with (at least) its diagonal
// 32 x 32 blocks valid.  One CTA per 32 rows.
constexpr int TRSM32S_SMEM = ((LB + 32) * PSL + 4 * 32 * 36 + 32 * 36) * (int)sizeof(double);

__global__ void __launch_bounds__(256) panel_trsm32s_kernel(double* __restrict__ B, long long ldb, int rows, int jb,
                                                            const double* __restrict__ Lj, long long ldl,
                                                            const double* __restrict__ dinv) {
  extern __shared__ __align__(16) double psm[];
  double* sL = psm;                    // [128][PSL] lower triangle of L (identity beyond jb)
  double* sX = sL + LB * PSL;          // [32][PSL]  B rows -> X in place
  double* sD = sX + 32 * PSL;          // [4][32][36] diagonal-block inverses
  double* sR = sD + 4 * 32 * 36;       // [32][36]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lq = lane & 3;
  const int r0 = blockIdx.x * 32;
#pragma unroll 8
  for (int idx = tid; idx < LB * LB; idx += 256) {
    const int i = idx >> 7, k = idx & (LB - 1);
    double v = 0.0;
    if (i < jb) { if (k < i) v = Lj[(long long)i * ldl + k]; }
    else if (k == i) v = 1.0;
    sL[i * PSL + k] = v;
  }
  for (int idx = tid; idx < 4 * 32 * 32; idx += 256) {
    const int b = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
    sD[(b * 32 + r) * 36 + c] = dinv[(b * 32 + r) * LB + b * 32 + c];
  }
#pragma unroll 4
  for (int idx = tid; idx < 32 * LB; idx += 256) {
    const int i = idx >> 7, k = idx & (LB - 1);
    sX[i * PSL + k] = (r0 + i < rows && k < jb) ? B[(long long)(r0 + i) * ldb + k] : 0.0;
  }
  __syncthreads();
  const int mf = warp & 3, nf0 = (warp >> 2) * 2;  // this warp's two 8 x 8 fragments of a 32 x 32 block
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // R = B_q - X_{<q} L_{q,<q}^T
      const int nf = nf0 + h;
      const double* ap = sX + (mf * 8 + lr) * PSL + lq;
      const double* bp = sL + (q * 32 + nf * 8 + lr) * PSL + lq;  // B[k][n] = L[n][k]
      double c0 = 0.0, c1 = 0.0;
      for (int k = 0; k < 32 * q; k += 4) dmma884(c0, c1, ap[k], bp[k]);
      const double* xp = sX + (mf * 8 + lr) * PSL + q * 32 + nf * 8 + 2 * lq;
      double* rp = sR + (mf * 8 + lr) * 36 + nf * 8 + 2 * lq;
      rp[0] = xp[0] - c0;
      rp[1] = xp[1] - c1;
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // X_q = R D_q^-T
      const int nf = nf0 + h;
      const double* ap = sR + (mf * 8 + lr) * 36 + lq;
      const double* bp = sD + (q * 32 + nf * 8 + lr) * 36 + lq;   // B[k][n] = Dinv[n][k]
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int k = 0; k < 32; k += 4) dmma884(c0, c1, ap[k], bp[k]);
      double* xp = sX + (mf * 8 + lr) * PSL + q * 32 + nf * 8 + 2 * lq;
      xp[0] = c0;
      xp[1] = c1;
    }
    __syncthreads();
  }
#pragma unroll 4
  for (int idx = tid; idx < 32 * LB; idx += 256) {
    const int i = idx >> 7, k = idx & (LB - 1);
    if (r0 + i < rows && k < jb) B[(long long)(r0 + i) * ldb + k] = sX[i * PSL + k];
  }
}

// C[nbn x nbn] (lower) -= P P^T, P = [nbn x jb] row-major (ld ldp): one CTA per 32 x 32 tile of the lower
// triangle, 4 warps x (16 x 16).
__global__ void __launch_bounds__(128) panel_syrk32_kernel(const double* __restrict__ P, long long ldp,
                                                           double* __restrict__ C, long long ldc, int nbn, int jb,
                                                           int bulk_io) {
  extern __shared__ __align__(16) double psm[];
  double* sA = psm;              // rows of tile row ti   [32][PSL]
  double* sT = psm + 32 * PSL;   // rows of tile column tj
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lq = lane & 3;
  int ti = 0, t = blockIdx.x;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int r0 = ti * 32, c0 = tj * 32;
  CHAIN_BEGIN(2)
  // no-ops unless launched as a programmatic dependent: wait for the head solve, then let the next leaf become
  // resident (it claims its SM before the bulk update's CTAs arrive, and waits for this grid itself)
  GPXB_PDL_WAIT_THEN_RELEASE();
  __shared__ __align__(8) uint64_t bar;
  const bool bulk = bulk_io && jb == LB && r0 + 32 <= nbn && c0 + 32 <= nbn && (ldp & 1) == 0 &&
                    (reinterpret_cast<uintptr_t>(P) & 15) == 0;
  if (bulk) {  // 64 row copies of 1 KB in flight at once instead of 16 exposed L2 round trips
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) mbar_arrive_expect_tx(&bar, 64 * 1024);
    if (tid < 32) tma_bulk_g2s(sA + tid * PSL, P + (long long)(r0 + tid) * ldp, 1024, &bar);
    else if (tid < 64) tma_bulk_g2s(sT + (tid - 32) * PSL, P + (long long)(c0 + tid - 32) * ldp, 1024, &bar);
    mbar_wait(&bar, 0);
  } else {
#pragma unroll 4
    for (int idx = tid; idx < 32 * LB; idx += 128) {
      const int i = idx >> 7, k = idx & (LB - 1);
      sA[i * PSL + k] = (r0 + i < nbn && k < jb) ? P[(long long)(r0 + i) * ldp + k] : 0.0;
      sT[i * PSL + k] = (c0 + i < nbn && k < jb) ? P[(long long)(c0 + i) * ldp + k] : 0.0;
    }
    __syncthreads();
  }
  const int wm = warp >> 1, wn = warp & 1;
  double c[2][2][2];
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) c[mf][nf][0] = c[mf][nf][1] = 0.0;
  const double* ap = sA + (wm * 16 + lr) * PSL + lq;
  const double* bp = sT + (wn * 16 + lr) * PSL + lq;
  const int kend = (jb + 3) & ~3;
  for (int k = 0; k < kend; k += 4) {
    double a[2], b[2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf) a[mf] = ap[mf * 8 * PSL + k];
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) b[nf] = bp[nf * 8 * PSL + k];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
      for (int nf = 0; nf < 2; ++nf) dmma884(c[mf][nf][0], c[mf][nf][1], a[mf], b[nf]);
  }
#pragma unroll
  for (int mf = 0; mf < 2; ++mf) {
    const int row = r0 + wm * 16 + mf * 8 + lr;
    if (row >= nbn) continue;
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      const int col = c0 + wn * 16 + nf * 8 + 2 * lq;
      if (col <= row) C[(long long)row * ldc + col] -= c[mf][nf][0];
      if (col + 1 <= row) C[(long long)row * ldc + col + 1] -= c[mf][nf][1];
    }
  }
  CHAIN_END();
}

// Both links in ONE launch (the common case: the next block's columns are inside the panel): CTA (ti, tj), tj <= ti,
// solves the two 32-row pieces ti and tj of the head rows against inv^T itself (a 32 x 128 x 128 triangular product
// is 0.5 MFLOP: cheaper than a kernel boundary), keeps them in shared memory, subtracts their product from its
// 32 x 32 tile of the next diagonal block, and -- diagonal CTAs only -- writes its solved rows back.  One dependent
// launch less per 128 columns and no second trip of the head rows through L2.
constexpr int HEAD_SMEM = (LB + 64) * PSL * (int)sizeof(double);

// The head rows are solved IN PLACE while other CTAs of the same launch still have to read their unsolved values:
// every CTA announces the end of its loads on a ticket counter, and a diagonal CTA writes back only when all CTAs of
// the launch have announced (the launches are ordered on one stream, so `target` = CTAs of all launches so far; ten
// CTAs are always co-resident or become so as independent kernels drain).
__global__ void __launch_bounds__(256) panel_head_kernel(double* __restrict__ B, long long ldb, int nbn, int jb,
                                                         const double* __restrict__ inv, double* __restrict__ C,
                                                         long long ldc, unsigned long long* __restrict__ counter,
                                                         unsigned long long target, int bulk_io) {
  extern __shared__ __align__(16) double psm[];
  double* sI = psm;                   // [128][PSL]
  double* sBi = psm + LB * PSL;       // [32][PSL] rows of piece ti, then their solved values
  double* sBj = sBi + 32 * PSL;       // [32][PSL] rows of piece tj
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lq = lane & 3;
  int ti = 0, t = blockIdx.x;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const int tj = t - ti * (ti + 1) / 2;
  const bool diag = (ti == tj);
  const int ri = ti * 32, rj = tj * 32;
  __shared__ __align__(8) uint64_t bar;
  const bool bulk = bulk_io && jb == LB && ri + 32 <= nbn && rj + 32 <= nbn && (ldb & 1) == 0 &&
                    ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(inv)) & 15) == 0;
  if (bulk) {  // as in panel_trsm32_kernel: every row copy in flight at once
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) mbar_arrive_expect_tx(&bar, (diag ? 32 : 64) * 1024 + 16 * 128 * 36);
    if (tid < LB) tma_bulk_g2s(sI + tid * PSL, inv + tid * LB, ((tid >> 4) + 1) * 128, &bar);
    else if (tid < LB + 32) tma_bulk_g2s(sBi + (tid - LB) * PSL, B + (long long)(ri + tid - LB) * ldb, 1024, &bar);
    else if (tid < LB + 64 && !diag)
      tma_bulk_g2s(sBj + (tid - LB - 32) * PSL, B + (long long)(rj + tid - LB - 32) * ldb, 1024, &bar);
    mbar_wait(&bar, 0);
    __syncthreads();  // every thread's wait is over before the ticket below announces the reads
  } else {
#pragma unroll 8
    for (int idx = tid; idx < LB * LB / 2; idx += 256) {
      const int i = idx >> 6, k2 = (idx & 63) * 2;
      const double2 v = *reinterpret_cast<const double2*>(inv + i * LB + k2);
      sI[i * PSL + k2] = v.x;
      sI[i * PSL + k2 + 1] = v.y;
    }
#pragma unroll 4
    for (int idx = tid; idx < 32 * LB; idx += 256) {
      const int i = idx >> 7, k = idx & (LB - 1);
      sBi[i * PSL + k] = (ri + i < nbn && k < jb) ? B[(long long)(ri + i) * ldb + k] : 0.0;
      if (!diag) sBj[i * PSL + k] = (rj + i < nbn && k < jb) ? B[(long long)(rj + i) * ldb + k] : 0.0;
    }
    __syncthreads();
  }
  if (tid == 0) {  // this CTA holds its copies of the unsolved rows
    __threadfence();
    atomicAdd(counter, 1ULL);
  }
  // ---- the two solves: X = B inv^T, warp w owns columns 16 w .. 16 w + 15 ----
  double ci[4][2][2], cj[4][2][2];
#pragma unroll
  for (int mf = 0; mf < 4; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) ci[mf][nf][0] = ci[mf][nf][1] = cj[mf][nf][0] = cj[mf][nf][1] = 0.0;
  {
    const double* bp = sI + (warp * 16 + lr) * PSL + lq;
    const double* api = sBi + lr * PSL + lq;
    const double* apj = sBj + lr * PSL + lq;
    const int kend = min((jb + 3) & ~3, warp * 16 + 16);
    for (int k = 0; k < kend; k += 4) {
      double b[2];
#pragma unroll
      for (int nf = 0; nf < 2; ++nf) b[nf] = bp[nf * 8 * PSL + k];
#pragma unroll
      for (int mf = 0; mf < 4; ++mf) {
        const double a = api[mf * 8 * PSL + k];
#pragma unroll
        for (int nf = 0; nf < 2; ++nf) dmma884(ci[mf][nf][0], ci[mf][nf][1], a, b[nf]);
      }
      if (!diag) {
#pragma unroll
        for (int mf = 0; mf < 4; ++mf) {
          const double a = apj[mf * 8 * PSL + k];
#pragma unroll
          for (int nf = 0; nf < 2; ++nf) dmma884(cj[mf][nf][0], cj[mf][nf][1], a, b[nf]);
        }
      }
    }
  }
  if (diag && tid == 0) {  // in-place write-back only after every CTA of this launch has read (bounded wait)
    unsigned long long v;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
      if (++spins > (1LL << 31)) __trap();
    } while (v < target);
  }
  __syncthreads();  // every warp has read the unsolved rows (and, diagonal CTAs: so has every other CTA)
#pragma unroll
  for (int mf = 0; mf < 4; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      const int col = warp * 16 + nf * 8 + 2 * lq, r = mf * 8 + lr;
      sBi[r * PSL + col] = ci[mf][nf][0];
      sBi[r * PSL + col + 1] = ci[mf][nf][1];
      if (!diag) {
        sBj[r * PSL + col] = cj[mf][nf][0];
        sBj[r * PSL + col + 1] = cj[mf][nf][1];
      } else if (ri + r < nbn) {  // the diagonal CTA of a piece owns its write-back
        if (col < jb) B[(long long)(ri + r) * ldb + col] = ci[mf][nf][0];
        if (col + 1 < jb) B[(long long)(ri + r) * ldb + col + 1] = ci[mf][nf][1];
      }
    }
  __syncthreads();
  // ---- C[ti][tj] -= X_i X_j^T: 16 fragments of 8 x 8 over 8 warps ----
  const double* sXj = diag ? sBi : sBj;
  const int mf = warp & 3;
  const int kend = (jb + 3) & ~3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int nf = (warp >> 2) * 2 + h;
    const double* ap = sBi + (mf * 8 + lr) * PSL + lq;
    const double* bp = sXj + (nf * 8 + lr) * PSL + lq;
    double c0 = 0.0, c1 = 0.0;
    for (int k = 0; k < kend; k += 4) dmma884(c0, c1, ap[k], bp[k]);
    const int row = ri + mf * 8 + lr, col = rj + nf * 8 + 2 * lq;
    if (row < nbn) {
      if (col <= row) C[(long long)row * ldc + col] -= c0;
      if (col + 1 <= row) C[(long long)row * ldc + col + 1] -= c1;
    }
  }
}

// Joins a forked stream back into the caller's stream on EVERY exit path: if a launch fails mid-loop the callers'
// stream-ordered frees (DevBuf destructors on s) must still come after the work already queued on the side stream.
struct StreamJoin {
  Ctx* ctx;
  cudaStream_t from, into;
  ~StreamJoin() {
    if (!from || from == into) return;
    cudaEvent_t e;
    if (ctx->next_event(&e) != 0) return;
    if (cudaEventRecord(e, from) == cudaSuccess) cudaStreamWaitEvent(into, e, 0);
  }
};

// Factor the tall panel A[k0:n, k0:k0+w] (its w x w head is the diagonal block); everything this panel reads has been
// written on stream s.  First generation (GPXB_PANEL=v1): leaf -> tall TRSM -> rank-128 update, three dependent
// launches per 128 columns on one stream.
int factor_panel_v1(Ctx* ctx, double* A, long long lda, int n, int k0, int w, double* inv,
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

// Second generation (default): the chain  leaf(j) -> [rows of block j+1] inv^T -> diagonal block j+1 -= ...  runs on
// the high-priority stream ctx->crit through the 32-row kernels above, while the tall part of the solve and the
// rest of the rank-128 update (every 128-tile but the next diagonal block) run on s, one step behind:
//   crit:  leaf j | trsm32 head j | syrk32 head j | leaf j+1 | ...
//   s   :           TRSM rows below the head of j (after leaf j) | update j minus tile (0,0) (after head j) | ...
// crit waits for the bulk update of step j-1 before it touches block j+1 (that update wrote the head rows' entries
// in block column j+1 and the diagonal block itself).
int factor_panel(Ctx* ctx, double* A, long long lda, int n, int k0, int w, double* inv,
                 cudaStream_t s) {
  static const bool v1 = [] {
    const char* e = getenv("GPXB_PANEL");
    return e && e[0] == 'v' && e[1] == '1';
  }();
  if (v1 || ctx->serialize || !ctx->crit || w <= LB) return factor_panel_v1(ctx, A, lda, n, k0, w, inv, s);
  cudaStream_t cs = ctx->crit;
  StreamJoin join_crit{ctx, cs, s};
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)panel_trsm32_kernel, TRSM32_SMEM));
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)panel_syrk32_kernel, SYRK32_SMEM));
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)panel_head_kernel, HEAD_SMEM));
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)panel_trsm32s_kernel, TRSM32S_SMEM));
  // GPXB_LEAF_SPLIT=1: merge-less leaf + substitution head solve + merge kernel beside the chain.  Measured SLOWER
  // than the leaf that finishes its own packed inverse (potrf 2000 / 4096 / 8192: 1.69 / 3.76 / 11.0 ms against
  // 1.52 / 3.39 / 10.2 ms): the substitution solve pays eight block barriers and a predicated load of L, and the
  // merge launch delays the tall solve on the other stream.  Off by default, kept for the record.
  static const bool no_split_leaf = [] {
    const char* e = getenv("GPXB_LEAF_SPLIT");
    return !(e && e[0] == '1');
  }();
  // Default: head solve and head update as two launches.  GPXB_PANEL_HEAD=fused selects panel_head_kernel (both in
  // one launch, every CTA solving the pieces it multiplies): measured equal within noise on potrf 2000 / 4096 / 8192
  // (1.55 / 3.43 / 10.24 ms fused vs 1.52-1.55 / 3.39-3.46 / 10.17-10.29 ms split over two boxes) -- each of its ten CTAs stages the whole 128 x 128
  // inverse -- so the simpler pair stays the default.
  static const bool split_head = [] {
    const char* e = getenv("GPXB_PANEL_HEAD");
    return !(e && e[0] == 'f');
  }();
  // GPXB_PANEL_PDL: 0 plain stream order; 1 head update as a programmatic dependent of the head solve; 2 (default)
  // also the next leaf as a dependent of the head update -- the event the bulk update waits for is then recorded
  // right after the head solve (the update skips the diagonal block the head update writes).
  static const int pdl_mode = [] {
    const char* e = getenv("GPXB_PANEL_PDL");
    return e ? atoi(e) : 2;
  }();
  bool leaf_follows_update = false;  // the previous operation on cs is a head update launched in this loop
  cudaEvent_t ev, ev_bulk_prev = nullptr;
  GPXB_TRY(ctx->next_event(&ev));
  GPXB_CUDA_CHECK(cudaEventRecord(ev, s));
  GPXB_CUDA_CHECK(cudaStreamWaitEvent(cs, ev, 0));
  for (int j = 0; j < w; j += LB) {
    const int jb = min(LB, w - j);
    const int c0 = k0 + j;
    double* Ajj = A + (long long)c0 * lda + c0;
    double* invj = inv + (long long)(c0 / LB) * LB * LB;
    const int below = n - c0 - jb;
    // merge-less leaf when a head solve follows: the merge levels of the packed inverse run on s, beside the chain
    const bool split_leaf = below > 0 && jb == LB && leaf_generation() == 2 && !no_split_leaf;
    GPXB_TRY(leaf(ctx, Ajj, lda, jb, invj, cs, split_leaf ? 0 : 1, pdl_mode >= 2 && leaf_follows_update));
    leaf_follows_update = false;
    if (below <= 0) continue;
    cudaEvent_t ev_leaf, ev_head;
    GPXB_TRY(ctx->next_event(&ev_leaf));
    GPXB_CUDA_CHECK(cudaEventRecord(ev_leaf, cs));
    double* B = A + (long long)(c0 + jb) * lda + c0;  // below x jb
    const int wr = w - j - jb;                         // panel columns still to update
    const int nbn = min(LB, below);                    // rows of the next block (the "head")
    // ---- critical chain on cs ----
    bool head_recorded = false;
    if (ev_bulk_prev) GPXB_CUDA_CHECK(cudaStreamWaitEvent(cs, ev_bulk_prev, 0));
    if (wr >= nbn && !split_head && !split_leaf) {
      // both links in one launch: every CTA solves the head pieces it multiplies
      const int nt = ceil_div(nbn, 32), nctas = nt * (nt + 1) / 2;
      if (!ctx->head_counter) {
        GPXB_CUDA_CHECK(cudaMalloc(&ctx->head_counter, sizeof(unsigned long long)));
        GPXB_CUDA_CHECK(cudaMemsetAsync(ctx->head_counter, 0, sizeof(unsigned long long), cs));
        ctx->head_ticket = 0;
      }
      ctx->head_ticket += (unsigned long long)nctas;
      panel_head_kernel<<<nctas, 256, HEAD_SMEM, cs>>>(B, lda, nbn, jb, invj, B + jb, lda, ctx->head_counter,
                                                        ctx->head_ticket, leaf_bulk_io());
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
    } else {
      if (split_leaf) panel_trsm32s_kernel<<<ceil_div(nbn, 32), 256, TRSM32S_SMEM, cs>>>(B, lda, nbn, jb, Ajj, lda, invj);
      else if (pdl_mode >= 3)
        GPXB_CUDA_CHECK(launch_pdl(panel_trsm32_kernel, dim3(ceil_div(nbn, 32)), dim3(256), TRSM32_SMEM, cs, B,
                                   (long long)lda, nbn, jb, (const double*)invj, leaf_bulk_io()));
      else panel_trsm32_kernel<<<ceil_div(nbn, 32), 256, TRSM32_SMEM, cs>>>(B, lda, nbn, jb, invj, leaf_bulk_io());
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
      if (pdl_mode >= 2 && !split_leaf) {
        GPXB_TRY(ctx->next_event(&ev_head));
        GPXB_CUDA_CHECK(cudaEventRecord(ev_head, cs));
        head_recorded = true;
      }
      if (wr > 0) {
        const int nt = ceil_div(min(nbn, wr), 32);
        const bool pdl = pdl_mode >= 1;
        if (pdl && !split_leaf) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(nt * (nt + 1) / 2);
          cfg.blockDim = dim3(128);
          cfg.dynamicSmemBytes = SYRK32_SMEM;
          cfg.stream = cs;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at;
          cfg.numAttrs = 1;
          GPXB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, panel_syrk32_kernel, (const double*)B, (long long)lda, B + jb,
                                             (long long)lda, min(nbn, wr), jb, leaf_bulk_io()));
        } else {
          panel_syrk32_kernel<<<nt * (nt + 1) / 2, 128, SYRK32_SMEM, cs>>>(B, lda, B + jb, lda, min(nbn, wr), jb,
                                                                           leaf_bulk_io());
        }
        GPXB_LAUNCH_CHECK();
        ctx->launches++;
        leaf_follows_update = head_recorded;
      }
    }
    if (!head_recorded) {
      GPXB_TRY(ctx->next_event(&ev_head));
      GPXB_CUDA_CHECK(cudaEventRecord(ev_head, cs));
    }
    // ---- bulk on s ----
    GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_leaf, 0));
    if (split_leaf) GPXB_TRY(leaf_merge(ctx, Ajj, lda, jb, invj, s));  // the packed inverse, off the chain
    if (below > nbn) {
      // rows below the head: B <- B inv^T (in place).  Full 128-column steps go through the 32-row solve kernel as
      // well (one CTA per 32 rows at the DMMA rate of its SM, half the flops of a general product: 2.2 us per CTA
      // against 17-27 us for the latency of one GEMM tile at K = 128); GPXB_PANEL_TRSM=gemm keeps the GEMM.
      static const bool trsm_gemm = [] {
        const char* e = getenv("GPXB_PANEL_TRSM");
        return e && e[0] == 'g';
      }();
      if (jb == LB && !trsm_gemm && !split_leaf && below - nbn <= 4096) {  // beyond: equal (n = 8192: 9.45 vs 9.30 ms)
        panel_trsm32_kernel<<<ceil_div(below - nbn, 32), 256, TRSM32_SMEM, s>>>(B + (long long)nbn * lda, lda,
                                                                                  below - nbn, jb, invj, leaf_bulk_io());
        GPXB_LAUNCH_CHECK();
        ctx->launches++;
      } else {
        GemmArgs g;
        g.M = below - nbn; g.N = jb; g.K = jb;
        g.A = B + (long long)nbn * lda; g.lda = lda; g.a_kmajor = true;
        g.B = invj; g.ldb = LB; g.b_kmajor = true;
        g.C = B + (long long)nbn * lda; g.ldc = lda;
        g.small_tile_max = 35;  // <= 140 CTAs of 64 x 64: SMs stay free for the chain kernels
        GPXB_TRY(gemm_f64(ctx, g, s));
      }
    }
    GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_head, 0));
    if (wr > 0 && below > nbn) {
      GemmArgs g;  // C -= B B[0:wr]^T on the lower trapezoid, except the diagonal block the head kernels own
      g.M = below; g.N = wr; g.K = jb;
      g.A = B; g.lda = lda; g.a_kmajor = true;
      g.B = B; g.ldb = lda; g.b_kmajor = true;
      g.C = B + jb; g.ldc = lda;
      g.alpha = -1.0; g.beta = 1.0;
      g.lower_only = 1;
      g.skip_tile0 = 1;
      g.small_tile_max = 35;
      GPXB_TRY(gemm_f64(ctx, g, s));
    }
    GPXB_TRY(ctx->next_event(&ev_bulk_prev));
    GPXB_CUDA_CHECK(cudaEventRecord(ev_bulk_prev, s));
  }
  GPXB_TRY(ctx->next_event(&ev));
  GPXB_CUDA_CHECK(cudaEventRecord(ev, cs));
  GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
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
  static const int nb_env = [] {
    const char* e = getenv("GPXB_POTRF_NB");  // panel width (multiple of 128); default 1024
    const int v = e ? atoi(e) : 0;
    return (v >= LB && v % LB == 0) ? v : 0;
  }();
  // one panel up to n = 2048: the rank-1024 update between two panels stalls the leaf chain for ~180 us
  const int NB = nb_env ? nb_env : (n <= 4096 ? 2048 : 1024);  // measured: 4096: 2.55 vs 2.70 ms; 8192: 9.7 vs 9.3 ms
  if (n <= NB) return factor_panel(ctx, A, lda, n, 0, n, inv, s);
  // Depth-1 look-ahead: panel p+1 is factorised on the high-priority stream hp while the
  // caller's stream s applies panel p's rank-NB update to the columns right of panel p+1.
  cudaStream_t hp = ctx->serialize ? s : ctx->side;
  StreamJoin join_hp{ctx, hp, s};
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
  cudaStream_t hp = ctx->serialize ? s : ctx->side;
  StreamJoin join_hp{ctx, hp, s};
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
// scratch of the level-wise triangular inverse: one level's T blocks (at most n^2/4 at the top level, plus the
// ragged pairs), rounded up generously
size_t trtri_scratch(int n) {
  if (n <= LB) return 0;
  return (size_t)n * (size_t)n / 2 + (size_t)LB * n;
}

// all leaf inverses into the diagonal blocks of W in one launch
__global__ void copy_leaves_kernel(const double* __restrict__ inv, double* __restrict__ W, long long ldw, int n) {
  const int k = blockIdx.x, c0 = k * LB, nb = (n - c0 < LB) ? n - c0 : LB;
  const double* src = inv + (long long)k * LB * LB;
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
    const int i = idx / nb, j = idx - i * nb;
    if (j <= i) W[(long long)(c0 + i) * ldw + c0 + j] = src[i * LB + j];
  }
}

// W (n x n, zero-initialised, lower) <- L^-1, bottom-up: the leaf inverses are in place; at every level
// neighbouring diagonal blocks [W11, W22] are merged by  W21 = -W22 (L21 W11)  (two GEMMs that skip the zero
// halves of the triangular operands).  All merges of a level have the same shape except possibly the last, so
// each level is TWO batched GEMM launches: the many small products deep in the tree fill the GPU together
// instead of running as single-tile launches.  T: trtri_scratch(n) doubles.
int trtri_levels(Ctx* ctx, const double* L, long long ldl, const double* inv, double* W, long long ldw, int n,
                 double* T, cudaStream_t s) {
  const int nleaf = (n + LB - 1) / LB;
  copy_leaves_kernel<<<nleaf, 256, 0, s>>>(inv, W, ldw, n);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  struct Blk { int start, size; };
  std::vector<Blk> cur;
  for (int k = 0; k < nleaf; ++k) cur.push_back({k * LB, min(LB, n - k * LB)});
  auto merge = [&](int start, int n1, int n2, int batch, double* Tb) -> int {
    const long long step = (long long)(n1 + n2);  // distance between consecutive pairs along the diagonal
    GemmArgs g;  // T = L21 * W11   (W11 lower: k starts at the tile's first column)
    g.M = n2; g.N = n1; g.K = n1;
    g.A = L + (long long)(start + n1) * ldl + start; g.lda = ldl; g.a_kmajor = true;
    g.B = W + (long long)start * ldw + start; g.ldb = ldw; g.b_kmajor = false;
    g.C = Tb; g.ldc = n1;
    g.klo_mode = 1;
    g.batch = batch; g.strideA = step * ldl + step; g.strideB = step * ldw + step; g.strideC = (long long)n1 * n2;
    GPXB_TRY(gemm_f64(ctx, g, s));
    GemmArgs h;  // W21 = -W22 * T   (W22 lower: k ends at the tile's last row)
    h.M = n2; h.N = n1; h.K = n2;
    h.A = W + (long long)(start + n1) * ldw + start + n1; h.lda = ldw; h.a_kmajor = true;
    h.B = Tb; h.ldb = n1; h.b_kmajor = false;
    h.C = W + (long long)(start + n1) * ldw + start; h.ldc = ldw;
    h.alpha = -1.0;
    h.khi_mode = 1;
    h.batch = batch; h.strideA = step * ldw + step; h.strideB = (long long)n1 * n2; h.strideC = step * ldw + step;
    return gemm_f64(ctx, h, s);
  };
  while (cur.size() > 1) {
    const int npairs = (int)cur.size() / 2;
    const int b = cur[0].size;
    int nu = 0;  // leading run of uniform pairs
    while (nu < npairs && cur[2 * nu].size == b && cur[2 * nu + 1].size == b && cur[2 * nu].start == 2 * nu * b) ++nu;
    double* Tb = T;
    if (nu > 0) {
      GPXB_TRY(merge(0, b, b, nu, Tb));
      Tb += (size_t)nu * b * b;
    }
    for (int i = nu; i < npairs; ++i) {
      GPXB_TRY(merge(cur[2 * i].start, cur[2 * i].size, cur[2 * i + 1].size, 1, Tb));
      Tb += (size_t)cur[2 * i].size * cur[2 * i + 1].size;
    }
    std::vector<Blk> nxt;
    for (int i = 0; i < npairs; ++i) nxt.push_back({cur[2 * i].start, cur[2 * i].size + cur[2 * i + 1].size});
    if (cur.size() % 2) nxt.push_back(cur.back());
    cur.swap(nxt);
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
  GPXB_TRY(trtri_levels(ctx, L, ldl, inv, W, n, n, T, s));
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
