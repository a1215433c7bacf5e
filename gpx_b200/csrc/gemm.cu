// FP64 tensor-core GEMM for sm_100a: 128x128x16 CTA tiles, 8 warps (2x4), each warp a
// 64x32 accumulator tile of DMMA.8x8x4 fragments, 4-stage LDGSTS (cp.async) pipeline
// into padded shared memory (conflict-free fragment reads).  Used by the blocked
// Cholesky / TRSM / inverse (chol.cu) and the SGPR Woodbury path (sgpr.cu).
//
// Replaces the XLA dot_general / cholesky / triangular_solve lowering the reference
// relies on (gpx/models/_gpr.py:759-763, _sgpr.py:684-691).
#include "gemm.cuh"
#include "mma.cuh"

namespace gpxb {

namespace {

constexpr int BM = GEMM_BM, BN = GEMM_BN, BK = 16, STAGES = 4, NT = 256;
constexpr int LDK = BK + 4;   // row stride of a K-major tile  [128][16]  (doubles)
constexpr int LDM = BM + 4;   // row stride of an MN-major tile [16][128] (doubles)
constexpr int TILE_DBL = BM * LDK;  // 2560 doubles >= BK*LDM = 2112
constexpr int STAGE_DBL = 2 * TILE_DBL;
constexpr int SMEM_BYTES = STAGES * STAGE_DBL * (int)sizeof(double);  // 163840

struct GemmParams {
  int M, N, K;
  const double* A;
  long long lda;
  const double* B;
  long long ldb;
  double* C;
  long long ldc;
  double alpha, beta;
  int lower_only, klo_mode, khi_mode;
  int tiles_m, tiles_n;
  int vecA, vecB, vecC;
};

// Loads one operand tile (128 rows of the M/N index x 16 of K) into shared memory.
template <bool KMAJ>
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* g,
                                          long long ld, int r0, int rmax, int k0, int kmax, int vec,
                                          int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = tid + i * NT;
    if (KMAJ) {
      const int r = c >> 3, kc = (c & 7) * 2;
      const int gr = r0 + r, gk = k0 + kc;
      double* dst = s + r * LDK + kc;
      if (vec) {
        int nb = (gr < rmax) ? (kmax - gk) * 8 : 0;
        nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
        const double* src = nb > 0 ? g + (long long)gr * ld + gk : g;
        cp_async16(dst, src, nb);
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = (gr < rmax) && (gk + e < kmax);
          dst[e] = ok ? g[(long long)gr * ld + gk + e] : 0.0;  // unaligned slow path
        }
      }
    } else {
      const int kk = c >> 6, mc = (c & 63) * 2;
      const int gk = k0 + kk, gm = r0 + mc;
      double* dst = s + kk * LDM + mc;
      if (vec) {
        int nb = (gk < kmax) ? (rmax - gm) * 8 : 0;
        nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
        const double* src = nb > 0 ? g + (long long)gk * ld + gm : g;
        cp_async16(dst, src, nb);
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = (gk < kmax) && (gm + e < rmax);
          dst[e] = ok ? g[(long long)gk * ld + gm + e] : 0.0;  // unaligned slow path
        }
      }
    }
  }
}

template <bool AK, bool BKM>
__global__ void __launch_bounds__(NT, 1) gemm_f64_kernel(const GemmParams p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane >> 2, lq = lane & 3;

  // ---- tile coordinates ----
  int ti, tj;
  if (p.lower_only) {
    // trapezoid (M >= N): tiles_n*(tiles_n+1)/2 triangular tiles, then full tile rows
    const long long t = blockIdx.x;
    const long long ntri = (long long)p.tiles_n * (p.tiles_n + 1) / 2;
    if (t < ntri) {
      long long i = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while (i * (i + 1) / 2 > t) --i;
      while ((i + 1) * (i + 2) / 2 <= t) ++i;
      ti = (int)i;
      tj = (int)(t - i * (i + 1) / 2);
    } else {
      const long long r = t - ntri;
      ti = p.tiles_n + (int)(r / p.tiles_n);
      tj = (int)(r % p.tiles_n);
    }
  } else {
    // grouped raster: 8 tile-rows share their B column panels in L2
    const int GM = 8;
    const int t = blockIdx.x;
    const int per_group = GM * p.tiles_n;
    const int gidx = t / per_group;
    const int first_m = gidx * GM;
    const int gsz = min(p.tiles_m - first_m, GM);
    ti = first_m + (t % per_group) % gsz;
    tj = (t % per_group) / gsz;
  }
  const int row0 = ti * BM, col0 = tj * BN;

  int k_lo = (p.klo_mode == 1) ? col0 : (p.klo_mode == 2 ? row0 : 0);
  if (k_lo > p.K) k_lo = p.K;
  k_lo &= ~(BK - 1);
  int k_hi = (p.khi_mode == 1) ? min(p.K, row0 + BM) : p.K;
  const int nk = k_hi > k_lo ? (k_hi - k_lo + BK - 1) / BK : 0;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // ---- pipeline prologue ----
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      double* sa = smem + s * STAGE_DBL;
      load_tile<AK>(sa, p.A, p.lda, row0, p.M, k_lo + s * BK, p.K, p.vecA, tid);
      load_tile<BKM>(sa + TILE_DBL, p.B, p.ldb, col0, p.N, k_lo + s * BK, p.K, p.vecB, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) {
        double* sa = smem + (nxt % STAGES) * STAGE_DBL;
        load_tile<AK>(sa, p.A, p.lda, row0, p.M, k_lo + nxt * BK, p.K, p.vecA, tid);
        load_tile<BKM>(sa + TILE_DBL, p.B, p.ldb, col0, p.N, k_lo + nxt * BK, p.K, p.vecB, tid);
      }
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * STAGE_DBL;
    const double* sB = sA + TILE_DBL;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      double a[8], b[4];
#pragma unroll
      for (int mf = 0; mf < 8; ++mf) {
        a[mf] = AK ? sA[(wm * 64 + mf * 8 + lr) * LDK + ks * 4 + lq]
                   : sA[(ks * 4 + lq) * LDM + wm * 64 + mf * 8 + lr];
      }
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        b[nf] = BKM ? sB[(wn * 32 + nf * 8 + lr) * LDK + ks * 4 + lq]
                    : sB[(ks * 4 + lq) * LDM + wn * 32 + nf * 8 + lr];
      }
#pragma unroll
      for (int mf = 0; mf < 8; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf], b[nf]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ----
  const bool use_beta = (p.beta != 0.0);
#pragma unroll
  for (int mf = 0; mf < 8; ++mf) {
    const int row = row0 + wm * 64 + mf * 8 + lr;
    if (row >= p.M) continue;
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
      const int col = col0 + wn * 32 + nf * 8 + lq * 2;
      if (col >= p.N) continue;
      double* cp = p.C + (long long)row * p.ldc + col;
      const bool ok0 = !p.lower_only || col <= row;
      const bool ok1 = (col + 1 < p.N) && (!p.lower_only || col + 1 <= row);
      double v0 = p.alpha * acc[mf][nf][0], v1 = p.alpha * acc[mf][nf][1];
      if (p.vecC && ok0 && ok1) {
        if (use_beta) {
          const double2 o = *reinterpret_cast<const double2*>(cp);
          v0 += p.beta * o.x;
          v1 += p.beta * o.y;
        }
        *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
      } else {
        if (ok0) cp[0] = use_beta ? v0 + p.beta * cp[0] : v0;
        if (ok1) cp[1] = use_beta ? v1 + p.beta * cp[1] : v1;
      }
    }
  }
}

template <bool AK, bool BKM>
int launch(Ctx* ctx, const GemmParams& p, int grid, cudaStream_t stream) {
  GPXB_CUDA_CHECK(cudaFuncSetAttribute(gemm_f64_kernel<AK, BKM>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  gemm_f64_kernel<AK, BKM><<<grid, NT, SMEM_BYTES, stream>>>(p);
  GPXB_LAUNCH_CHECK();
  if (ctx) ctx->launches++;
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int gemm_f64(Ctx* ctx, const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0) return 0;
  GPXB_ARG_CHECK(g.A && g.B && g.C, -1);
  GPXB_ARG_CHECK(g.K >= 0, -1);
  GemmParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.A = g.A; p.lda = g.lda; p.B = g.B; p.ldb = g.ldb; p.C = g.C; p.ldc = g.ldc;
  p.alpha = g.alpha; p.beta = g.beta;
  p.lower_only = g.lower_only; p.klo_mode = g.klo_mode; p.khi_mode = g.khi_mode;
  p.tiles_m = ceil_div(g.M, BM);
  p.tiles_n = ceil_div(g.N, BN);
  p.vecA = aligned16(g.A) && (g.lda % 2 == 0);
  p.vecB = aligned16(g.B) && (g.ldb % 2 == 0);
  p.vecC = aligned16(g.C) && (g.ldc % 2 == 0);
  long long grid;
  if (g.lower_only) {
    GPXB_ARG_CHECK(g.M >= g.N, -2);
    grid = (long long)p.tiles_n * (p.tiles_n + 1) / 2 + (long long)(p.tiles_m - p.tiles_n) * p.tiles_n;
  } else {
    grid = (long long)p.tiles_m * p.tiles_n;
  }
  GPXB_ARG_CHECK(grid < (1LL << 31), -3);
  if (g.a_kmajor && g.b_kmajor) return launch<true, true>(ctx, p, (int)grid, stream);
  if (g.a_kmajor && !g.b_kmajor) return launch<true, false>(ctx, p, (int)grid, stream);
  if (!g.a_kmajor && g.b_kmajor) return launch<false, true>(ctx, p, (int)grid, stream);
  return launch<false, false>(ctx, p, (int)grid, stream);
}

}  // namespace gpxb
