// FP64 tensor-core GEMM for sm_100a: 128x128x32 CTA tiles, 16 warps (4x4), each warp a
// 32x32 accumulator tile of DMMA.8x8x4 fragments, 3-stage LDGSTS (cp.async) pipeline
// into padded shared memory (conflict-free fragment reads).  Used by the blocked
// Cholesky / TRSM / inverse (chol.cu) and the SGPR Woodbury path (sgpr.cu).
//
// Replaces the XLA dot_general / cholesky / triangular_solve lowering the reference
// relies on (gpx/models/_gpr.py:759-763, _sgpr.py:684-691).
#include <algorithm>

#include "gemm.cuh"
#include "mma.cuh"

#include <cstdlib>

namespace gpxb {

namespace {

#ifndef GPXB_GEMM_BK
#define GPXB_GEMM_BK 32
#endif
// BK = 32 with 3 stages (216 KB of shared memory) halves the number of CTA-wide barriers per flop
// relative to BK = 16 with 4 stages: measured 34.5 vs 33.4 TFLOP/s at 8192^3.
constexpr int BK = GPXB_GEMM_BK, STAGES = (BK == 32 ? 3 : 4);
constexpr int KCH = BK / 2;  // 16-byte chunks per K-major tile row
constexpr int LDK = BK + 4;   // row stride of a K-major tile  [T][BK]  (doubles)
// The CTA tile is T x T: 128 (16 warps, one CTA per SM) for anything that fills the machine, 64 (4 warps, two CTAs
// per SM) for launches of a few hundred big tiles or less -- a 128 x 128 tile costs 17 us per 128 of K on ONE SM
// whatever the size of the problem, which is what bounds the small GEMMs of a mid-size factorisation.
template <int T> struct TileGeom {
  static constexpr int LDM = T + 4;                 // row stride of an MN-major tile [BK][T] (doubles)
  static constexpr int TILE_DBL = T * LDK;          // >= BK * LDM
  static constexpr int STAGE_DBL = 2 * TILE_DBL;
  static constexpr int SMEM_BYTES = STAGES * STAGE_DBL * (int)sizeof(double);
};
static_assert(TileGeom<128>::TILE_DBL >= BK * TileGeom<128>::LDM && TileGeom<64>::TILE_DBL >= BK * TileGeom<64>::LDM,
              "tile buffer too small for the MN-major layout");

struct GemmParams {
  int M, N, K;
  const double* A;
  long long lda;
  const double* B;
  long long ldb;
  double* C;
  long long ldc;
  double alpha, beta;
  int lower_only, klo_mode, khi_mode, skip_tile0, tri_group;
  int tiles_m, tiles_n, tile;
  int vecA, vecB, vecC;
  long long strideA, strideB, strideC;  // batched launch: blockIdx.y selects the product
};

// Per-thread loader state for one operand: the thread's 16-byte chunks of a tile
// (128 rows of the M/N index x 16 of K).  Row/column validity is fixed for the whole K
// loop, so interior K tiles cost one LDGSTS per chunk and a pointer bump.
template <bool KMAJ, int NTHR, int T>
struct TileLoader {
  static constexpr int CH = (T * BK / 2) / NTHR;
  static constexpr int LDM = TileGeom<T>::LDM;
  const double* ptr[CH];  // global address of the chunk at k = k_lo (clamped when invalid)
  int soff[CH];           // smem offset in doubles
  int nb[CH];             // bytes valid along the fixed index (0, 8 or 16)
  int kk[CH];             // chunk's k offset inside the tile
  const double* base;
  long long ld;
  long long kstep;        // pointer increment per K tile
  int vec;
  // slow-path bookkeeping
  int r0, rmax, tid;

  __device__ __forceinline__ void init(const double* g, long long ld_, int r0_, int rmax_, int k_lo,
                                       int vec_, int tid_) {
    base = g; ld = ld_; vec = vec_; r0 = r0_; rmax = rmax_; tid = tid_;
    kstep = KMAJ ? (long long)BK : (long long)BK * ld_;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = tid_ + i * NTHR;
      if (KMAJ) {
        const int r = c / KCH, kc = (c % KCH) * 2;
        const int gr = r0_ + r;
        soff[i] = r * LDK + kc;
        kk[i] = kc;
        nb[i] = gr < rmax_ ? 16 : 0;
        ptr[i] = nb[i] ? g + (long long)gr * ld_ + k_lo + kc : g;
      } else {
        const int k = c / (T / 2), mc = (c % (T / 2)) * 2;
        const int gm = r0_ + mc;
        soff[i] = k * LDM + mc;
        kk[i] = k;
        int b = (rmax_ - gm) * 8;
        nb[i] = b < 0 ? 0 : (b > 16 ? 16 : b);
        ptr[i] = nb[i] ? g + (long long)(k_lo + k) * ld_ + gm : g;
      }
    }
  }

  // interior K tile of an aligned operand: one LDGSTS per chunk, nothing else
  __device__ __forceinline__ void load_fast(double* __restrict__ s, int tile_idx) {
    const long long off = (long long)tile_idx * kstep;
#pragma unroll
    for (int i = 0; i < CH; ++i) cp_async16(s + soff[i], ptr[i] + off, nb[i]);
  }

  // k0: first k of this tile; kmax: true K bound; interior tiles skip every check.
  __device__ __forceinline__ void load(double* __restrict__ s, int k0, int kmax, int tile_idx) {
    if (vec) {
      const bool interior = (k0 + BK <= kmax);
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        int b = nb[i];
        if (!interior) {
          if (KMAJ) {
            int kb = (kmax - (k0 + kk[i])) * 8;
            kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
            b = b < kb ? b : kb;
          } else {
            if (k0 + kk[i] >= kmax) b = 0;
          }
        }
        const double* src = ptr[i] + (long long)tile_idx * kstep;
        cp_async16(s + soff[i], b > 0 ? src : base, b);
      }
    } else {
      // unaligned slow path: plain loads, element by element
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c = tid + i * NTHR;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (KMAJ) {
            const int gr = r0 + (c / KCH), gk = k0 + (c % KCH) * 2 + e;
            const bool ok = (gr < rmax) && (gk < kmax);
            s[soff[i] + e] = ok ? base[(long long)gr * ld + gk] : 0.0;
          } else {
            const int gk = k0 + c / (T / 2), gm = r0 + (c % (T / 2)) * 2 + e;
            const bool ok = (gk < kmax) && (gm < rmax);
            s[soff[i] + e] = ok ? base[(long long)gk * ld + gm] : 0.0;
          }
        }
      }
    }
  }
};

// WM x WN warps, each a (8*MF) x (8*NF) tile:  2x4 warps of 64x32, or 4x4 warps of 32x32 (T = 128); 2x2 warps of
// 32x32 (T = 64).
// FAST (chosen by the host: both operands 16-byte aligned and K a multiple of BK, so every K tile is interior): the
// main loop carries only the one-LDGSTS-per-chunk loader.  The general loader (K-edge clamping, element-wise path for
// unaligned operands) inlined into the same loop made its body 1 845 instructions (30 KB) for 128 DMMAs.
template <bool AK, bool BKM, int WM, int WN, int T, bool FAST>
__global__ void __launch_bounds__(WM * WN * 32, T == 64 ? 2 : 1) gemm_f64_kernel(const GemmParams p) {
  constexpr int NTHR = WM * WN * 32;
  constexpr int BM = T, BN = T;
  constexpr int LDM = TileGeom<T>::LDM, TILE_DBL = TileGeom<T>::TILE_DBL, STAGE_DBL = TileGeom<T>::STAGE_DBL;
  constexpr int MF = BM / (8 * WM), NF = BN / (8 * WN);
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int lr = lane >> 2, lq = lane & 3;

  // ---- tile coordinates ----
  int ti, tj;
  if (p.lower_only) {
    // trapezoid (M >= N): tiles_n*(tiles_n+1)/2 triangular tiles, then full tile rows
    const long long t = blockIdx.x;
    const long long ntri = (long long)p.tiles_n * (p.tiles_n + 1) / 2;
    if (t < ntri && p.tri_group > 1) {
      // grouped raster on the triangle (as on the rectangle below): tile rows in groups of G, column-major inside a
      // group, so that concurrently resident CTAs share B column panels as well as A row panels
      const int G = p.tri_group;
      long long g = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5) / G;  // first guess from the row index
      auto before = [&](long long gg) { const long long r = gg * G; return r * (r + 1) / 2; };
      while (before(g) > t) --g;
      while (before(g + 1) <= t && (g + 1) * G < p.tiles_n) ++g;
      const int row_lo = (int)(g * G), rows = min(G, p.tiles_n - row_lo);
      long long u = t - before(g);
      const long long full = (long long)rows * row_lo;  // columns left of the group's own diagonal block: all rows
      if (u < full) {
        tj = (int)(u / rows);
        ti = row_lo + (int)(u % rows);
      } else {
        u -= full;  // the rows x rows lower triangle on the diagonal, column-major: column c holds rows c .. rows-1
        int c = 0;
        while (u >= rows - c) { u -= rows - c; ++c; }
        tj = row_lo + c;
        ti = row_lo + c + (int)u;
      }
    } else if (t < ntri) {
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
  if (p.skip_tile0 && row0 < 128 && col0 < 128) return;  // the leading 128 x 128 block belongs to another kernel

  int k_lo = (p.klo_mode == 1) ? col0 : (p.klo_mode == 2 ? row0 : 0);
  if (k_lo > p.K) k_lo = p.K;
  k_lo &= ~(BK - 1);
  int k_hi = (p.khi_mode == 1) ? min(p.K, row0 + BM) : p.K;
  const int nk = k_hi > k_lo ? (k_hi - k_lo + BK - 1) / BK : 0;

  double acc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  TileLoader<AK, NTHR, T> la;
  TileLoader<BKM, NTHR, T> lb;
  const double* Ab = p.A + (long long)blockIdx.y * p.strideA;
  const double* Bb = p.B + (long long)blockIdx.y * p.strideB;
  double* Cb = p.C + (long long)blockIdx.y * p.strideC;
  la.init(Ab, p.lda, row0, p.M, k_lo, p.vecA, tid);
  lb.init(Bb, p.ldb, col0, p.N, k_lo, p.vecB, tid);

  // ---- pipeline prologue ----
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      double* sa = smem + s * STAGE_DBL;
      if (FAST) {
        la.load_fast(sa, s);
        lb.load_fast(sa + TILE_DBL, s);
      } else {
        la.load(sa, k_lo + s * BK, p.K, s);
        lb.load(sa + TILE_DBL, k_lo + s * BK, p.K, s);
      }
    }
    cp_async_commit();
  }

  const int a_off = AK ? (wm * (8 * MF) + lr) * LDK + lq : lq * LDM + wm * (8 * MF) + lr;
  const int b_off = BKM ? (wn * (8 * NF) + lr) * LDK + lq : lq * LDM + wn * (8 * NF) + lr;

  // Warps of one scheduler (warp % 4) issue their global->shared copies at different k-steps
  // of the tile, so the tensor pipe always has DMMAs from the other warps to run.
  const int lpos = ((warp >> 2) & 3) * (BK / 16);
  const bool fast = p.vecA && p.vecB;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    const bool do_load = nxt < nk;
    const bool interior = fast && (k_lo + (nxt + 1) * BK <= p.K);
    double* sn = smem + (nxt % STAGES) * STAGE_DBL;
    const double* sA = smem + (kt % STAGES) * STAGE_DBL + a_off;
    const double* sB = smem + (kt % STAGES) * STAGE_DBL + TILE_DBL + b_off;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      if (ks == lpos) {
        if (do_load) {
          if (FAST || interior) {
            la.load_fast(sn, nxt);
            lb.load_fast(sn + TILE_DBL, nxt);
          } else {
            la.load(sn, k_lo + nxt * BK, p.K, nxt);
            lb.load(sn + TILE_DBL, k_lo + nxt * BK, p.K, nxt);
          }
        }
        cp_async_commit();
      }
      double a[MF], b[NF];
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) a[mf] = AK ? sA[mf * 8 * LDK + ks * 4] : sA[ks * 4 * LDM + mf * 8];
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) b[nf] = BKM ? sB[nf * 8 * LDK + ks * 4] : sB[ks * 4 * LDM + nf * 8];
#pragma unroll
      for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf], b[nf]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ----
  const bool use_beta = (p.beta != 0.0);
#pragma unroll
  for (int mf = 0; mf < MF; ++mf) {
    const int row = row0 + wm * (8 * MF) + mf * 8 + lr;
    if (row >= p.M) continue;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) {
      const int col = col0 + wn * (8 * NF) + nf * 8 + lq * 2;
      if (col >= p.N) continue;
      double* cp = Cb + (long long)row * p.ldc + col;
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

int gemm_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GPXB_GEMM_WARPS");
    v = (e && atoi(e) == 8) ? 8 : 16;
  }
  return v;
}

// Algorithmic flops of one launch (one batch member), used only by the profiling hooks: 2 x (output entries
// that are part of the result) x (length of the k range the tile's mode prescribes).  Lower-only launches count
// the lower triangle of the diagonal tiles only; triangular-operand modes count each tile's own k range, as the
// kernel executes it (within 128/N of the exact triangular count).
double launch_flops(const GemmParams& p, long long grid) {
  (void)grid;
  if (p.klo_mode == 0 && p.khi_mode == 0) {
    double area;
    if (p.lower_only) {
      const double n = (double)p.N, m = (double)p.M;
      area = n * (n + 1.0) * 0.5 + (m - n) * n;
    } else {
      area = (double)p.M * (double)p.N;
    }
    return 2.0 * area * (double)p.K;
  }
  double fl = 0.0;
  const int BM = p.tile, BN = p.tile;
  for (int ti = 0; ti < p.tiles_m; ++ti) {
    const int row0 = ti * BM, rows = std::min(BM, p.M - row0);
    const int tj_end = p.lower_only ? std::min(p.tiles_n, ti + 1) : p.tiles_n;
    for (int tj = 0; tj < tj_end; ++tj) {
      const int col0 = tj * BN, cols = std::min(BN, p.N - col0);
      int k_lo = (p.klo_mode == 1) ? col0 : (p.klo_mode == 2 ? row0 : 0);
      k_lo = std::min(k_lo, p.K);
      const int k_hi = (p.khi_mode == 1) ? std::min(p.K, row0 + BM) : p.K;
      if (k_hi <= k_lo) continue;
      double area = (double)rows * cols;
      if (p.lower_only && ti == tj) area = (double)rows * (rows + 1) * 0.5;
      fl += 2.0 * area * (double)(k_hi - k_lo);
    }
  }
  return fl;
}

template <bool AK, bool BKM>
int launch(Ctx* ctx, const GemmParams& p, int grid1, int batch, cudaStream_t stream) {
  const dim3 grid(grid1, batch);
  Ctx::ProfRec rec{nullptr, nullptr, 0.0};
  const bool prof = ctx && ctx->prof;
  if (prof) {
    GPXB_CUDA_CHECK(cudaEventCreate(&rec.e0));
    GPXB_CUDA_CHECK(cudaEventCreate(&rec.e1));
    rec.flops = launch_flops(p, grid1) * batch;
    GPXB_CUDA_CHECK(cudaEventRecord(rec.e0, stream));
  }
  // every K tile interior and both operands aligned: the instantiation whose main loop has only the fast loader
  static const bool allow_fast = [] {
    const char* e = getenv("GPXB_GEMM_FAST");  // "0": always the general loader (comparison)
    return !(e && e[0] == '0');
  }();
  const bool fast = allow_fast && p.vecA && p.vecB && (p.K % BK == 0) && p.K > 0;
  auto go = [&](void (*kern)(const GemmParams), int threads, int smem) -> cudaError_t {
    const cudaError_t e = ensure_dyn_smem((const void*)kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, smem, stream>>>(p);
    return cudaSuccess;
  };
  constexpr int SM64 = TileGeom<64>::SMEM_BYTES, SM128 = TileGeom<128>::SMEM_BYTES;
  if (p.tile == 64) {
    if (fast) GPXB_CUDA_CHECK(go(gemm_f64_kernel<AK, BKM, 2, 2, 64, true>, 128, SM64));
    else GPXB_CUDA_CHECK(go(gemm_f64_kernel<AK, BKM, 2, 2, 64, false>, 128, SM64));
  } else if (gemm_variant() == 8) {
    GPXB_CUDA_CHECK(go(gemm_f64_kernel<AK, BKM, 2, 4, 128, false>, 256, SM128));
  } else {
    if (fast) GPXB_CUDA_CHECK(go(gemm_f64_kernel<AK, BKM, 4, 4, 128, true>, 512, SM128));
    else GPXB_CUDA_CHECK(go(gemm_f64_kernel<AK, BKM, 4, 4, 128, false>, 512, SM128));
  }
  GPXB_LAUNCH_CHECK();
  if (prof) {
    GPXB_CUDA_CHECK(cudaEventRecord(rec.e1, stream));
    ctx->prof_recs.push_back(rec);
  }
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
  p.skip_tile0 = g.lower_only ? g.skip_tile0 : 0;
  static const int tri_group = [] {
    // tile rows per raster group on the lower triangle; 1 (default): row-major.  Measured on the trailing-update
    // shapes (tools/bench_syrk.py, profiles/r2_syrk_raster.jsonl): groups of 4 / 8 / 16 change nothing (31.7 vs
    // 31.8 TFLOP/s at n = 16384, K = 1024) -- the 134 MB operand is L2-resident either way; the gap to the square
    // GEMM is per-tile prologue / read-modify-write epilogue at K = 1024, not operand locality.
    const char* e = getenv("GPXB_GEMM_TRI_GROUP");
    const int v = e ? atoi(e) : 1;
    return v >= 1 && v <= 64 ? v : 1;
  }();
  p.tri_group = (g.skip_tile0 || g.klo_mode || g.khi_mode) ? 1 : tri_group;  // tile 0 must stay block 0 when it is skipped
  // 64 x 64 tiles when 128 x 128 ones would not give every SM a few CTAs (GPXB_GEMM_SMALL_MAX: the largest count of
  // big tiles still served by the small kernel; 0 turns it off)
  static const long long small_max = [] {
    const char* e = getenv("GPXB_GEMM_SMALL_MAX");
    return e ? atoll(e) : -1;
  }();
  auto tiles_of = [&](int T, int& tm, int& tn) -> long long {
    tm = ceil_div(g.M, T);
    tn = ceil_div(g.N, T);
    return g.lower_only ? (long long)tn * (tn + 1) / 2 + (long long)(tm - tn) * tn : (long long)tm * tn;
  };
  int tm128, tn128;
  const long long grid128 = tiles_of(128, tm128, tn128) * g.batch;
  // default: less than ~3/4 of a wave of big tiles (measured, tools/bench_gemm_small.py: 1024^3 85 vs 152 us, skinny
  // K = 128 products 17 vs 27 us; from one full wave on the big tile wins: 2048^2 x 1024 lower 155 vs 169 us)
  // triangular-operand products (per-tile k ranges of very different length) balance better on small tiles: up to
  // four waves of big ones (LML + gradient at n = 2000: 2.06 vs 2.34 ms)
  const long long lim = small_max >= 0 ? small_max
                        : (g.small_tile_max > 0 ? g.small_tile_max : ((g.klo_mode || g.khi_mode) ? 592 : 110));
  // In-place products (C aliasing an operand: the callers' B <- B inv^T with N <= 128) rely on ONE column tile per
  // tile row -- each CTA consumes its whole operand rows before its epilogue writes them -- so they keep the 128-wide
  // tile: with 64-wide tiles the CTA of columns 0..63 may overwrite what the CTA of columns 64..127 still reads.
  const bool in_place = (g.C == g.A) || (g.C == g.B);
  GPXB_ARG_CHECK(!in_place || g.N <= 128, -5);
  p.tile = (grid128 <= lim && !in_place) ? 64 : 128;
  p.tiles_m = ceil_div(g.M, p.tile);
  p.tiles_n = ceil_div(g.N, p.tile);
  GPXB_ARG_CHECK(g.batch >= 1 && g.batch <= 65535, -4);
  p.strideA = g.strideA; p.strideB = g.strideB; p.strideC = g.strideC;
  const bool even_strides = g.batch == 1 || (g.strideA % 2 == 0 && g.strideB % 2 == 0 && g.strideC % 2 == 0);
  p.vecA = aligned16(g.A) && (g.lda % 2 == 0) && even_strides;
  p.vecB = aligned16(g.B) && (g.ldb % 2 == 0) && even_strides;
  p.vecC = aligned16(g.C) && (g.ldc % 2 == 0) && even_strides;
  long long grid;
  if (g.lower_only) {
    GPXB_ARG_CHECK(g.M >= g.N, -2);
    grid = (long long)p.tiles_n * (p.tiles_n + 1) / 2 + (long long)(p.tiles_m - p.tiles_n) * p.tiles_n;
  } else {
    grid = (long long)p.tiles_m * p.tiles_n;
  }
  GPXB_ARG_CHECK(grid < (1LL << 31), -3);
  if (g.a_kmajor && g.b_kmajor) return launch<true, true>(ctx, p, (int)grid, g.batch, stream);
  if (g.a_kmajor && !g.b_kmajor) return launch<true, false>(ctx, p, (int)grid, g.batch, stream);
  if (!g.a_kmajor && g.b_kmajor) return launch<false, true>(ctx, p, (int)grid, g.batch, stream);
  return launch<false, false>(ctx, p, (int)grid, g.batch, stream);
}

}  // namespace gpxb
