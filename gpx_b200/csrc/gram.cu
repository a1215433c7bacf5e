// Family 1: fused stationary kernel-matrix construction for sm_100a.
//
// Replaces gpx/utils.py:55-72 (squared_distances) + gpx/kernels/kernels.py:751-757 (SE),
// 922-927 (Matern-1/2), 966-971 (Matern-3/2), 1050-1055 (Matern-5/2), the noise diagonal of
// gpx/models/_gpr.py:39-63, and -- in contraction mode -- the d/d lengthscale part of
// jit(value_and_grad(loss)) (gpx/optimizers/scipy_optimize.py:72-78).
//
// One CTA = one 128x128 tile.  Both pre-scaled operand tiles (128 rows x S doubles,
// contiguous) arrive by one 1-D bulk TMA copy each (cp.async.bulk + mbarrier); the pair
// dot products run on the FP64 tensor pipe (DMMA.8x8x4, K = D); the epilogue forms
//   d2 = |z_i|^2 - 2 z_i.z_j + |z_j|^2 + 1e-12
// applies the kernel nonlinearity and either stores K (+ diagonal shift) or contracts
// dK/d ell with W = alpha alpha^T - A^-1 without materialising dK.
#include "gram.cuh"
#include "kfun.cuh"
#include "mma.cuh"

namespace gpxb {

namespace {

constexpr int TB = 128;
constexpr int NT = 256;

struct GramParams {
  int kind;
  const double* Z1;
  int n1;
  const double* Z2;
  int n2;
  int S, Dp;
  double ell, diag_add;
  int sym, lower_only;
  long long doff;
  double* out;
  long long ldo;
  int vec_out;
  const double* Ainv;
  long long ldainv;
  const double* alpha;   // row weights  [t][n1]
  const double* alpha2;  // column weights [t][n2]
  int t;
  double* partials;
  int tiles_m, tiles_n;
};

// Kernel nonlinearity with the kind fixed at compile time: the 64-element epilogue is fully unrolled, and
// with a run-time switch per element it grew to ~29 000 SASS instructions (470 KB: every tile streamed
// through more code than the instruction caches hold).  exp is the table-driven exp_tab (kfun.cuh).
template <int KIND>
__device__ __forceinline__ double kfun_t(double d2, const double* __restrict__ tab) {
  if (KIND == KIND_SE) return exp_tab(-d2, tab);
  const double r = sqrt(fmax(d2, 1e-36));
  if (KIND == KIND_M12) return exp_tab(-r, tab);
  if (KIND == KIND_M32) {
    const double d = 1.7320508075688772 * r;
    return (1.0 + d) * exp_tab(-d, tab);
  }
  const double d = 2.23606797749979 * r;
  return (1.0 + d + d * d * 0.3333333333333333) * exp_tab(-d, tab);
}

// dK/d ell given d2 = s + 1e-12 (the jitter is a constant; the clamp passes no gradient).
// rl = 1/ell is hoisted by the caller: no FP64 division in the inner loop except Matern-1/2's 1/r.
template <int KIND>
__device__ __forceinline__ double kfun_dl_t(double d2, double s, double rl, const double* __restrict__ tab) {
  if (KIND == KIND_SE) return exp_tab(-d2, tab) * (2.0 * rl) * s;
  if (!(d2 > 1e-36)) return 0.0;
  const double r = sqrt(d2);
  if (KIND == KIND_M12) return exp_tab(-r, tab) * s * rl / r;
  if (KIND == KIND_M32) {
    const double d = 1.7320508075688772 * r;
    return (3.0 * rl) * s * exp_tab(-d, tab);
  }
  const double d = 2.23606797749979 * r;
  return ((5.0 / 3.0) * rl) * s * (1.0 + d) * exp_tab(-d, tab);
}

__global__ void prep_z_kernel(const double* __restrict__ x, long long ldx, int n, int D, int Dp,
                              int S, double ell, double* __restrict__ Z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* xi = x + (long long)i * ldx;
  double* zi = Z + (long long)i * S;
  double nrm = 0.0;
  for (int d = 0; d < D; ++d) {
    const double z = xi[d] / ell;
    zi[d] = z;
    nrm += z * z;
  }
  for (int d = D; d < S; ++d) zi[d] = 0.0;
  zi[Dp] = nrm;
}

// KT = false: the whole feature range of both operand tiles fits in shared memory (S <= GRAM_MAX_S): one bulk copy
// per operand.  KT = true (any feature count): the features stream through two stages of KTC-column slabs; a slab
// of a row-major operand is one bulk copy per row (KTC * 8 bytes each), issued by warp 0 one slab ahead of the DMMA
// loop; |z|^2 comes from two small shared arrays filled from column Dp of the operands.
constexpr int KTC = 32;   // features per slab
constexpr int KTS = 36;   // slab row stride in shared memory (== 4 mod 8: conflict-free fragment reads)
constexpr int KT_STAGE_DBL = 2 * TB * KTS;

template <bool CONTRACT, int KIND, bool KT>
__global__ void __launch_bounds__(NT, 1) gram_kernel(const GramParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sZ1 = reinterpret_cast<double*>(smem_raw + 128);
  double* sZ2 = sZ1 + (KT ? TB * KTS : TB * p.S);
  __shared__ double red[NT / 32];
  __shared__ double sN[KT ? 2 * TB : 2];  // KT: |z_i|^2 of the row tile, then of the column tile
  __shared__ double etab[64];  // 2^(j/64) for exp_tab

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  if (tid < 64) etab[tid] = exp2((double)tid * (1.0 / 64.0));
  const int lr = lane >> 2, lq = lane & 3;

  int ti, tj;
  if (p.lower_only) {
    const long long t = blockIdx.x;
    long long i = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (i * (i + 1) / 2 > t) --i;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    ti = (int)i;
    tj = (int)(t - i * (i + 1) / 2);
  } else {
    ti = blockIdx.x / p.tiles_n;
    tj = blockIdx.x % p.tiles_n;
  }
  const int row0 = ti * TB, col0 = tj * TB;
  const int rows1 = min(TB, p.n1 - row0), rows2 = min(TB, p.n2 - col0);

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (!KT) {
    if (tid == 0) {
      const uint32_t b1 = (uint32_t)rows1 * p.S * 8u, b2 = (uint32_t)rows2 * p.S * 8u;
      mbar_arrive_expect_tx(bar, b1 + b2);
      tma_bulk_g2s(sZ1, p.Z1 + (long long)row0 * p.S, b1, bar);
      tma_bulk_g2s(sZ2, p.Z2 + (long long)col0 * p.S, b2, bar);
    }
    mbar_wait(bar, 0);
    const double* a_base = sZ1 + (wm * 64 + lr) * p.S + lq;
    const double* b_base = sZ2 + (wn * 32 + lr) * p.S + lq;
    const int S8 = 8 * p.S;
    for (int k = 0; k < p.Dp; k += 4) {
      double a[8], b[4];
#pragma unroll
      for (int mf = 0; mf < 8; ++mf) a[mf] = a_base[mf * S8 + k];
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) b[nf] = b_base[nf * S8 + k];
#pragma unroll
      for (int mf = 0; mf < 8; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf], b[nf]);
    }
  } else {
    // rows beyond the matrix edge are never copied: zero both stages once so that their (masked) products stay finite
    for (int i = tid; i < 2 * KT_STAGE_DBL; i += NT) sZ1[i] = 0.0;
    if (tid < TB) sN[tid] = (tid < rows1) ? p.Z1[(long long)(row0 + tid) * p.S + p.Dp] : 0.0;
    else sN[tid] = (tid - TB < rows2) ? p.Z2[(long long)(col0 + tid - TB) * p.S + p.Dp] : 0.0;
    fence_proxy_async();  // the generic-proxy zero fill is ordered before the async-proxy bulk copies
    __syncthreads();
    const int nslab = (p.Dp + KTC - 1) / KTC;
    auto issue = [&](int sl) {  // warp 0: slab sl of both operands into stage sl & 1
      const int k0 = sl * KTC, kc = min(KTC, p.Dp - k0);
      double* d1 = sZ1 + (sl & 1) * KT_STAGE_DBL;
      double* d2 = d1 + TB * KTS;
      if (lane == 0) mbar_arrive_expect_tx(&bar[sl & 1], (uint32_t)(rows1 + rows2) * kc * 8u);
      __syncwarp();
      for (int r = lane; r < rows1; r += 32)
        tma_bulk_g2s(d1 + r * KTS, p.Z1 + (long long)(row0 + r) * p.S + k0, (uint32_t)kc * 8u, &bar[sl & 1]);
      for (int r = lane; r < rows2; r += 32)
        tma_bulk_g2s(d2 + r * KTS, p.Z2 + (long long)(col0 + r) * p.S + k0, (uint32_t)kc * 8u, &bar[sl & 1]);
    };
    if (warp == 0) issue(0);
    for (int sl = 0; sl < nslab; ++sl) {
      if (warp == 0 && sl + 1 < nslab) issue(sl + 1);  // stage (sl + 1) & 1 was released by the barrier below
      mbar_wait(&bar[sl & 1], (sl >> 1) & 1);
      const int kc = min(KTC, p.Dp - sl * KTC);
      const double* a_base = sZ1 + (sl & 1) * KT_STAGE_DBL + (wm * 64 + lr) * KTS + lq;
      const double* b_base = sZ1 + (sl & 1) * KT_STAGE_DBL + TB * KTS + (wn * 32 + lr) * KTS + lq;
      for (int k = 0; k < kc; k += 4) {
        double a[8], b[4];
#pragma unroll
        for (int mf = 0; mf < 8; ++mf) a[mf] = a_base[mf * 8 * KTS + k];
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) b[nf] = b_base[nf * 8 * KTS + k];
#pragma unroll
        for (int mf = 0; mf < 8; ++mf)
#pragma unroll
          for (int nf = 0; nf < 4; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf], b[nf]);
      }
      __syncthreads();
    }
  }
  // |z|^2 of local row r / local column c
  auto nrm1 = [&](int r) { return KT ? sN[r] : sZ1[r * p.S + p.Dp]; };
  auto nrm2 = [&](int c) { return KT ? sN[TB + c] : sZ2[c * p.S + p.Dp]; };

  // ---- fused epilogue ----
  double nj[4][2];
#pragma unroll
  for (int nf = 0; nf < 4; ++nf) {
    const int c = wn * 32 + nf * 8 + lq * 2;
    nj[nf][0] = nrm2(c);
    nj[nf][1] = nrm2(c + 1);
  }
  if (!CONTRACT) {
#pragma unroll
    for (int mf = 0; mf < 8; ++mf) {
      const int rloc = wm * 64 + mf * 8 + lr;
      const int gi = row0 + rloc;
      const double ni = nrm1(rloc);
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        const int gj = col0 + wn * 32 + nf * 8 + lq * 2;
        double v[2];
        bool ok[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = gj + e;
          ok[e] = (gi < p.n1) && (j < p.n2) && (!p.lower_only || j <= gi);
          double d2 = ((ni - 2.0 * acc[mf][nf][e]) + nj[nf][e]) + 1e-12;
          const bool on_diag = (gi + p.doff == j);
          if (p.sym && on_diag) d2 = 1e-12;
          v[e] = kfun_t<KIND>(d2, etab) + (on_diag ? p.diag_add : 0.0);
        }
        double* o = p.out + (long long)gi * p.ldo + gj;
        if (p.vec_out && ok[0] && ok[1]) {
          *reinterpret_cast<double2*>(o) = make_double2(v[0], v[1]);
        } else {
          if (ok[0]) o[0] = v[0];
          if (ok[1]) o[1] = v[1];
        }
      }
    }
    return;
  }
  // contraction: sum_ij (sum_t a_t[i] b_t[j] - Y[i][j]) dK_ij/d ell.  Tiles that lie fully inside the matrix
  // and strictly below the diagonal take the unmasked path, where every load of Y is unconditional and can
  // be scheduled ahead of the FP64 work; the (few) edge / diagonal tiles keep the per-element masks.
  const double rl = 1.0 / p.ell;
  const bool full = (row0 + TB <= p.n1) && (col0 + TB <= p.n2) && (!p.lower_only || col0 + TB <= row0) && (p.t <= 1) &&
                    !(p.sym && row0 == col0);
  double local = 0.0;
  if (full) {
    const double wscale = p.lower_only ? 2.0 : 1.0;  // off-diagonal entries of a symmetric sum count twice
    double bj[4][2];
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
      const int gj = col0 + wn * 32 + nf * 8 + lq * 2;
      bj[nf][0] = p.t ? p.alpha2[gj] : 0.0;
      bj[nf][1] = p.t ? p.alpha2[gj + 1] : 0.0;
    }
#pragma unroll
    for (int mf = 0; mf < 8; ++mf) {
      const int rloc = wm * 64 + mf * 8 + lr;
      const int gi = row0 + rloc;
      const double ni = nrm1(rloc);
      const double ai = p.t ? p.alpha[gi] : 0.0;
      double2 y[4];
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        const int gj = col0 + wn * 32 + nf * 8 + lq * 2;
        const double* yp = p.Ainv + (long long)gi * p.ldainv + gj;
        y[nf] = make_double2(yp[0], yp[1]);
      }
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        const double d20 = ((ni - 2.0 * acc[mf][nf][0]) + nj[nf][0]) + 1e-12;
        const double d21 = ((ni - 2.0 * acc[mf][nf][1]) + nj[nf][1]) + 1e-12;
        const double dk0 = kfun_dl_t<KIND>(d20, d20 - 1e-12, rl, etab);
        const double dk1 = kfun_dl_t<KIND>(d21, d21 - 1e-12, rl, etab);
        local += (ai * bj[nf][0] - y[nf].x) * dk0 + (ai * bj[nf][1] - y[nf].y) * dk1;
      }
    }
    local *= wscale;
  } else {
#pragma unroll
    for (int mf = 0; mf < 8; ++mf) {
      const int rloc = wm * 64 + mf * 8 + lr;
      const int gi = row0 + rloc;
      const double ni = nrm1(rloc);
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        const int gj = col0 + wn * 32 + nf * 8 + lq * 2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = gj + e;
          const bool ok = (gi < p.n1) && (j < p.n2) && (!p.lower_only || j <= gi);
          if (!ok) continue;
          double d2 = ((ni - 2.0 * acc[mf][nf][e]) + nj[nf][e]) + 1e-12;
          const bool on_diag = p.sym && (gi == j);
          if (on_diag) d2 = 1e-12;
          const double s = on_diag ? 0.0 : d2 - 1e-12;
          const double dk = kfun_dl_t<KIND>(d2, s, rl, etab);
          double w = -p.Ainv[(long long)gi * p.ldainv + j];
          for (int t = 0; t < p.t; ++t)
            w += p.alpha[(long long)t * p.n1 + gi] * p.alpha2[(long long)t * p.n2 + j];
          // lower-triangle evaluation of a symmetric sum counts off-diagonal entries twice
          local += ((p.lower_only && gi != j) ? 2.0 : 1.0) * w * dk;
        }
      }
    }
  }
  const double tot = block_sum<NT>(local, red);
  if (tid == 0) p.partials[blockIdx.x] = tot;
}

}  // namespace

int prep_z(Ctx* ctx, const double* x, long long ldx, int n, ZLayout zl, double ell, double* Z,
           cudaStream_t s) {
  if (n <= 0) return 0;
  prep_z_kernel<<<ceil_div(n, 128), 128, 0, s>>>(x, ldx, n, zl.D, zl.Dp, zl.S, ell, Z);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

long long gram_num_ctas(const GramArgs& a) {
  const long long tm = ceil_div(a.n1, TB), tn = ceil_div(a.n2, TB);
  return a.lower_only ? tm * (tm + 1) / 2 : tm * tn;
}

int gram_launch(Ctx* ctx, const GramArgs& a, cudaStream_t s) {
  if (a.n1 <= 0 || a.n2 <= 0) return 0;
  GPXB_ARG_CHECK(!a.lower_only || a.n1 == a.n2, -11);
  const bool kt = a.zl.S > GRAM_MAX_S;  // more features than two whole operand tiles hold: stream them in slabs
  GramParams p;
  p.kind = a.kind;
  p.Z1 = a.Z1; p.n1 = a.n1; p.Z2 = a.Z2; p.n2 = a.n2;
  p.S = a.zl.S; p.Dp = a.zl.Dp;
  p.ell = a.ell; p.diag_add = a.diag_add;
  p.sym = a.sym; p.lower_only = a.lower_only; p.doff = a.diag_off;
  GPXB_ARG_CHECK(a.diag_off == 0 || (a.out && !a.lower_only), -15);
  p.out = a.out; p.ldo = a.ldo;
  p.vec_out = a.out && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) && (a.ldo % 2 == 0);
  p.Ainv = a.Ainv; p.ldainv = a.ldainv; p.alpha = a.alpha; p.alpha2 = a.alpha2 ? a.alpha2 : a.alpha;
  p.t = a.t; p.partials = a.partials;
  p.tiles_m = ceil_div(a.n1, TB);
  p.tiles_n = ceil_div(a.n2, TB);
  const long long grid = gram_num_ctas(a);
  GPXB_ARG_CHECK(grid < (1LL << 31), -12);
  const int smem = 128 + (kt ? 2 * KT_STAGE_DBL : 2 * TB * p.S) * (int)sizeof(double);
  if (!a.out) {
    GPXB_ARG_CHECK(a.Ainv && a.partials && (a.alpha || a.t == 0), -13);
    GPXB_ARG_CHECK(a.t == 0 || a.alpha2 || (a.sym && a.n1 == a.n2), -14);
  }
#define GPXB_GRAM_GO2(C, K, T)                                                                                    \
  do {                                                                                                            \
    GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)(gram_kernel<C, K, T>), smem)); \
    gram_kernel<C, K, T><<<(int)grid, NT, smem, s>>>(p);                                                          \
  } while (0)
#define GPXB_GRAM_GO(C, K)                 \
  do {                                     \
    if (kt) GPXB_GRAM_GO2(C, K, true);     \
    else GPXB_GRAM_GO2(C, K, false);       \
  } while (0)
#define GPXB_GRAM_KIND(C)                                  \
  do {                                                     \
    switch (a.kind) {                                      \
      case KIND_SE: GPXB_GRAM_GO(C, KIND_SE); break;       \
      case KIND_M12: GPXB_GRAM_GO(C, KIND_M12); break;     \
      case KIND_M32: GPXB_GRAM_GO(C, KIND_M32); break;     \
      default: GPXB_GRAM_GO(C, KIND_M52); break;           \
    }                                                      \
  } while (0)
  if (a.out) GPXB_GRAM_KIND(false);
  else GPXB_GRAM_KIND(true);
#undef GPXB_GRAM_KIND
#undef GPXB_GRAM_GO
#undef GPXB_GRAM_GO2
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

}  // namespace gpxb
