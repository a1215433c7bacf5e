// Family 1b: Hessian-Jacobian kernel matrix (training on derivative values) for sm_100a.
//
// Replaces _squared_exponential_kernel_d01kj (gpx/kernels/kernels.py:804-831) and
// _matern52_kernel_d01kj (gpx/kernels/kernels.py:1253-1283), _A_derivs_lhs
// (gpx/models/_gpr.py:66-91), _lml_derivs_dense (824-870) and _fit_derivs_dense (313-344).
//
// With row index i = s*nv + v and A[i][f] = J[s][f][v] the reference's einsums collapse to
//   out[i][j] = c1[s][t] * (A1 A2^T)[i][j] - c2[s][t] * a_i,t * b_j,s
//   a_i,t = kappa (q1_i - (A1 x2^T)[i][t]),  b_j,s = kappa ((A2 x1^T)[j][s] - q2_j),  q_i = A[i] . x_s
//   SE:   kappa = 2/l^2,   c2 = exp(-d2),             c1 = (2/l^2) c2
//   M52:  kappa = sqrt5/l, c2 = 5/(3 l^2) exp(-d),    c1 = c2 (1 + d)          
// d/d ell keeps this form with other pair scalars (s = d2 - 1e-12, rho = s/d2; kappa's own ell
// dependence is folded in), so the hyper-parameter gradient sum_ij W_ij dA_ij/d ell is the same kernel
// in contraction mode (no dA is ever stored):
//   SE:   c1' = c1 (2s - 2)/l,                   c2' = c2 (2s - 4)/l
//   M52:  c1' = c2 (d^2 rho - 2d - 2)/l,         c2' = c2 (d rho - 4)/l
// so the (N nv) x (N nv) matrix is ONE FP64 tensor-core GEMM (K = nf, operand tiles by bulk TMA)
// with a fused per-(s,t) scalar / rank-1 epilogue; the reference's (N, N, nf) temporaries are
// never formed.
#include "../../include/gpxb200.h"
#include "chol.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "hess.cuh"
#include "matvec.cuh"
#include "mma.cuh"
#include "rpchol.cuh"
#include "small_kernels.cuh"

namespace gpxb {

namespace {

constexpr int TB = 128;
constexpr int NT = 256;

struct HessParams {
  const double* A1;
  int n1;
  const double* A2;
  int n2;
  int S, Dp, nv1, nv2, Ns1, Ns2;
  const double* C1;
  const double* C2;
  const double* P12;  // [n1][Ns2]
  const double* P21;  // [n2][Ns1]
  double kappa, diag_add;
  int sym, lower_only;
  double* out;
  long long ldo;  // ldo < 0: block-column packed lower storage (chol.cuh: BcpLayout) of an n1 x n1 matrix
  int vec_out, tiles_m, tiles_n;
  // contraction mode: partials[cta] = sum over the lower triangle of (2 - [i==j]) (alpha_i alpha_j - Ainv_ij) v_ij
  const double* Ainv;
  long long ldainv;
  const double* alpha;
  double* partials;
};

// AJ[i = s*nv + v][f] = J[s][f][v];  AJ[i][Dp] = q_i = sum_f J[s][f][v] x[s][f]
__global__ void prep_hess_kernel(const double* __restrict__ x, const double* __restrict__ J, int N, int nf,
                                 int nv, int S, int Dp, double* __restrict__ AJ) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * nv) return;
  const int s = (int)(i / nv), v = (int)(i % nv);
  const double* Js = J + (long long)s * nf * nv + v;
  const double* xs = x + (long long)s * nf;
  double* a = AJ + i * S;
  double q = 0.0;
  for (int f = 0; f < nf; ++f) {
    const double jv = Js[(long long)f * nv];
    a[f] = jv;
    q += jv * xs[f];
  }
  for (int f = nf; f < S; ++f) a[f] = 0.0;
  a[Dp] = q;
}

// per sample pair: c1, c2 from d2 = |z_s|^2 - 2 z_s.z_t + |z_t|^2 + 1e-12 (exact 1e-12 on the diagonal)
__global__ void pair_coef_kernel(const double* __restrict__ Z1, int Ns1, const double* __restrict__ Z2,
                                 int Ns2, int S, int Dp, int kind, double ell, int sym, int dl,
                                 double* __restrict__ C1, double* __restrict__ C2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Ns1 * Ns2) return;
  const int s = (int)(idx / Ns2), t = (int)(idx % Ns2);
  const double* a = Z1 + (long long)s * S;
  const double* b = Z2 + (long long)t * S;
  double dot = 0.0;
  for (int f = 0; f < Dp; ++f) dot += a[f] * b[f];
  double d2 = ((a[Dp] - 2.0 * dot) + b[Dp]) + 1e-12;
  if (sym && s == t) d2 = 1e-12;
  const double sq = (sym && s == t) ? 0.0 : d2 - 1e-12;  // the jitter is a constant
  if (kind == KIND_SE) {
    const double e = exp(-d2);
    const double c1 = (2.0 / (ell * ell)) * e;
    C2[idx] = dl ? e * (2.0 * sq - 4.0) / ell : e;
    C1[idx] = dl ? c1 * (2.0 * sq - 2.0) / ell : c1;
  } else {
    const double d = 2.23606797749979 * sqrt(fmax(d2, 1e-36));
    const double c = (5.0 / (3.0 * ell * ell)) * exp(-d);
    const double rho = d2 > 1e-36 ? sq / d2 : 0.0;  // the clamp passes no gradient
    C2[idx] = dl ? c * (d * rho - 4.0) / ell : c;
    C1[idx] = dl ? c * (d * d * rho - 2.0 * d - 2.0) / ell : c * (1.0 + d);
  }
}

template <bool CONTRACT>
__global__ void __launch_bounds__(NT, 1) hess_kernel(const HessParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[NT / 32];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sA1 = reinterpret_cast<double*>(smem_raw + 128);
  double* sA2 = sA1 + TB * p.S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
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
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t b1 = (uint32_t)rows1 * p.S * 8u, b2 = (uint32_t)rows2 * p.S * 8u;
    mbar_arrive_expect_tx(bar, b1 + b2);
    tma_bulk_g2s(sA1, p.A1 + (long long)row0 * p.S, b1, bar);
    tma_bulk_g2s(sA2, p.A2 + (long long)col0 * p.S, b2, bar);
  }
  mbar_wait(bar, 0);

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* a_base = sA1 + (wm * 64 + lr) * p.S + lq;
  const double* b_base = sA2 + (wn * 32 + lr) * p.S + lq;
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

  // ---- fused epilogue: per-(s,t) scalars and the rank-1 term ----
  double local = 0.0;
#pragma unroll
  for (int nf = 0; nf < 4; ++nf) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cl = wn * 32 + nf * 8 + lq * 2 + e;
      const int gj = col0 + cl;
      if (gj >= p.n2) continue;
      const int t = gj / p.nv2;
      const double q2 = sA2[cl * p.S + p.Dp];
      const double alj = CONTRACT ? p.alpha[gj] : 0.0;
#pragma unroll
      for (int mf = 0; mf < 8; ++mf) {
        const int rl = wm * 64 + mf * 8 + lr;
        const int gi = row0 + rl;
        if (gi >= p.n1 || (p.lower_only && gj > gi)) continue;
        const int s = gi / p.nv1;
        const double q1 = sA1[rl * p.S + p.Dp];
        const double a = p.kappa * (q1 - p.P12[(long long)gi * p.Ns2 + t]);
        const double b = p.kappa * (p.P21[(long long)gj * p.Ns1 + s] - q2);
        const long long st = (long long)s * p.Ns2 + t;
        double v = p.C1[st] * acc[mf][nf][e] - p.C2[st] * a * b;
        if (CONTRACT) {
          const double w = p.alpha[gi] * alj - p.Ainv[(long long)gi * p.ldainv + gj];
          local += (gi != gj ? 2.0 : 1.0) * w * v;
        } else {
          if (p.sym && gi == gj) v += p.diag_add;
          const long long o = p.ldo >= 0 ? (long long)gi * p.ldo + gj : BcpLayout{p.n1}.index(gi, gj);
          p.out[o] = v;
        }
      }
    }
  }
  if (CONTRACT) {
    const double tot = block_sum<NT>(local, red);
    if (tid == 0) p.partials[blockIdx.x] = tot;
  }
}

}  // namespace

// out[(n1*nv1) x (n2*nv2)] = d01kj(x1, x2; J1, J2) (+ diag_add I when the operands are the same).
// ldo < 0 (with lower_only): out is block-column packed (BcpLayout).
//
// Contraction mode (out == nullptr; symmetric operands): *contract_out = sum_ij W_ij dA_ij/d ell over the full
// symmetric matrix with W = alpha alpha^T - Ainv (Ainv lower-stored, ld ldainv), evaluated on the lower triangle.
static int hess_run(Ctx* ctx, int kind, const double* x1, const double* J1, int Ns1, int nv1, const double* x2,
                    const double* J2, int Ns2, int nv2, int nf, double ell, double diag_add, double* out,
                    long long ldo, int lower_only, const double* Ainv, long long ldainv, const double* alpha,
                    double* contract_out, cudaStream_t s) {
  const bool contract = out == nullptr;
  GPXB_ARG_CHECK(kind == KIND_SE || kind == KIND_M52, -50);  // the reference hand-writes d01kj for these two
  GPXB_ARG_CHECK(Ns1 > 0 && Ns2 > 0 && nv1 > 0 && nv2 > 0 && nf > 0, -51);
  const ZLayout zl = z_layout(nf);
  GPXB_ARG_CHECK(zl.S <= GRAM_MAX_S, -52);
  const bool sym = (x1 == x2) && (J1 == J2) && (Ns1 == Ns2) && (nv1 == nv2);
  GPXB_ARG_CHECK(!lower_only || sym, -53);
  GPXB_ARG_CHECK(ldo >= 0 || lower_only, -56);
  GPXB_ARG_CHECK(!contract || (sym && lower_only && Ainv && alpha && contract_out), -57);
  const long long n1 = (long long)Ns1 * nv1, n2 = (long long)Ns2 * nv2;
  GPXB_ARG_CHECK(n1 < (1LL << 31) && n2 < (1LL << 31), -54);
  DevBuf bA1, bA2, bZ1, bZ2, bX1, bX2, bC1, bC2, bP12, bP21;
  GPXB_TRY(bA1.alloc(sizeof(double) * (size_t)n1 * zl.S, s));
  GPXB_TRY(bZ1.alloc(sizeof(double) * (size_t)Ns1 * zl.S, s));
  GPXB_TRY(bX1.alloc(sizeof(double) * (size_t)Ns1 * zl.S, s));
  GPXB_TRY(bC1.alloc(sizeof(double) * (size_t)Ns1 * Ns2, s));
  GPXB_TRY(bC2.alloc(sizeof(double) * (size_t)Ns1 * Ns2, s));
  GPXB_TRY(bP12.alloc(sizeof(double) * (size_t)n1 * Ns2, s));
  double *A1 = bA1.as<double>(), *Z1 = bZ1.as<double>(), *X1 = bX1.as<double>();
  double *A2 = A1, *Z2 = Z1, *X2 = X1;
  prep_hess_kernel<<<ceil_div(n1, 128), 128, 0, s>>>(x1, J1, Ns1, nf, nv1, zl.S, zl.Dp, A1);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(prep_z(ctx, x1, nf, Ns1, zl, ell, Z1, s));
  GPXB_TRY(prep_z(ctx, x1, nf, Ns1, zl, 1.0, X1, s));
  if (!sym) {
    GPXB_TRY(bA2.alloc(sizeof(double) * (size_t)n2 * zl.S, s));
    GPXB_TRY(bZ2.alloc(sizeof(double) * (size_t)Ns2 * zl.S, s));
    GPXB_TRY(bX2.alloc(sizeof(double) * (size_t)Ns2 * zl.S, s));
    A2 = bA2.as<double>(); Z2 = bZ2.as<double>(); X2 = bX2.as<double>();
    prep_hess_kernel<<<ceil_div(n2, 128), 128, 0, s>>>(x2, J2, Ns2, nf, nv2, zl.S, zl.Dp, A2);
    GPXB_LAUNCHED(ctx);
    GPXB_TRY(prep_z(ctx, x2, nf, Ns2, zl, ell, Z2, s));
    GPXB_TRY(prep_z(ctx, x2, nf, Ns2, zl, 1.0, X2, s));
  }
  pair_coef_kernel<<<ceil_div((long long)Ns1 * Ns2, 256), 256, 0, s>>>(Z1, Ns1, Z2, Ns2, zl.S, zl.Dp, kind, ell,
                                                                         sym ? 1 : 0, contract ? 1 : 0,
                                                                         bC1.as<double>(), bC2.as<double>());
  GPXB_LAUNCHED(ctx);
  {
    GemmArgs g;  // P12 = A1 X2^T   [n1 x Ns2]
    g.M = (int)n1; g.N = Ns2; g.K = zl.Dp;
    g.A = A1; g.lda = zl.S; g.a_kmajor = true;
    g.B = X2; g.ldb = zl.S; g.b_kmajor = true;
    g.C = bP12.as<double>(); g.ldc = Ns2;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  const double* P21 = bP12.as<double>();
  if (!sym) {
    GPXB_TRY(bP21.alloc(sizeof(double) * (size_t)n2 * Ns1, s));
    GemmArgs g;  // P21 = A2 X1^T   [n2 x Ns1]
    g.M = (int)n2; g.N = Ns1; g.K = zl.Dp;
    g.A = A2; g.lda = zl.S; g.a_kmajor = true;
    g.B = X1; g.ldb = zl.S; g.b_kmajor = true;
    g.C = bP21.as<double>(); g.ldc = Ns1;
    GPXB_TRY(gemm_f64(ctx, g, s));
    P21 = bP21.as<double>();
  }
  HessParams p;
  p.A1 = A1; p.n1 = (int)n1; p.A2 = A2; p.n2 = (int)n2;
  p.S = zl.S; p.Dp = zl.Dp; p.nv1 = nv1; p.nv2 = nv2; p.Ns1 = Ns1; p.Ns2 = Ns2;
  p.C1 = bC1.as<double>(); p.C2 = bC2.as<double>(); p.P12 = bP12.as<double>(); p.P21 = P21;
  p.kappa = (kind == KIND_SE) ? 2.0 / (ell * ell) : 2.23606797749979 / ell;
  p.diag_add = diag_add; p.sym = sym ? 1 : 0; p.lower_only = lower_only;
  p.out = out; p.ldo = ldo; p.vec_out = 0;
  p.tiles_m = ceil_div(n1, TB); p.tiles_n = ceil_div(n2, TB);
  const long long grid = lower_only ? (long long)p.tiles_m * (p.tiles_m + 1) / 2 : (long long)p.tiles_m * p.tiles_n;
  GPXB_ARG_CHECK(grid < (1LL << 31), -55);
  const int smem = 128 + 2 * TB * p.S * (int)sizeof(double);
  p.Ainv = Ainv; p.ldainv = ldainv; p.alpha = alpha; p.partials = nullptr;
  if (!contract) {
    GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)(hess_kernel<false>), smem));
    hess_kernel<false><<<(int)grid, NT, smem, s>>>(p);
    GPXB_LAUNCHED(ctx);
    return 0;
  }
  DevBuf bPart;
  GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)grid, s));
  p.partials = bPart.as<double>();
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)(hess_kernel<true>), smem));
  hess_kernel<true><<<(int)grid, NT, smem, s>>>(p);
  GPXB_LAUNCHED(ctx);
  sk::sum_kernel<<<1, 1024, 0, s>>>(p.partials, grid, contract_out);  // fixed order: deterministic
  GPXB_LAUNCHED(ctx);
  return 0;
}

// *contract_out = sum_ij (alpha_i alpha_j - Ainv_ij) d(d01kj(x, x; J, J))_ij / d lengthscale over the full symmetric
// matrix (Ainv lower-stored with ld ldainv); dA is never formed.
int hess_contract_dl(Ctx* ctx, int kind, const double* x, const double* J, int N, int nv, int nf, double ell,
                     const double* Ainv, long long ldainv, const double* alpha, double* contract_out, cudaStream_t s) {
  return hess_run(ctx, kind, x, J, N, nv, x, J, N, nv, nf, ell, 0.0, nullptr, ldainv, 1, Ainv, ldainv, alpha,
                  contract_out, s);
}

int hess_gram(Ctx* ctx, int kind, const double* x1, const double* J1, int Ns1, int nv1, const double* x2,
              const double* J2, int Ns2, int nv2, int nf, double ell, double diag_add, double* out, long long ldo,
              int lower_only, cudaStream_t s) {
  GPXB_ARG_CHECK(out != nullptr, -58);
  return hess_run(ctx, kind, x1, J1, Ns1, nv1, x2, J2, Ns2, nv2, nf, ell, diag_add, out, ldo, lower_only, nullptr, 0,
                  nullptr, nullptr, s);
}

namespace {
// scal: [0] quad, [1] sum log L_ii
__global__ void finalize_derivs_kernel(const double* __restrict__ scal, long long n, double* __restrict__ out) {
  const double m = (double)n;
  out[0] = (-0.5 * scal[0] - scal[1] - m * 0.5 * log(2.0 * 3.141592653589793)) / m;
}
// jaccoef[s][f] = sum_v c[s*nv + v] J[s][f][v]     (gpx/models/gpr.py:475-476)
__global__ void jaccoef_kernel(const double* __restrict__ c, const double* __restrict__ J, int N, int nf, int nv,
                               double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * nf) return;
  const int s = (int)(idx / nf);
  const double* Jr = J + idx * nv;
  const double* cs = c + (long long)s * nv;
  double a = 0.0;
  for (int v = 0; v < nv; ++v) a += cs[v] * Jr[v];
  out[idx] = a;
}
}  // namespace

// lml/n of _lml_derivs_dense (zero mean); if c_out: also c = A^-1 y (n) and jaccoef (N x nf).
int gpr_derivs(Ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf, int nv,
               double ell, double sigma, double* lml_out, double* c_out, double* jaccoef_out, cudaStream_t s) {
  const long long n = (long long)N * nv;
  GPXB_ARG_CHECK(n > 0 && n < (1LL << 31), -1);
  const double s2 = sigma * sigma + 1e-10;
  // block-column packed lower storage for large systems: half the memory (the 180 000-row system is
  // 259 GB square, 131 GB packed) and, with its 8 KB instead of n*8 B row stride, measured faster than
  // the square layout at 108 000 rows (14.0 s vs 17.05 s) though not at 32 760 (0.435 s vs 0.409 s).
  // GPXB_DERIVS_LAYOUT=packed|square overrides; read per call so that tests can flip it.
  bool packed = n >= 65536;
  if (const char* e = getenv("GPXB_DERIVS_LAYOUT")) packed = (e[0] == 'p') ? true : (e[0] == 's' ? false : packed);
  const BcpLayout lay{n};
  DevBuf bA, bInv, bR, bSc;
  GPXB_TRY(bA.alloc(sizeof(double) * (packed ? lay.doubles() : (size_t)n * n), s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles((int)n), s));
  GPXB_TRY(bR.alloc(sizeof(double) * (size_t)n, s));
  GPXB_TRY(bSc.alloc(sizeof(double) * 4, s));
  double *A = bA.as<double>(), *inv = bInv.as<double>(), *r = bR.as<double>(), *scal = bSc.as<double>();
  GPXB_TRY(hess_gram(ctx, kind, x, J, N, nv, x, J, N, nv, nf, ell, s2, A, packed ? -1 : n, 1, s));
  GPXB_CUDA_CHECK(cudaMemcpyAsync(r, y, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  if (packed) {
    GPXB_TRY(potrf_lower_bcp(ctx, A, (int)n, inv, s));
    GPXB_TRY(trsv_bcp(ctx, A, (int)n, inv, r, n, 1, true, s));
  } else {
    GPXB_TRY(potrf_lower(ctx, A, n, (int)n, inv, s));
    GPXB_TRY(trsm_right_lt(ctx, A, n, inv, r, n, 1, (int)n, s));
  }
  if (lml_out) {
    if (packed) GPXB_TRY(logdet_bcp(ctx, A, (int)n, scal + 1, 1.0, s));
    else GPXB_TRY(logdet_lower(ctx, A, n, (int)n, scal + 1, 1.0, s));
    sk::sumsq_kernel<<<1, 1024, 0, s>>>(r, n, scal + 0, 0);
    GPXB_LAUNCHED(ctx);
    finalize_derivs_kernel<<<1, 1, 0, s>>>(scal, n, lml_out);
    GPXB_LAUNCHED(ctx);
  }
  if (c_out) {
    if (packed) GPXB_TRY(trsv_bcp(ctx, A, (int)n, inv, r, n, 1, false, s));
    else GPXB_TRY(trsm_right_l(ctx, A, n, inv, r, n, 1, (int)n, s));
    GPXB_CUDA_CHECK(cudaMemcpyAsync(c_out, r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    if (jaccoef_out) {
      jaccoef_kernel<<<ceil_div((long long)N * nf, 256), 256, 0, s>>>(r, J, N, nf, nv, jaccoef_out);
      GPXB_LAUNCHED(ctx);
    }
  }
  return 0;
}

namespace {
// scal: [0] quad, [1] sum log L_ii, [2] sum W dA/d ell, [3] sum alpha^2, [4] tr(Ainv)
__global__ void finalize_derivs_grad_kernel(const double* __restrict__ scal, long long n, double sigma,
                                            int want_grad, double* __restrict__ out3) {
  const double m = (double)n;
  out3[0] = (-0.5 * scal[0] - scal[1] - m * 0.5 * log(2.0 * 3.141592653589793)) / m;
  out3[1] = want_grad ? 0.5 * scal[2] / m : 0.0;
  out3[2] = want_grad ? sigma * (scal[3] - scal[4]) / m : 0.0;
}
}  // namespace

// out3 = {lml, d lml/d ell, d lml/d sigma} of _lml_derivs_dense: what jit(value_and_grad(loss)) yields for
// neg_log_marginal_likelihood_derivs (gpx/models/gpr.py:276-300, gpx/optimizers/scipy_optimize.py:72-78,
// 92-122), in analytic form: alpha = A^-1 y, W = alpha alpha^T - A^-1,
// d lml/d ell = (1/2 sum_ij W_ij dA_ij/d ell)/n, d lml/d sigma = sigma tr(W)/n, n = N nv.
int gpr_derivs_grad(Ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf, int nv,
                    double ell, double sigma, int want_grad, double* out3, cudaStream_t s) {
  const long long n = (long long)N * nv;
  GPXB_ARG_CHECK(n > 0 && n < (1LL << 31), -1);
  if (!want_grad) {
    GPXB_TRY(gpr_derivs(ctx, kind, x, J, y, N, nf, nv, ell, sigma, out3, nullptr, nullptr, s));
    GPXB_CUDA_CHECK(cudaMemsetAsync(out3 + 1, 0, 2 * sizeof(double), s));
    return 0;
  }
  // the inverse needs square storage: three n x n buffers (factor / inverse, trtri workspace)
  size_t free_b = 0, total_b = 0;
  GPXB_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  GPXB_ARG_CHECK(2.2 * sizeof(double) * (double)n * (double)n < (double)total_b, -60);
  const double s2 = sigma * sigma + 1e-10;
  DevBuf bA, bInv, bR, bSc, bW, bT;
  GPXB_TRY(bA.alloc(sizeof(double) * (size_t)n * n, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles((int)n), s));
  GPXB_TRY(bR.alloc(sizeof(double) * (size_t)n, s));
  GPXB_TRY(bSc.alloc(sizeof(double) * 8, s));
  double *A = bA.as<double>(), *inv = bInv.as<double>(), *r = bR.as<double>(), *scal = bSc.as<double>();
  GPXB_CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double) * 8, s));
  GPXB_TRY(hess_gram(ctx, kind, x, J, N, nv, x, J, N, nv, nf, ell, s2, A, n, 1, s));
  GPXB_TRY(potrf_lower(ctx, A, n, (int)n, inv, s));
  GPXB_TRY(logdet_lower(ctx, A, n, (int)n, scal + 1, 1.0, s));
  GPXB_CUDA_CHECK(cudaMemcpyAsync(r, y, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  GPXB_TRY(trsm_right_lt(ctx, A, n, inv, r, n, 1, (int)n, s));
  sk::sumsq_kernel<<<1, 1024, 0, s>>>(r, n, scal + 0, 0);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(trsm_right_l(ctx, A, n, inv, r, n, 1, (int)n, s));  // r <- alpha
  sk::sumsq_kernel<<<1, 1024, 0, s>>>(r, n, scal + 3, 0);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(bW.alloc(sizeof(double) * (size_t)n * n, s));
  GPXB_TRY(bT.alloc(sizeof(double) * potri_scratch_doubles((int)n), s));
  GPXB_TRY(potri_lower(ctx, A, n, inv, (int)n, bW.as<double>(), bT.as<double>(), A, n, s));
  bW.release();
  sk::diag_sum_kernel<<<1, 1024, 0, s>>>(A, n, (int)n, scal + 4);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(hess_run(ctx, kind, x, J, N, nv, x, J, N, nv, nf, ell, 0.0, nullptr, n, 1, A, n, r, scal + 2, s));
  finalize_derivs_grad_kernel<<<1, 1, 0, s>>>(scal, n, sigma, 1, out3);
  GPXB_LAUNCHED(ctx);
  return 0;
}

namespace {
// V[s] = (q_s, jaccoef[s][0..nf)),  q_s = jaccoef[s] . x[s]
__global__ void d0_pack_kernel(const double* __restrict__ x, const double* __restrict__ jc, int N, int nf,
                               double* __restrict__ V) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  double q = 0.0;
  for (int f = 0; f < nf; ++f) {
    const double c = jc[(long long)s * nf + f];
    V[(long long)s * (nf + 1) + 1 + f] = c;
    q += c * x[(long long)s * nf + f];
  }
  V[(long long)s * (nf + 1)] = q;
}
// mean[t] = mu - factor (W[t][0] - sum_f x[t][f] W[t][1+f])
__global__ void d0_finish_kernel(const double* __restrict__ W, const double* __restrict__ x, int ns, int nf,
                                 double factor, double mu, double* __restrict__ mean) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ns) return;
  const double* w = W + (long long)t * (nf + 1);
  double acc = w[0];
  for (int f = 0; f < nf; ++f) acc -= x[(long long)t * nf + f] * w[1 + f];
  mean[t] = mu - factor * acc;
}
}  // namespace

// mean[ns] = mu + sum_s d0kjc(x_train, x, jaccoef)[s, :]  (_predict_y_derivs_dense fast path,
// gpx/models/_gpr.py:630-638).  d k(x_s, x_t)/d x_s = -g(x_s, x_t) (x_s - x_t) with the radial profile
//   SE:  g = (2/l^2) exp(-d2)                              = (2/l^2) k_SE(l)
//   M52: g = (5/(3 l^2)) (1 + d) exp(-d), d = sqrt5 r/l    = (5/(3 l^2)) k_M32(l sqrt(3/5))
// so sum_s jc_s . dk/dx_s = -factor [ (G q)_t - x_t . (G JC)_t ]: ONE matrix-free block mat-vec
// G(x, x_train) [q | JC] with nf + 1 columns; no (N, n*, nf) temporary is formed.
int gpr_predict_y_derivs(Ctx* ctx, int kind, const double* x_train, const double* jaccoef, int N, int nf,
                         const double* x, int ns, double ell, double mu, double* mean_out, cudaStream_t s) {
  GPXB_ARG_CHECK(kind == KIND_SE || kind == KIND_M52, -50);
  GPXB_ARG_CHECK(N > 0 && nf > 0 && ns > 0, -51);
  const int P = nf + 1;
  DevBuf bV, bW;
  GPXB_TRY(bV.alloc(sizeof(double) * (size_t)N * P, s));
  GPXB_TRY(bW.alloc(sizeof(double) * (size_t)ns * P, s));
  d0_pack_kernel<<<ceil_div(N, 128), 128, 0, s>>>(x_train, jaccoef, N, nf, bV.as<double>());
  GPXB_LAUNCHED(ctx);
  const int pkind = kind == KIND_SE ? KIND_SE : KIND_M32;
  const double pell = kind == KIND_SE ? ell : ell * 0.7745966692414834;  // sqrt(3/5)
  const double factor = kind == KIND_SE ? 2.0 / (ell * ell) : 5.0 / (3.0 * ell * ell);
  if (kmatvec_supports(nf)) {
    GPXB_TRY(kmatvec(ctx, pkind, x, ns, -1, x_train, N, nf, pell, 0.0, bV.as<double>(), P, P, bW.as<double>(), P, s));
  } else {
    // wide features (the 90 coordinates of config 2): the profile block G(x, x_train) is stored in row
    // chunks (<= 1 GiB) by the fused Gram kernel and multiplied on the tensor pipe
    const ZLayout zl = z_layout(nf);
    GPXB_ARG_CHECK(zl.S <= GRAM_MAX_S, -52);
    const int chunk = (int)std::max<long long>(128, std::min<long long>(ns, (1LL << 27) / N / 128 * 128));
    DevBuf bZt, bZs, bG;
    GPXB_TRY(bZt.alloc(sizeof(double) * (size_t)N * zl.S, s));
    GPXB_TRY(bZs.alloc(sizeof(double) * (size_t)ns * zl.S, s));
    GPXB_TRY(bG.alloc(sizeof(double) * (size_t)chunk * N, s));
    GPXB_TRY(prep_z(ctx, x_train, nf, N, zl, pell, bZt.as<double>(), s));
    GPXB_TRY(prep_z(ctx, x, nf, ns, zl, pell, bZs.as<double>(), s));
    for (int r0 = 0; r0 < ns; r0 += chunk) {
      const int nr = std::min(chunk, ns - r0);
      GramArgs g;
      g.kind = pkind;
      g.Z1 = bZs.as<double>() + (size_t)r0 * zl.S; g.n1 = nr; g.Z2 = bZt.as<double>(); g.n2 = N; g.zl = zl;
      g.ell = pell; g.out = bG.as<double>(); g.ldo = N;
      GPXB_TRY(gram_launch(ctx, g, s));
      GemmArgs m;  // W[r0:r0+nr] = G [q | JC]
      m.M = nr; m.N = P; m.K = N;
      m.A = bG.as<double>(); m.lda = N; m.a_kmajor = true;
      m.B = bV.as<double>(); m.ldb = P; m.b_kmajor = false;
      m.C = bW.as<double>() + (size_t)r0 * P; m.ldc = P;
      GPXB_TRY(gemm_f64(ctx, m, s));
    }
  }
  d0_finish_kernel<<<ceil_div(ns, 128), 128, 0, s>>>(bW.as<double>(), x, ns, nf, factor, mu, mean_out);
  GPXB_LAUNCHED(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Matrix-free Hessian-kernel operator  W = (d01kj(x1, x2; J1, J2) [+ diag_add I]) V   (replaces
// _Ax_derivs_lhs_fun, gpx/models/_gpr.py:122-152: one kernel ROW of (N nv) entries per lax.scan step).
// With m_t = sum_u A2_(t,u) V_(t,u) (the contracted jacobian of V) the sum over the (t, u) columns collapses to
//   (H V)_(s,v) = A1_(s,v) . h_s,   h_s = sum_t [ c1_st m_t - kappa^2 c2_st (m_t . D_st) D_st ],  D_st = x1_s - x2_t,
// i.e. h = C1 M - (rowsum(E) o X1 - E X2) with E = kappa^2 C2 o (X1 M^T - 1 r^T), r_t = m_t . x2_t:
// three (N1 x N2 x nf) GEMMs on the configuration level instead of an (N1 nv)(N2 nv) nf product.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void rowdot_kernel(const double* __restrict__ A, const double* __restrict__ B, int n, int d,
                              double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0;
  for (int f = 0; f < d; ++f) a += A[(long long)i * d + f] * B[(long long)i * d + f];
  out[i] = a;
}
// R[s][t] <- kappa2 * C2[s][t] * (R[s][t] - r[t])
__global__ void hessop_e_kernel(double* __restrict__ R, const double* __restrict__ C2, const double* __restrict__ r,
                                long long n1, int n2, double kappa2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n1 * n2) return;
  R[idx] = kappa2 * C2[idx] * (R[idx] - r[idx % n2]);
}
// X2e[t] = (x2[t][0..nf), 1)
__global__ void append_one_kernel(const double* __restrict__ x, int n, int d, double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * (d + 1)) return;
  const int i = (int)(idx / (d + 1)), c = (int)(idx % (d + 1));
  out[idx] = c < d ? x[(long long)i * d + c] : 1.0;
}
// W[(s,v)] = sum_f J1[s][f][v] (H1[s][f] - (H2[s][nf] x1[s][f] - H2[s][f])) + diag_add * V[(s,v)]
__global__ void hessop_out_kernel(const double* __restrict__ J1, const double* __restrict__ x1,
                                  const double* __restrict__ H1, const double* __restrict__ H2, int N1, int nf, int nv,
                                  double diag_add, const double* __restrict__ V, long long ldv,
                                  double* __restrict__ W, long long ldw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N1 * nv) return;
  const int sidx = (int)(i / nv), v = (int)(i % nv);
  const double* Js = J1 + (long long)sidx * nf * nv + v;
  const double* h1 = H1 + (long long)sidx * nf;
  const double* h2 = H2 + (long long)sidx * (nf + 1);
  const double* xs = x1 + (long long)sidx * nf;
  const double rs = h2[nf];
  double a = 0.0;
  for (int f = 0; f < nf; ++f) a += Js[(long long)f * nv] * (h1[f] - (rs * xs[f] - h2[f]));
  if (diag_add != 0.0) a += diag_add * V[i * ldv];
  W[i * ldw] = a;
}
// strided jaccoef: M[t][f] = sum_u J[t][f][u] V[(t nv + u) ldv]
__global__ void jaccoef_strided_kernel(const double* __restrict__ V, long long ldv, const double* __restrict__ J, int N,
                                       int nf, int nv, double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * nf) return;
  const int t = (int)(idx / nf);
  const double* Jr = J + idx * nv;
  double a = 0.0;
  for (int u = 0; u < nv; ++u) a += V[((long long)t * nv + u) * ldv] * Jr[u];
  out[idx] = a;
}
}  // namespace

struct HessOp {
  int kind = 0, N1 = 0, nv1 = 0, N2 = 0, nv2 = 0, nf = 0;
  const double *x1 = nullptr, *J1 = nullptr, *x2 = nullptr, *J2 = nullptr;
  double ell = 1.0, kappa2 = 0.0, diag_add = 0.0;
  int dl = 0;  // 1: the operator is d(d01kj)/d lengthscale (same form, d/d ell pair scalars; no shift)
  DevBuf C1, C2, X2e, M, r, R, H1, H2;
  long long rows() const { return (long long)N1 * nv1; }
  long long cols() const { return (long long)N2 * nv2; }
};

int hess_op_prepare(Ctx* ctx, HessOp& op, cudaStream_t s) {
  GPXB_ARG_CHECK(op.kind == KIND_SE || op.kind == KIND_M52, -50);
  GPXB_ARG_CHECK(op.N1 > 0 && op.N2 > 0 && op.nv1 > 0 && op.nv2 > 0 && op.nf > 0, -51);
  const ZLayout zl = z_layout(op.nf);
  const bool sym = (op.x1 == op.x2) && (op.J1 == op.J2) && (op.N1 == op.N2) && (op.nv1 == op.nv2);
  const double kappa = (op.kind == KIND_SE) ? 2.0 / (op.ell * op.ell) : 2.23606797749979 / op.ell;
  op.kappa2 = kappa * kappa;
  const size_t nn = (size_t)op.N1 * op.N2;
  DevBuf bZ1, bZ2;
  GPXB_TRY(bZ1.alloc(sizeof(double) * (size_t)op.N1 * zl.S, s));
  GPXB_TRY(bZ2.alloc(sizeof(double) * (size_t)op.N2 * zl.S, s));
  GPXB_TRY(op.C1.alloc(sizeof(double) * nn, s));
  GPXB_TRY(op.C2.alloc(sizeof(double) * nn, s));
  GPXB_TRY(op.R.alloc(sizeof(double) * nn, s));
  GPXB_TRY(op.X2e.alloc(sizeof(double) * (size_t)op.N2 * (op.nf + 1), s));
  GPXB_TRY(op.M.alloc(sizeof(double) * (size_t)op.N2 * op.nf, s));
  GPXB_TRY(op.r.alloc(sizeof(double) * (size_t)op.N2, s));
  GPXB_TRY(op.H1.alloc(sizeof(double) * (size_t)op.N1 * op.nf, s));
  GPXB_TRY(op.H2.alloc(sizeof(double) * (size_t)op.N1 * (op.nf + 1), s));
  GPXB_TRY(prep_z(ctx, op.x1, op.nf, op.N1, zl, op.ell, bZ1.as<double>(), s));
  GPXB_TRY(prep_z(ctx, op.x2, op.nf, op.N2, zl, op.ell, bZ2.as<double>(), s));
  pair_coef_kernel<<<ceil_div((long long)nn, 256), 256, 0, s>>>(bZ1.as<double>(), op.N1, bZ2.as<double>(), op.N2, zl.S,
                                                                 zl.Dp, op.kind, op.ell, sym ? 1 : 0, op.dl,
                                                                 op.C1.as<double>(), op.C2.as<double>());
  GPXB_LAUNCHED(ctx);
  append_one_kernel<<<ceil_div((long long)op.N2 * (op.nf + 1), 256), 256, 0, s>>>(op.x2, op.N2, op.nf,
                                                                                    op.X2e.as<double>());
  GPXB_LAUNCHED(ctx);
  return 0;
}

// one column: W[i * ldw] = (H V)[i] + diag_add V[i] (diag_add only makes sense for the symmetric operator)
int hess_op_apply(Ctx* ctx, HessOp& op, const double* V, long long ldv, double* W, long long ldw, cudaStream_t s) {
  const int N1 = op.N1, N2 = op.N2, nf = op.nf;
  double *M = op.M.as<double>(), *R = op.R.as<double>();
  jaccoef_strided_kernel<<<ceil_div((long long)N2 * nf, 256), 256, 0, s>>>(V, ldv, op.J2, N2, nf, op.nv2, M);
  GPXB_LAUNCHED(ctx);
  rowdot_kernel<<<ceil_div(N2, 128), 128, 0, s>>>(M, op.x2, N2, nf, op.r.as<double>());
  GPXB_LAUNCHED(ctx);
  {
    GemmArgs g;  // R = X1 M^T
    g.M = N1; g.N = N2; g.K = nf;
    g.A = op.x1; g.lda = nf; g.a_kmajor = true;
    g.B = M; g.ldb = nf; g.b_kmajor = true;
    g.C = R; g.ldc = N2;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  hessop_e_kernel<<<ceil_div((long long)N1 * N2, 256), 256, 0, s>>>(R, op.C2.as<double>(), op.r.as<double>(), N1, N2,
                                                                     op.kappa2);
  GPXB_LAUNCHED(ctx);
  {
    GemmArgs g;  // H1 = C1 M
    g.M = N1; g.N = nf; g.K = N2;
    g.A = op.C1.as<double>(); g.lda = N2; g.a_kmajor = true;
    g.B = M; g.ldb = nf; g.b_kmajor = false;
    g.C = op.H1.as<double>(); g.ldc = nf;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  {
    GemmArgs g;  // H2 = E [X2 | 1]
    g.M = N1; g.N = nf + 1; g.K = N2;
    g.A = R; g.lda = N2; g.a_kmajor = true;
    g.B = op.X2e.as<double>(); g.ldb = nf + 1; g.b_kmajor = false;
    g.C = op.H2.as<double>(); g.ldc = nf + 1;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  hessop_out_kernel<<<ceil_div(op.rows(), 128), 128, 0, s>>>(op.J1, op.x1, op.H1.as<double>(), op.H2.as<double>(), N1, nf,
                                                            op.nv1, op.diag_add, V, ldv, W, ldw);
  GPXB_LAUNCHED(ctx);
  return 0;
}

int hess_matvec(Ctx* ctx, int kind, const double* x1, const double* J1, int N1, int nv1, const double* x2,
                const double* J2, int N2, int nv2, int nf, double ell, double diag_add, int dl, const double* V,
                long long ldv, int P, double* W, long long ldw, cudaStream_t s) {
  HessOp op;
  op.dl = dl;
  op.kind = kind; op.N1 = N1; op.nv1 = nv1; op.N2 = N2; op.nv2 = nv2; op.nf = nf;
  op.x1 = x1; op.J1 = J1; op.x2 = x2; op.J2 = J2; op.ell = ell; op.diag_add = diag_add;
  GPXB_ARG_CHECK(diag_add == 0.0 || op.rows() == op.cols(), -59);
  GPXB_TRY(hess_op_prepare(ctx, op, s));
  for (int p = 0; p < P; ++p) GPXB_TRY(hess_op_apply(ctx, op, V + p, ldv, W + p, ldw, s));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Iterative derivative-GPR: PCG fit with the rpcholesky_derivs preconditioner and the SLQ log-determinant
// on the matrix-free Hessian operator (_fit_derivs_iter gpx/models/_gpr.py:347-390, _precond_rpcholesky_derivs
// :219-256, rpcholesky_derivs gpx/kernels/approximations.py:132-242, _lml_derivs_iter _gpr.py:873-935).
// Replicas only: the operator costs O(N^2 nf) per product, there is nothing worth sharding.
// ------------------------------------------------------------------------------------------------
int precond_build(Ctx* ctx, const double* F, long long ldf, int k, int n_loc, double sigma, double* Pkk,
                  cudaStream_t s);
int pcg_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
             const double* b, const double* F, long long ldf, int k, const double* Pkk, double sigma, double tol,
             double atol, int maxiter, double* x, int* iters_out, cudaStream_t s);
int lanczos_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                 const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                 cudaStream_t s);

namespace {
// diag0[(s,v)] = d01kj(x_s, x_s, J_s, J_s)[v][v] = c1(d2 = 1e-12) |A_(s,v)|^2
__global__ void hess_diag_kernel(const double* __restrict__ AJ, int S, int nf, long long n, double c1_0,
                                 double* __restrict__ diag0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* a = AJ + i * S;
  double q = 0.0;
  for (int f = 0; f < nf; ++f) q += a[f] * a[f];
  diag0[i] = c1_0 * q;
}
// col[(t,u)] = d01kj(x_s1, x, J_s1, J)[v1][(t,u)] for the pivot row (s1, v1) = pbuf[0]; its A row is pbuf[2..]
__global__ void hess_column_kernel(const double* __restrict__ AJ, int S, const double* __restrict__ X, int nf, int nv,
                                   long long n, int kind, double ell, const double* __restrict__ pbuf,
                                   double* __restrict__ col) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const long long piv = (long long)pbuf[0];
  const long long s1 = piv / nv, t = r / nv;
  const double* a = pbuf + 2;
  const double* ar = AJ + r * S;
  const double* xt = X + t * nf;
  const double* xs = X + s1 * nf;
  double dd = 0.0, ara = 0.0, ard = 0.0, ad = 0.0;
  for (int f = 0; f < nf; ++f) {
    const double d = xt[f] - xs[f];
    dd += d * d;
    ara += ar[f] * a[f];
    ard += ar[f] * d;
    ad += a[f] * d;
  }
  double d2 = dd / (ell * ell) + 1e-12;
  if (t == s1) d2 = 1e-12;
  double c1, c2, kappa;
  if (kind == KIND_SE) {
    c2 = exp(-d2);
    kappa = 2.0 / (ell * ell);
    c1 = kappa * c2;
  } else {
    const double d = 2.23606797749979 * sqrt(fmax(d2, 1e-36));
    c2 = (5.0 / (3.0 * ell * ell)) * exp(-d);
    c1 = c2 * (1.0 + d);
    kappa = 2.23606797749979 / ell;
  }
  col[r] = c1 * ara - kappa * kappa * c2 * ard * ad;
}
}  // namespace

// rpcholesky_derivs (gpx/kernels/approximations.py:132-242): randomly pivoted Cholesky of the Hessian kernel
// over the (sample, variable) rows.  FB: row-blocked factor (rp_factor_doubles(N nv, M)); pivots: [M] int64.
int hess_rpcholesky(Ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, double ell,
                    const uint32_t key[2], int M, int partitionable, double* FB, long long* pivots,
                    cudaStream_t s) {
  GPXB_ARG_CHECK(kind == KIND_SE || kind == KIND_M52, -50);
  const long long n = (long long)N * nv;
  const ZLayout zl = z_layout(nf);
  DevBuf bAJ, bD0, bCol;
  GPXB_TRY(bAJ.alloc(sizeof(double) * (size_t)n * zl.S, s));
  GPXB_TRY(bD0.alloc(sizeof(double) * (size_t)n, s));
  GPXB_TRY(bCol.alloc(sizeof(double) * (size_t)n, s));
  double *AJ = bAJ.as<double>(), *col = bCol.as<double>();
  prep_hess_kernel<<<ceil_div(n, 128), 128, 0, s>>>(x, J, N, nf, nv, zl.S, zl.Dp, AJ);
  GPXB_LAUNCHED(ctx);
  double c1_0;
  if (kind == KIND_SE) c1_0 = (2.0 / (ell * ell)) * std::exp(-1e-12);
  else {
    const double d = 2.23606797749979 * 1e-6;
    c1_0 = (5.0 / (3.0 * ell * ell)) * std::exp(-d) * (1.0 + d);
  }
  hess_diag_kernel<<<ceil_div(n, 256), 256, 0, s>>>(AJ, zl.S, nf, n, c1_0, bD0.as<double>());
  GPXB_LAUNCHED(ctx);
  RpColumnGen gen = [&](const double* pbuf, cudaStream_t st) -> int {
    hess_column_kernel<<<ceil_div(n, 256), 256, 0, st>>>(AJ, zl.S, x, nf, nv, n, kind, ell, pbuf, col);
    GPXB_LAUNCHED(ctx);
    return 0;
  };
  return rpcholesky_ext(ctx, AJ, (int)n, zl, bD0.as<double>(), col, gen, key, M, partitionable, FB, pivots, s);
}

// c[n] = (d01kj + (sigma^2+1e-10) I)^-1 y by PCG (tol, atol as jax.scipy.sparse.linalg.cg); n_pivots = 0: no
// preconditioner.  lanczos_* (optional): also run the m-step block Lanczos on the same operator for the P
// start vectors U (n x ldu), as _lml_derivs_iter does after the fit.
int gpr_derivs_iter(Ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf, int nv,
                    double ell, double sigma, int n_pivots, const uint32_t key[2], int partitionable, double tol,
                    double atol, int maxiter, double* c_out, double* jaccoef_out, int* iters_out, const double* U,
                    long long ldu, int P, int m, double* alpha_out, double* beta_out, int* flag, cudaStream_t s) {
  GPXB_ARG_CHECK(ctx->world == 1, -61);  // replicas only
  const long long n = (long long)N * nv;
  GPXB_ARG_CHECK(n > 0 && n < (1LL << 31) && n_pivots >= 0, -1);
  const double s2 = sigma * sigma + 1e-10;
  HessOp op;
  op.kind = kind; op.N1 = N; op.nv1 = nv; op.N2 = N; op.nv2 = nv; op.nf = nf;
  op.x1 = x; op.J1 = J; op.x2 = x; op.J2 = J; op.ell = ell; op.diag_add = s2;
  GPXB_TRY(hess_op_prepare(ctx, op, s));
  if (c_out) {
    DevBuf bFB, bFT, bPiv, bPkk;
    const long long ldf = (n + 1) / 2 * 2;
    const int k = n_pivots;
    if (k > 0) {
      GPXB_TRY(bFB.alloc(sizeof(double) * std::max<size_t>(rp_factor_doubles((int)n, k), 8), s));
      GPXB_TRY(bFT.alloc(sizeof(double) * (size_t)k * ldf, s));
      GPXB_TRY(bPiv.alloc(sizeof(long long) * k, s));
      GPXB_TRY(bPkk.alloc(sizeof(double) * (size_t)k * k, s));
      GPXB_TRY(hess_rpcholesky(ctx, kind, x, J, N, nf, nv, ell, key, k, partitionable, bFB.as<double>(),
                               bPiv.as<long long>(), s));
      GPXB_TRY(rp_unblock_transposed(ctx, bFB.as<double>(), (int)n, k, bFT.as<double>(), ldf, s));
      bFB.release();
      GPXB_TRY(precond_build(ctx, bFT.as<double>(), ldf, k, (int)n, sigma, bPkk.as<double>(), s));
    }
    auto apply = [&](const double* p, double* Ap) -> int { return hess_op_apply(ctx, op, p, 2, Ap, 1, s); };
    GPXB_TRY(pcg_core(ctx, (int)n, n, apply, y, bFT.as<double>(), ldf, k, bPkk.as<double>(), sigma, tol, atol,
                      maxiter > 0 ? maxiter : (int)std::min<long long>(10 * n, 2000000000LL), c_out, iters_out, s));
    if (jaccoef_out) {
      jaccoef_kernel<<<ceil_div((long long)N * nf, 256), 256, 0, s>>>(c_out, J, N, nf, nv, jaccoef_out);
      GPXB_LAUNCHED(ctx);
    }
  }
  if (U && alpha_out) {
    GPXB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
    DevBuf up;
    for (int c0 = 0; c0 < P; c0 += 32) {
      const int pb = std::min(32, P - c0);
      const int PS = v_padded_stride(pb);
      GPXB_TRY(up.alloc(sizeof(double) * ((size_t)n * PS + 16), s));
      GPXB_TRY(pack_v(ctx, U, ldu, (int)n, c0, pb, PS, up.as<double>(), s));
      auto apply = [&](const double* v, double* w) -> int {
        for (int p = 0; p < pb; ++p) GPXB_TRY(hess_op_apply(ctx, op, v + p, PS, w + p, PS, s));
        return 0;
      };
      GPXB_TRY(lanczos_core(ctx, (int)n, n, apply, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                            beta_out + (size_t)c0 * m, flag, s));
      up.release();
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Iterative GPR on targets AND derivatives (gpx/models/gpr_td.py: _Ax_lhs_fun 162-310, rpcholesky_TD 313-460,
// _precond_rpcholesky_TD 463-504, _fit_iter 685-747, _lml_iter 562-637; one derivative block).
//         / K + s_t I      d1kj      \      rows / columns ordered [targets (N); derivatives (sample, variable)]
//     A = |                           |
//         \   d0kj     d01kj + s_d I /
// With C1_st the gradient profile (d k/d x2 = C1 (x1 - x2), d k/d x1 = -C1 (x1 - x2); HessOp's C1) and m_t the
// jacobian-contracted derivative part of the vector (HessOp's M), r_t = m_t . x_t:
//   (d1kj z2)_s     =  x_s . (C1 M)_s - (C1 r)_s
//   (d0kj z1)_(s,v) = -J_s[:, v] . (x_s (C1 z1)_s - (C1 (X o z1))_s)
// so one extra N x N x (nf + 2) GEMM with C1 and one GEMV with K ride on the three GEMMs of the Hessian operator.
// ------------------------------------------------------------------------------------------------
namespace {
// B[i] = (z1[i], x[i][:] * z1[i], r[i])      N x (nf + 2)
__global__ void td_build_b_kernel(const double* __restrict__ z1, long long ldv, const double* __restrict__ x,
                                  const double* __restrict__ r, int N, int nf, double* __restrict__ B) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = nf + 2;
  if (idx >= (long long)N * w) return;
  const int i = (int)(idx / w), c = (int)(idx % w);
  const double z = z1[(long long)i * ldv];
  B[idx] = c == 0 ? z : (c <= nf ? x[(long long)i * nf + c - 1] * z : r[i]);
}
// W1[i] = (K z1)[i] + s_t z1[i] + x_i . H1_i - T[i][nf + 1]
__global__ void td_out1_kernel(const double* __restrict__ t0, const double* __restrict__ z1, long long ldv,
                               const double* __restrict__ x, const double* __restrict__ H1,
                               const double* __restrict__ T, int N, int nf, double s_t, double* __restrict__ W,
                               long long ldw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double a = 0.0;
  for (int f = 0; f < nf; ++f) a += x[(long long)i * nf + f] * H1[(long long)i * nf + f];
  W[(long long)i * ldw] = t0[i] + s_t * z1[(long long)i * ldv] + a - T[(long long)i * (nf + 2) + nf + 1];
}
// W2[(i,u)] -= sum_f J[i][f][u] (x[i][f] T[i][0] - T[i][1 + f])
__global__ void td_out2_kernel(const double* __restrict__ J, const double* __restrict__ x,
                               const double* __restrict__ T, int N, int nf, int nv, double* __restrict__ W,
                               long long ldw) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)N * nv) return;
  const int i = (int)(r / nv), u = (int)(r % nv);
  const double* Ji = J + (long long)i * nf * nv + u;
  const double* Ti = T + (long long)i * (nf + 2);
  const double* xi = x + (long long)i * nf;
  double a = 0.0;
  for (int f = 0; f < nf; ++f) a += Ji[(long long)f * nv] * (xi[f] * Ti[0] - Ti[1 + f]);
  W[r * ldw] -= a;
}
// Row r of the REFERENCE's ordering [targets (N); block 1 (sample, variable) (N nv1); block 2 (N nv2)] ->
// (is_target, sample t, column u of the concatenated jacobian [J1 | J2], nv = nv1 + nv2).  nv2 = 0: one block.
struct TdRow {
  bool target;
  long long t;
  int u;
};
__device__ __forceinline__ TdRow td_row(long long r, int N, int nv1, int nv2) {
  TdRow o;
  o.target = r < N;
  if (o.target) { o.t = r; o.u = 0; return o; }
  long long q = r - N;
  if (q < (long long)N * nv1) { o.t = q / nv1; o.u = (int)(q % nv1); return o; }
  q -= (long long)N * nv1;
  o.t = q / nv2; o.u = nv1 + (int)(q % nv2);
  return o;
}
// rows of the pivot-buffer layout: zero for the targets, J_t[:, u] for the derivative rows
__global__ void td_rows_kernel(const double* __restrict__ J, int N, int nf, int nv1, int nv2, int S,
                               double* __restrict__ rows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = nv1 + nv2;
  const long long n = (long long)N + (long long)N * nv;
  if (idx >= n * S) return;
  const long long r = idx / S;
  const int f = (int)(idx % S);
  double v = 0.0;
  const TdRow w = td_row(r, N, nv1, nv2);
  if (!w.target && f < nf) v = J[w.t * nf * nv + (long long)f * nv + w.u];
  rows[idx] = v;
}
// diag0 = [k(x, x) (d2 = 1e-12); c1(1e-12) |J_t[:, u]|^2]
__global__ void td_diag_kernel(const double* __restrict__ J, int N, int nf, int nv1, int nv2, double k0, double c1_0,
                               double* __restrict__ diag0) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = nv1 + nv2;
  if (r >= (long long)N + (long long)N * nv) return;
  const TdRow w = td_row(r, N, nv1, nv2);
  if (w.target) {
    diag0[r] = k0;
    return;
  }
  const double* Jt = J + w.t * nf * nv + w.u;
  double a = 0.0;
  for (int f = 0; f < nf; ++f) a += Jt[(long long)f * nv] * Jt[(long long)f * nv];
  diag0[r] = c1_0 * a;
}
// column `piv` of the block matrix without the noise terms (rpcholesky_TD first_row / second_row / third_row,
// gpr_td.py:389-436), rows in the reference's order
__global__ void td_column_kernel(const double* __restrict__ x, const double* __restrict__ J, int N, int nf, int nv1,
                                 int nv2, int kind, double ell, const double* __restrict__ pbuf,
                                 double* __restrict__ col) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = nv1 + nv2;
  if (r >= (long long)N + (long long)N * nv) return;
  const TdRow pw = td_row((long long)pbuf[0], N, nv1, nv2), rw = td_row(r, N, nv1, nv2);
  const bool prow_t = pw.target, rrow_t = rw.target;
  const long long s1 = pw.t, t = rw.t;
  const double* xs = x + s1 * nf;
  const double* xt = x + t * nf;
  const double* a = J + s1 * nf * nv + pw.u;  // J_s1[:, v1] (stride nv), used when the pivot is a derivative row
  const double* b = J + t * nf * nv + rw.u;   // J_t[:, u]
  double dd = 0.0, ab = 0.0, ad = 0.0, bd = 0.0;
  for (int f = 0; f < nf; ++f) {
    const double d = xs[f] - xt[f];  // x_pivot - x_row
    dd += d * d;
    const double af = prow_t ? 0.0 : a[(long long)f * nv], bf = rrow_t ? 0.0 : b[(long long)f * nv];
    ab += af * bf;
    ad += af * d;
    bd += bf * d;
  }
  double d2 = dd / (ell * ell) + 1e-12;
  if (t == s1) d2 = 1e-12;
  double k, c1, c2, kappa;
  if (kind == KIND_SE) {
    k = exp(-d2);
    kappa = 2.0 / (ell * ell);
    c2 = k;
    c1 = kappa * k;
  } else {
    const double d = 2.23606797749979 * sqrt(fmax(d2, 1e-36));
    const double e = exp(-d);
    k = (1.0 + d + d * d / 3.0) * e;
    c2 = (5.0 / (3.0 * ell * ell)) * e;
    c1 = c2 * (1.0 + d);
    kappa = 2.23606797749979 / ell;
  }
  double out;
  if (prow_t && rrow_t) out = k;                       // k(x_s, x_t)
  else if (prow_t) out = c1 * bd;                      // d1kj: d k(x_s, x_t)/d x_t . J_t[:, u] = c1 (x_s - x_t) . b
  else if (rrow_t) out = -c1 * ad;                     // d0kj: d k(x_s1, x_t)/d x_s1 . a = -c1 (x_s1 - x_t) . a
  else out = c1 * ab - kappa * kappa * c2 * ad * bd;   // d01kj (hess_column_kernel)
  col[r] = out;
}
// FT_lib[c][l] = FT_ref[c][r(l)]: the factor's rows from the reference's order to the library's
// [targets; (sample, [block 1 | block 2])]
__global__ void td_permute_ft_kernel(const double* __restrict__ src, double* __restrict__ dst, long long ldf, int k,
                                     int N, int nv1, int nv2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = nv1 + nv2;
  const long long n = (long long)N + (long long)N * nv;
  if (idx >= n * k) return;
  const long long l = idx % n;
  const int c = (int)(idx / n);
  long long r = l;
  if (l >= N) {
    const long long q = l - N, t = q / nv;
    const int u = (int)(q % nv);
    r = u < nv1 ? N + t * nv1 + u : N + (long long)N * nv1 + t * nv2 + (u - nv1);
  }
  dst[(long long)c * ldf + l] = src[(long long)c * ldf + r];
}
// W2[(i,u)] += shift(u) V2[(i,u)]: sigma_derivs for the first nv1 variables, sigma_derivs2 for the rest
__global__ void td_shift2_kernel(double* __restrict__ W, long long ldw, const double* __restrict__ V, long long ldv,
                                 long long n, int nv, int nv1, double s1, double s2) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  W[r * ldw] += ((int)(r % nv) < nv1 ? s1 : s2) * V[r * ldv];
}
}  // namespace

struct TdOp {
  HessOp h;
  int N = 0, nf = 0, nv = 0, nv1 = 0;  // nv1 < nv: a second derivative block with its own noise
  const double *x = nullptr, *J = nullptr;
  double s_t = 0.0, s_d = 0.0, s_d2 = 0.0;
  DevBuf C0, B, T, t0;
  long long rows() const { return (long long)N + (long long)N * nv; }
};

int td_op_prepare(Ctx* ctx, TdOp& op, int kind, const double* x, const double* J, int N, int nf, int nv, int nv1,
                  double ell, double s_t, double s_d, double s_d2, cudaStream_t s) {
  op.N = N; op.nf = nf; op.nv = nv; op.nv1 = nv1; op.x = x; op.J = J; op.s_t = s_t; op.s_d = s_d; op.s_d2 = s_d2;
  op.h.kind = kind; op.h.N1 = op.h.N2 = N; op.h.nv1 = op.h.nv2 = nv; op.h.nf = nf;
  op.h.x1 = op.h.x2 = x; op.h.J1 = op.h.J2 = J; op.h.ell = ell; op.h.diag_add = nv1 < nv ? 0.0 : s_d;
  GPXB_TRY(hess_op_prepare(ctx, op.h, s));
  const ZLayout zl = z_layout(nf);
  DevBuf bZ;
  GPXB_TRY(bZ.alloc(sizeof(double) * (size_t)N * zl.S, s));
  GPXB_TRY(op.C0.alloc(sizeof(double) * (size_t)N * N, s));
  GPXB_TRY(op.B.alloc(sizeof(double) * (size_t)N * (nf + 2), s));
  GPXB_TRY(op.T.alloc(sizeof(double) * (size_t)N * (nf + 2), s));
  GPXB_TRY(op.t0.alloc(sizeof(double) * (size_t)N, s));
  GPXB_TRY(prep_z(ctx, x, nf, N, zl, ell, bZ.as<double>(), s));
  GramArgs g;
  g.kind = kind; g.Z1 = bZ.as<double>(); g.n1 = N; g.Z2 = bZ.as<double>(); g.n2 = N; g.zl = zl; g.ell = ell;
  g.sym = 1; g.out = op.C0.as<double>(); g.ldo = N;
  return gram_launch(ctx, g, s);
}

// one column: W = A V, V / W of length N + N nv with strides ldv / ldw
int td_op_apply(Ctx* ctx, TdOp& op, const double* V, long long ldv, double* W, long long ldw, cudaStream_t s) {
  const int N = op.N, nf = op.nf, nv = op.nv;
  const double* z2 = V + (long long)N * ldv;
  double* W2 = W + (long long)N * ldw;
  GPXB_TRY(hess_op_apply(ctx, op.h, z2, ldv, W2, ldw, s));  // W2 = (d01kj + s_d) z2; leaves M, r, H1 = C1 M
  if (op.nv1 < nv) {  // two derivative blocks: the noise depends on the variable
    const long long n2 = (long long)N * nv;
    td_shift2_kernel<<<ceil_div(n2, 256), 256, 0, s>>>(W2, ldw, z2, ldv, n2, nv, op.nv1, op.s_d, op.s_d2);
    GPXB_LAUNCHED(ctx);
  }
  double *B = op.B.as<double>(), *T = op.T.as<double>(), *t0 = op.t0.as<double>();
  td_build_b_kernel<<<ceil_div((long long)N * (nf + 2), 256), 256, 0, s>>>(V, ldv, op.x, op.h.r.as<double>(), N, nf, B);
  GPXB_LAUNCHED(ctx);
  {
    GemmArgs g;  // T = C1 [z1 | X o z1 | r]
    g.M = N; g.N = nf + 2; g.K = N;
    g.A = op.h.C1.as<double>(); g.lda = N; g.a_kmajor = true;
    g.B = B; g.ldb = nf + 2; g.b_kmajor = false;
    g.C = T; g.ldc = nf + 2;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  {
    GemmArgs g;  // t0 = K z1
    g.M = N; g.N = 1; g.K = N;
    g.A = op.C0.as<double>(); g.lda = N; g.a_kmajor = true;
    g.B = V; g.ldb = ldv; g.b_kmajor = false;
    g.C = t0; g.ldc = 1;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  td_out1_kernel<<<ceil_div(N, 128), 128, 0, s>>>(t0, V, ldv, op.x, op.h.H1.as<double>(), T, N, nf, op.s_t, W, ldw);
  GPXB_LAUNCHED(ctx);
  td_out2_kernel<<<ceil_div((long long)N * nv, 128), 128, 0, s>>>(op.J, op.x, T, N, nf, nv, W2, ldw);
  GPXB_LAUNCHED(ctx);
  return 0;
}

// rpcholesky_TD (gpx/models/gpr_td.py:313-460): FB row-blocked factor over the N + N nv rows, pivots [M] int64
int td_rpcholesky(Ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, int nv1, double ell,
                  const uint32_t key[2], int M, int partitionable, double* FB, long long* pivots, cudaStream_t s) {
  GPXB_ARG_CHECK(kind == KIND_SE || kind == KIND_M52, -50);
  GPXB_ARG_CHECK(nv1 >= 1 && nv1 <= nv, -51);
  const int nv2 = nv - nv1;
  const long long n = (long long)N + (long long)N * nv;
  GPXB_ARG_CHECK(n < (1LL << 31), -2);
  const ZLayout zl = z_layout(nf);
  DevBuf bRows, bD0, bCol;
  GPXB_TRY(bRows.alloc(sizeof(double) * (size_t)n * zl.S, s));
  GPXB_TRY(bD0.alloc(sizeof(double) * (size_t)n, s));
  GPXB_TRY(bCol.alloc(sizeof(double) * (size_t)n, s));
  double* col = bCol.as<double>();
  td_rows_kernel<<<ceil_div(n * zl.S, 256), 256, 0, s>>>(J, N, nf, nv1, nv2, zl.S, bRows.as<double>());
  GPXB_LAUNCHED(ctx);
  double k0, c1_0;
  if (kind == KIND_SE) {
    k0 = std::exp(-1e-12);
    c1_0 = (2.0 / (ell * ell)) * k0;
  } else {
    const double d = 2.23606797749979 * 1e-6;
    k0 = (1.0 + d + d * d / 3.0) * std::exp(-d);
    c1_0 = (5.0 / (3.0 * ell * ell)) * std::exp(-d) * (1.0 + d);
  }
  td_diag_kernel<<<ceil_div(n, 256), 256, 0, s>>>(J, N, nf, nv1, nv2, k0, c1_0, bD0.as<double>());
  GPXB_LAUNCHED(ctx);
  RpColumnGen gen = [&](const double* pbuf, cudaStream_t st) -> int {
    td_column_kernel<<<ceil_div(n, 256), 256, 0, st>>>(x, J, N, nf, nv1, nv2, kind, ell, pbuf, col);
    GPXB_LAUNCHED(ctx);
    return 0;
  };
  return rpcholesky_ext(ctx, bRows.as<double>(), (int)n, zl, bD0.as<double>(), col, gen, key, M, partitionable, FB,
                        pivots, s);
}

namespace {
// b = [y - mu; y_derivs]
__global__ void td_rhs_kernel(const double* __restrict__ y, const double* __restrict__ yd, const double* __restrict__ mu,
                              int N, long long n, double* __restrict__ b) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  b[r] = r < N ? y[r] - mu[0] : yd[r - N];
}
}  // namespace

// c = A^-1 [y - mu; y_derivs] by PCG with the rpcholesky_TD preconditioner (which uses sigma_targets for every row,
// gpr_td.py:478-503) and / or the m-step block Lanczos on A for the start vectors U (SLQ part of _lml_iter).
int gpr_td_iter(Ctx* ctx, int kind, const double* x, const double* J, const double* y, const double* yd, int N, int nf,
                int nv, int nv1, double ell, double sigma_t, double sigma_d, double sigma_d2, int mean_mode, int n_pivots,
                const uint32_t key[2],
                int partitionable, double tol, double atol, int maxiter, double* c_out, double* mu_out, int* iters_out,
                const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out, int* flag,
                cudaStream_t s) {
  GPXB_ARG_CHECK(ctx->world == 1, -61);  // replicas only, like the derivative-only iterative path
  const long long n = (long long)N + (long long)N * nv;
  GPXB_ARG_CHECK(N > 0 && n < (1LL << 31) && n_pivots >= 0, -1);
  TdOp op;
  GPXB_ARG_CHECK(nv1 >= 1 && nv1 <= nv, -51);
  GPXB_TRY(td_op_prepare(ctx, op, kind, x, J, N, nf, nv, nv1, ell, sigma_t * sigma_t + 1e-10, sigma_d * sigma_d + 1e-10,
                         sigma_d2 * sigma_d2 + 1e-10, s));
  if (c_out) {
    DevBuf bFB, bFT, bPiv, bPkk, bB;
    const long long ldf = (n + 1) / 2 * 2;
    const int k = n_pivots;
    GPXB_TRY(bB.alloc(sizeof(double) * (size_t)n, s));
    sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu_out);
    GPXB_LAUNCHED(ctx);
    td_rhs_kernel<<<ceil_div(n, 256), 256, 0, s>>>(y, yd, mu_out, N, n, bB.as<double>());
    GPXB_LAUNCHED(ctx);
    if (k > 0) {
      GPXB_TRY(bFB.alloc(sizeof(double) * std::max<size_t>(rp_factor_doubles((int)n, k), 8), s));
      GPXB_TRY(bFT.alloc(sizeof(double) * (size_t)k * ldf, s));
      GPXB_TRY(bPiv.alloc(sizeof(long long) * k, s));
      GPXB_TRY(bPkk.alloc(sizeof(double) * (size_t)k * k, s));
      GPXB_TRY(td_rpcholesky(ctx, kind, x, J, N, nf, nv, nv1, ell, key, k, partitionable, bFB.as<double>(),
                             bPiv.as<long long>(), s));
      GPXB_TRY(rp_unblock_transposed(ctx, bFB.as<double>(), (int)n, k, bFT.as<double>(), ldf, s));
      bFB.release();
      if (nv1 < nv) {  // the factor's rows follow the reference's [targets; block 1; block 2]: bring them to the operator's order
        DevBuf bTmp;
        GPXB_TRY(bTmp.alloc(sizeof(double) * (size_t)k * ldf, s));
        GPXB_CUDA_CHECK(cudaMemcpyAsync(bTmp.as<double>(), bFT.as<double>(), sizeof(double) * (size_t)k * ldf,
                                        cudaMemcpyDeviceToDevice, s));
        td_permute_ft_kernel<<<ceil_div(n * k, 256), 256, 0, s>>>(bTmp.as<double>(), bFT.as<double>(), ldf, k, N, nv1,
                                                                  nv - nv1);
        GPXB_LAUNCHED(ctx);
      }
      GPXB_TRY(precond_build(ctx, bFT.as<double>(), ldf, k, (int)n, sigma_t, bPkk.as<double>(), s));
    }
    auto apply = [&](const double* p, double* Ap) -> int { return td_op_apply(ctx, op, p, 2, Ap, 1, s); };
    GPXB_TRY(pcg_core(ctx, (int)n, n, apply, bB.as<double>(), bFT.as<double>(), ldf, k, bPkk.as<double>(), sigma_t,
                      tol, atol, maxiter > 0 ? maxiter : (int)std::min<long long>(10 * n, 2000000000LL), c_out,
                      iters_out, s));
  }
  if (U && alpha_out) {
    GPXB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
    DevBuf up;
    for (int c0 = 0; c0 < P; c0 += 32) {
      const int pb = std::min(32, P - c0);
      const int PS = v_padded_stride(pb);
      GPXB_TRY(up.alloc(sizeof(double) * ((size_t)n * PS + 16), s));
      GPXB_TRY(pack_v(ctx, U, ldu, (int)n, c0, pb, PS, up.as<double>(), s));
      auto apply = [&](const double* v, double* w) -> int {
        for (int p = 0; p < pb; ++p) GPXB_TRY(td_op_apply(ctx, op, v + p, PS, w + p, PS, s));
        return 0;
      };
      GPXB_TRY(lanczos_core(ctx, (int)n, n, apply, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                            beta_out + (size_t)c0 * m, flag, s));
      up.release();
    }
  }
  return 0;
}

namespace {
// out[i] = a * w[i] + u[i] + eps * p[2 i]
__global__ void sgpr_derivs_lin_kernel(double* __restrict__ out, double a, const double* __restrict__ w,
                                       const double* __restrict__ u, double eps, const double* __restrict__ p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * w[i] + u[i] + eps * p[2 * i];
}
}  // namespace

// SGPR trained on derivative values, CG fit (_fit_derivs_iter gpx/models/_sgpr.py:399-428, _Ax_derivs_lhs_fun
// :145-201): c = cg(sigma^2 H_mm + H_mn H_nm + 1e-10 I, H_mn y), H = d01kj, zero mean, no preconditioner; every
// product goes through the matrix-free Hessian operator (three operators: mm, mn, nm).
int sgpr_fit_derivs_iter(Ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf, int nv,
                         const double* x_locs, const double* J_locs, int M, int nvl, double ell, double sigma,
                         double tol, double atol, int maxiter, double* c_out, int* iters_out, cudaStream_t s) {
  GPXB_ARG_CHECK(ctx->world == 1, -61);
  const long long nm = (long long)M * nvl, nn = (long long)N * nv;
  GPXB_ARG_CHECK(nm > 0 && nn > 0 && nm < (1LL << 31) && nn < (1LL << 31), -1);
  HessOp mm, mn, nmop;
  auto setup = [&](HessOp& op, const double* x1, const double* J1, int N1, int nv1, const double* x2, const double* J2,
                   int N2, int nv2) {
    op.kind = kind; op.nf = nf; op.ell = ell; op.diag_add = 0.0;
    op.x1 = x1; op.J1 = J1; op.N1 = N1; op.nv1 = nv1; op.x2 = x2; op.J2 = J2; op.N2 = N2; op.nv2 = nv2;
  };
  setup(mm, x_locs, J_locs, M, nvl, x_locs, J_locs, M, nvl);
  setup(mn, x_locs, J_locs, M, nvl, x, J, N, nv);
  setup(nmop, x, J, N, nv, x_locs, J_locs, M, nvl);
  GPXB_TRY(hess_op_prepare(ctx, mm, s));
  GPXB_TRY(hess_op_prepare(ctx, mn, s));
  GPXB_TRY(hess_op_prepare(ctx, nmop, s));
  DevBuf bB, bT, bU, bW;
  GPXB_TRY(bB.alloc(sizeof(double) * (size_t)nm, s));
  GPXB_TRY(bT.alloc(sizeof(double) * (size_t)nn, s));
  GPXB_TRY(bU.alloc(sizeof(double) * (size_t)nm, s));
  GPXB_TRY(bW.alloc(sizeof(double) * (size_t)nm, s));
  double *b = bB.as<double>(), *t = bT.as<double>(), *u = bU.as<double>(), *w = bW.as<double>();
  GPXB_TRY(hess_op_apply(ctx, mn, y, 1, b, 1, s));  // rhs = H_mn y
  const double s2 = sigma * sigma;
  auto apply = [&](const double* p, double* Ap) -> int {
    GPXB_TRY(hess_op_apply(ctx, nmop, p, 2, t, 1, s));  // t = H_nm p
    GPXB_TRY(hess_op_apply(ctx, mn, t, 1, u, 1, s));    // u = H_mn t
    GPXB_TRY(hess_op_apply(ctx, mm, p, 2, w, 1, s));    // w = H_mm p
    sgpr_derivs_lin_kernel<<<ceil_div(nm, 256), 256, 0, s>>>(Ap, s2, w, u, 1e-10, p, nm);
    GPXB_LAUNCHED(ctx);
    return 0;
  };
  return pcg_core(ctx, (int)nm, nm, apply, b, nullptr, 0, 0, nullptr, sigma, tol, atol,
                  maxiter > 0 ? maxiter : (int)std::min<long long>(10 * nm, 2000000000LL), c_out, iters_out, s);
}

int lanczos_grad_core(Ctx* ctx, int n_loc, long long N,
                      const std::function<int(const double*, double*, bool)>& matvec, double sigma, const double* U,
                      int P, int PS, int m, double* alpha_out, double* beta_out, double* dalpha_out,
                      double* dbeta_out, int Ptot, int* flag, cudaStream_t s);

// Block Lanczos on d01kj + (sigma^2+1e-10) I with its forward-mode derivative w.r.t. (lengthscale, sigma):
// the SLQ part of value_and_grad of _lml_derivs_iter (gpx/models/_gpr.py:873-935).
int lanczos_grad_derivs(Ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, double ell,
                        double sigma, const double* U, long long ldu, int P, int m, double* alpha_out,
                        double* beta_out, double* dalpha_out, double* dbeta_out, int* flag, cudaStream_t s) {
  GPXB_ARG_CHECK(ctx->world == 1, -61);
  const long long n = (long long)N * nv;
  GPXB_ARG_CHECK(n > 0 && n < (1LL << 31), -1);
  HessOp op, dop;
  op.kind = dop.kind = kind; op.N1 = op.N2 = dop.N1 = dop.N2 = N; op.nv1 = op.nv2 = dop.nv1 = dop.nv2 = nv;
  op.nf = dop.nf = nf; op.x1 = op.x2 = dop.x1 = dop.x2 = x; op.J1 = op.J2 = dop.J1 = dop.J2 = J;
  op.ell = dop.ell = ell; op.diag_add = sigma * sigma + 1e-10; dop.dl = 1;
  GPXB_TRY(hess_op_prepare(ctx, op, s));
  GPXB_TRY(hess_op_prepare(ctx, dop, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
  DevBuf up;
  for (int c0 = 0; c0 < P; c0 += 32) {
    const int pb = std::min(32, P - c0);
    const int PS = v_padded_stride(pb);
    GPXB_TRY(up.alloc(sizeof(double) * ((size_t)n * PS + 16), s));
    GPXB_TRY(pack_v(ctx, U, ldu, (int)n, c0, pb, PS, up.as<double>(), s));
    auto matvec = [&](const double* v, double* w, bool dl) -> int {
      for (int p = 0; p < pb; ++p) GPXB_TRY(hess_op_apply(ctx, dl ? dop : op, v + p, PS, w + p, PS, s));
      return 0;
    };
    GPXB_TRY(lanczos_grad_core(ctx, (int)n, n, matvec, sigma, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                               beta_out + (size_t)c0 * m, dalpha_out + (size_t)c0 * m, dbeta_out + (size_t)c0 * m, P,
                               flag, s));
    up.release();
  }
  return 0;
}

namespace {
__global__ void add_const_kernel(double* __restrict__ v, long long n, double c) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += c;
}
// d0kj(x1, x2, J1)[i][t] = -g_st (q_i - (A1 x2_t^T)[i]),  i = s nv + v, stored at out[i * ld_i + t * ld_t]
__global__ void d0kj_kernel(const double* __restrict__ A1, int S, int Dp, const double* __restrict__ P12,
                            const double* __restrict__ G, int nv, long long n, int ns, double* __restrict__ out,
                            long long ld_i, long long ld_t) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ns) return;
  const long long i = idx / ns;
  const int t = (int)(idx % ns);
  const long long sidx = i / nv;
  out[i * ld_i + t * ld_t] = -G[sidx * ns + t] * (A1[i * S + Dp] - P12[idx]);
}
// partials[cta] = sum over this CTA's (i, t) of (a_i b_t - Y[i * ldy + t]) * d0kj-like entry (G: d/d ell scalars)
__global__ void __launch_bounds__(256) d0kj_contract_kernel(const double* __restrict__ A1, int S, int Dp,
                                                            const double* __restrict__ P12,
                                                            const double* __restrict__ G, int nv, long long n, int ns,
                                                            const double* __restrict__ a, const double* __restrict__ b,
                                                            const double* __restrict__ Y, long long ldy,
                                                            double* __restrict__ partials) {
  __shared__ double red[8];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (idx < n * ns) {
    const long long i = idx / ns;
    const int t = (int)(idx % ns);
    const double e = -G[(i / nv) * ns + t] * (A1[i * S + Dp] - P12[idx]);
    v = (a[i] * b[t] - Y[i * ldy + t]) * e;
  }
  v = block_sum<256>(v, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

// Shared tail of the two full-covariance predictions: Knm (m x n) holds the cross block (rows: test).
//   mean = mu + Knm c;  cov (already holding the prior block K_**) -= (Knm L^-T)(Knm L^-T)^T,
//   L = chol(d01kj(x_train) + shift I).
int predict_cov_tail(Ctx* ctx, int kind, const double* xt, const double* Jt, int N, int nf, int nv, double ell,
                     double shift, double* Knm, long long n, int m, const double* c, double mu, double* mean_out,
                     double* cov_out, cudaStream_t s) {
  {
    GemmArgs g;  // mean = Knm c
    g.M = m; g.N = 1; g.K = (int)n;
    g.A = Knm; g.lda = n; g.a_kmajor = true;
    g.B = c; g.ldb = 1; g.b_kmajor = false;
    g.C = mean_out; g.ldc = 1;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  if (mu != 0.0) {
    add_const_kernel<<<ceil_div(m, 256), 256, 0, s>>>(mean_out, m, mu);
    GPXB_LAUNCHED(ctx);
  }
  if (!cov_out) return 0;
  DevBuf bA, bInv;
  GPXB_TRY(bA.alloc(sizeof(double) * (size_t)n * n, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles((int)n), s));
  double *A = bA.as<double>(), *inv = bInv.as<double>();
  GPXB_TRY(hess_gram(ctx, kind, xt, Jt, N, nv, xt, Jt, N, nv, nf, ell, shift, A, n, 1, s));
  GPXB_TRY(potrf_lower(ctx, A, n, (int)n, inv, s));
  GPXB_TRY(trsm_right_lt(ctx, A, n, inv, Knm, n, m, (int)n, s));  // Knm <- G^T = Knm L^-T
  GemmArgs g;  // cov -= G^T G
  g.M = m; g.N = m; g.K = (int)n;
  g.A = Knm; g.lda = n; g.a_kmajor = true;
  g.B = Knm; g.ldb = n; g.b_kmajor = true;
  g.C = cov_out; g.ldc = m;
  g.alpha = -1.0; g.beta = 1.0;
  return gemm_f64(ctx, g, s);
}
}  // namespace

// _predict_derivs_dense without the contracted jacobian (gpx/models/_gpr.py:488-560): mean[ns nvs] = mu +
// K_nm c with K = d01kj; cov_out (optional, (ns nvs)^2) = d01kj(x, x) - G^T G, G = chol(d01kj(x_train) +
// (sigma^2 + 1e-10) I)^-1 K_mn.
int gpr_predict_derivs_full(Ctx* ctx, int kind, const double* xt, const double* Jt, int N, int nf, int nv,
                            const double* x, const double* J, int ns, int nvs, const double* c, double mu,
                            double ell, double sigma, double* mean_out, double* cov_out, cudaStream_t s) {
  const long long n = (long long)N * nv, m = (long long)ns * nvs;
  GPXB_ARG_CHECK(n > 0 && m > 0 && n < (1LL << 31) && m < (1LL << 31), -1);
  DevBuf bK;
  GPXB_TRY(bK.alloc(sizeof(double) * (size_t)m * n, s));
  GPXB_TRY(hess_gram(ctx, kind, x, J, ns, nvs, xt, Jt, N, nv, nf, ell, 0.0, bK.as<double>(), n, 0, s));
  if (cov_out) GPXB_TRY(hess_gram(ctx, kind, x, J, ns, nvs, x, J, ns, nvs, nf, ell, 0.0, cov_out, m, 0, s));
  return predict_cov_tail(ctx, kind, xt, Jt, N, nf, nv, ell, sigma * sigma + 1e-10, bK.as<double>(), n, (int)m, c, mu,
                          mean_out, cov_out, s);
}

// Block d0kj(x1, x2, J1) ((N1 nv) x N2; kernels.py:1108-1149 / grad_kernelize(argnums=0, with_jacob)) written at
// out[i * ld_i + t * ld_t] -- or, when out == nullptr, its d/d lengthscale (dl = 1) / value contracted with
// the weights (a_i b_t - Y[i][t]) into *contract_out (deterministic).  d1kj(x2, x1, J1) is its transpose.
int d0kj_block(Ctx* ctx, int kind, const double* x1, const double* J1, int N1, int nv, const double* x2, int N2,
               int nf, double ell, int dl, double* out, long long ld_i, long long ld_t, const double* a,
               const double* b, const double* Y, long long ldy, double* contract_out, cudaStream_t s) {
  GPXB_ARG_CHECK(kind == KIND_SE || kind == KIND_M52, -50);
  const long long n = (long long)N1 * nv;
  GPXB_ARG_CHECK(n > 0 && N2 > 0 && n < (1LL << 31), -1);
  const ZLayout zl = z_layout(nf);
  GPXB_ARG_CHECK(zl.S <= GRAM_MAX_S, -52);
  DevBuf bA1, bZ1, bZ2, bX2, bG, bG2, bP, bPart;
  GPXB_TRY(bA1.alloc(sizeof(double) * (size_t)n * zl.S, s));
  GPXB_TRY(bZ1.alloc(sizeof(double) * (size_t)N1 * zl.S, s));
  GPXB_TRY(bZ2.alloc(sizeof(double) * (size_t)N2 * zl.S, s));
  GPXB_TRY(bX2.alloc(sizeof(double) * (size_t)N2 * zl.S, s));
  GPXB_TRY(bG.alloc(sizeof(double) * (size_t)N1 * N2, s));
  GPXB_TRY(bG2.alloc(sizeof(double) * (size_t)N1 * N2, s));
  GPXB_TRY(bP.alloc(sizeof(double) * (size_t)n * N2, s));
  prep_hess_kernel<<<ceil_div(n, 128), 128, 0, s>>>(x1, J1, N1, nf, nv, zl.S, zl.Dp, bA1.as<double>());
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(prep_z(ctx, x1, nf, N1, zl, ell, bZ1.as<double>(), s));
  GPXB_TRY(prep_z(ctx, x2, nf, N2, zl, ell, bZ2.as<double>(), s));
  GPXB_TRY(prep_z(ctx, x2, nf, N2, zl, 1.0, bX2.as<double>(), s));
  // c1 of the Hessian kernel is exactly the radial derivative profile g: SE (2/l^2) exp(-d2), M52 (5/(3 l^2)) (1+d)
  // exp(-d); with dl = 1 the kernel yields dg/d ell.  sym: exact d2 = 1e-12 on coincident samples of one set.
  const int sym = (x1 == x2 && N1 == N2) ? 1 : 0;
  pair_coef_kernel<<<ceil_div((long long)N1 * N2, 256), 256, 0, s>>>(bZ1.as<double>(), N1, bZ2.as<double>(), N2, zl.S,
                                                                      zl.Dp, kind, ell, sym, dl, bG.as<double>(),
                                                                      bG2.as<double>());
  GPXB_LAUNCHED(ctx);
  {
    GemmArgs g;  // P12 = A1 X2^T   [n x N2]
    g.M = (int)n; g.N = N2; g.K = zl.Dp;
    g.A = bA1.as<double>(); g.lda = zl.S; g.a_kmajor = true;
    g.B = bX2.as<double>(); g.ldb = zl.S; g.b_kmajor = true;
    g.C = bP.as<double>(); g.ldc = N2;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  const long long tot = n * N2;
  if (out) {
    d0kj_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(bA1.as<double>(), zl.S, zl.Dp, bP.as<double>(), bG.as<double>(), nv, n,
                                                   N2, out, ld_i, ld_t);
    GPXB_LAUNCHED(ctx);
    return 0;
  }
  GPXB_ARG_CHECK(a && b && Y && contract_out, -57);
  const long long nb = ceil_div(tot, 256);
  GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)nb, s));
  d0kj_contract_kernel<<<(unsigned)nb, 256, 0, s>>>(bA1.as<double>(), zl.S, zl.Dp, bP.as<double>(), bG.as<double>(), nv,
                                                    n, N2, a, b, Y, ldy, bPart.as<double>());
  GPXB_LAUNCHED(ctx);
  sk::sum_kernel<<<1, 1024, 0, s>>>(bPart.as<double>(), nb, contract_out);
  GPXB_LAUNCHED(ctx);
  return 0;
}

// _predict_y_derivs_dense without the contracted jacobian (gpx/models/_gpr.py:608-660): mean[ns] = mu +
// K_nm c with K_mn = d0kj(x_train, x, J_train); cov_out (optional, ns^2) = k(x, x) - G^T G,
// G = chol(d01kj(x_train) + sigma^2 I)^-1 K_mn (no 1e-10 here, as in the reference).
int gpr_predict_y_derivs_full(Ctx* ctx, int kind, const double* xt, const double* Jt, int N, int nf, int nv,
                              const double* x, int ns, const double* c, double mu, double ell, double sigma,
                              double* mean_out, double* cov_out, cudaStream_t s) {
  GPXB_ARG_CHECK(kind == KIND_SE || kind == KIND_M52, -50);
  const long long n = (long long)N * nv;
  GPXB_ARG_CHECK(n > 0 && ns > 0 && n < (1LL << 31), -1);
  const ZLayout zl = z_layout(nf);
  GPXB_ARG_CHECK(zl.S <= GRAM_MAX_S, -52);
  DevBuf bK, bZ2;
  GPXB_TRY(bK.alloc(sizeof(double) * (size_t)ns * n, s));
  GPXB_TRY(bZ2.alloc(sizeof(double) * (size_t)ns * zl.S, s));
  GPXB_TRY(prep_z(ctx, x, nf, ns, zl, ell, bZ2.as<double>(), s));
  // cross block stored transposed: Knm[t][i]
  GPXB_TRY(d0kj_block(ctx, kind, xt, Jt, N, nv, x, ns, nf, ell, 0, bK.as<double>(), 1, n, nullptr, nullptr, nullptr, 0,
                      nullptr, s));
  if (cov_out) {
    GramArgs h;  // prior block k(x, x)
    h.kind = kind; h.Z1 = bZ2.as<double>(); h.n1 = ns; h.Z2 = bZ2.as<double>(); h.n2 = ns; h.zl = zl; h.ell = ell;
    h.sym = 1; h.out = cov_out; h.ldo = ns;
    GPXB_TRY(gram_launch(ctx, h, s));
  }
  return predict_cov_tail(ctx, kind, xt, Jt, N, nf, nv, ell, sigma * sigma, bK.as<double>(), n, ns, c, mu, mean_out,
                          cov_out, s);
}

}  // namespace gpxb

using gpxb::Ctx;

extern "C" {

int gpxb_hess_gram(gpxb_ctx* ctx, int kind, const double* x1, const double* J1, int n1, int nv1,
                   const double* x2, const double* J2, int n2, int nv2, int nf, double ell, double diag_add,
                   double* out, long long ldo, int lower_only, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x1 && J1 && x2 && J2 && out, -1);
  return gpxb::hess_gram(reinterpret_cast<Ctx*>(ctx), kind, x1, J1, n1, nv1, x2, J2, n2, nv2, nf, ell, diag_add,
                         out, ldo, lower_only, (cudaStream_t)stream);
}

int gpxb_gpr_lml_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf,
                        int nv, double ell, double sigma, double* out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && out, -1);
  return gpxb::gpr_derivs(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, N, nf, nv, ell, sigma, out, nullptr, nullptr,
                          (cudaStream_t)stream);
}

int gpxb_gpr_lml_grad_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N,
                             int nf, int nv, double ell, double sigma, int want_grad, double* out3,
                             gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && out3, -1);
  return gpxb::gpr_derivs_grad(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, N, nf, nv, ell, sigma, want_grad, out3,
                               (cudaStream_t)stream);
}

int gpxb_gpr_predict_y_derivs(gpxb_ctx* ctx, int kind, const double* x_train, const double* jaccoef, int N, int nf,
                              const double* x, int ns, double ell, double mu, double* mean_out,
                              gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x_train && jaccoef && x && mean_out, -1);
  return gpxb::gpr_predict_y_derivs(reinterpret_cast<Ctx*>(ctx), kind, x_train, jaccoef, N, nf, x, ns, ell, mu,
                                    mean_out, (cudaStream_t)stream);
}

int gpxb_gpr_predict_derivs_full(gpxb_ctx* ctx, int kind, const double* x_train, const double* J_train, int N, int nf,
                                 int nv, const double* x, const double* J, int ns, int nvs, const double* c,
                                 double mu, double ell, double sigma, double* mean_out, double* cov_out,
                                 gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x_train && J_train && x && J && c && mean_out, -1);
  return gpxb::gpr_predict_derivs_full(reinterpret_cast<Ctx*>(ctx), kind, x_train, J_train, N, nf, nv, x, J, ns, nvs,
                                       c, mu, ell, sigma, mean_out, cov_out, (cudaStream_t)stream);
}

int gpxb_gpr_predict_y_derivs_full(gpxb_ctx* ctx, int kind, const double* x_train, const double* J_train, int N,
                                   int nf, int nv, const double* x, int ns, const double* c, double mu, double ell,
                                   double sigma, double* mean_out, double* cov_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x_train && J_train && x && c && mean_out, -1);
  return gpxb::gpr_predict_y_derivs_full(reinterpret_cast<Ctx*>(ctx), kind, x_train, J_train, N, nf, nv, x, ns, c, mu,
                                         ell, sigma, mean_out, cov_out, (cudaStream_t)stream);
}

int gpxb_hess_matvec(gpxb_ctx* ctx, int kind, const double* x1, const double* J1, int n1, int nv1, const double* x2,
                     const double* J2, int n2, int nv2, int nf, double ell, double diag_add, const double* V,
                     long long ldv, int P, double* W, long long ldw, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x1 && J1 && x2 && J2 && V && W && P > 0, -1);
  return gpxb::hess_matvec(reinterpret_cast<Ctx*>(ctx), kind, x1, J1, n1, nv1, x2, J2, n2, nv2, nf, ell, diag_add, 0, V,
                           ldv, P, W, ldw, (cudaStream_t)stream);
}

int gpxb_hess_matvec_dl(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nv, int nf, double ell,
                        const double* V, long long ldv, int P, double* W, long long ldw, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && V && W && P > 0, -1);
  return gpxb::hess_matvec(reinterpret_cast<Ctx*>(ctx), kind, x, J, N, nv, x, J, N, nv, nf, ell, 0.0, 1, V, ldv, P, W,
                           ldw, (cudaStream_t)stream);
}

int gpxb_lanczos_grad_derivs(gpxb_ctx* ctx_, int kind, const double* x, const double* J, int N, int nf, int nv,
                             double ell, double sigma, const double* U, long long ldu, int P, int m,
                             double* alpha_out, double* beta_out, double* dalpha_out, double* dbeta_out,
                             int* breakdown_flag, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && J && U && alpha_out && beta_out && dalpha_out && dbeta_out && breakdown_flag, -1);
  GPXB_ARG_CHECK(P > 0 && m > 0, -2);
  return gpxb::lanczos_grad_derivs(reinterpret_cast<Ctx*>(ctx_), kind, x, J, N, nf, nv, ell, sigma, U, ldu, P, m,
                                   alpha_out, beta_out, dalpha_out, dbeta_out, breakdown_flag, (cudaStream_t)stream);
}

int gpxb_gpr_fit_derivs_iter(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N,
                             int nf, int nv, double ell, double sigma, int n_pivots, uint32_t key0, uint32_t key1,
                             int partitionable, double tol, double atol, int maxiter, double* c_out,
                             double* jaccoef_out, int* iters_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && c_out, -1);
  const uint32_t key[2] = {key0, key1};
  return gpxb::gpr_derivs_iter(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, N, nf, nv, ell, sigma, n_pivots, key,
                               partitionable, tol, atol, maxiter, c_out, jaccoef_out, iters_out, nullptr, 0, 0, 0,
                               nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int gpxb_rpcholesky_derivs(gpxb_ctx* ctx_, int kind, const double* x, const double* J, int N, int nf, int nv,
                           double ell, uint32_t key0, uint32_t key1, int M, int partitionable, double* F_out,
                           long long* pivots_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && J && pivots_out && N > 0 && nf > 0 && nv > 0 && M > 0, -1);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)N * nv;
  GPXB_ARG_CHECK(n < (1LL << 31), -2);
  gpxb::DevBuf fb;
  GPXB_TRY(fb.alloc(sizeof(double) * std::max<size_t>(gpxb::rp_factor_doubles((int)n, M), 8), s));
  const uint32_t key[2] = {key0, key1};
  GPXB_TRY(gpxb::hess_rpcholesky(ctx, kind, x, J, N, nf, nv, ell, key, M, partitionable, fb.as<double>(), pivots_out,
                                 s));
  if (F_out) GPXB_TRY(gpxb::rp_unblock_rowmajor(ctx, fb.as<double>(), (int)n, M, F_out, s));
  return 0;
}

int gpxb_sgpr_fit_derivs_iter(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N,
                              int nf, int nv, const double* x_locs, const double* J_locs, int M, int nv_locs, double ell,
                              double sigma, double tol, double atol, int maxiter, double* c_out, int* iters_out,
                              gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && x_locs && J_locs && c_out && N > 0 && M > 0 && nf > 0 && nv > 0 && nv_locs > 0, -1);
  return gpxb::sgpr_fit_derivs_iter(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, N, nf, nv, x_locs, J_locs, M, nv_locs,
                                    ell, sigma, tol, atol, maxiter, c_out, iters_out, (cudaStream_t)stream);
}

int gpxb_gpr_td_fit_iter(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y,
                         const double* y_derivs, int N, int nf, int nv, int nv_first, double ell, double sigma_targets,
                         double sigma_derivs, double sigma_derivs2, int mean_mode, int n_pivots, uint32_t key0,
                         uint32_t key1, int partitionable, double tol, double atol, int maxiter, double* c_out,
                         double* mu_out, int* iters_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && y_derivs && c_out && mu_out && N > 0 && nf > 0 && nv > 0, -1);
  const uint32_t key[2] = {key0, key1};
  return gpxb::gpr_td_iter(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, y_derivs, N, nf, nv, nv_first, ell,
                           sigma_targets, sigma_derivs, sigma_derivs2, mean_mode, n_pivots, key, partitionable, tol, atol,
                           maxiter, c_out, mu_out, iters_out, nullptr, 0, 0, 0, nullptr, nullptr, nullptr,
                           (cudaStream_t)stream);
}

int gpxb_lanczos_td(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, int nv_first,
                    double ell, double sigma_targets, double sigma_derivs, double sigma_derivs2, const double* U,
                    long long ldu, int P, int m, double* alpha_out, double* beta_out, int* breakdown_flag,
                    gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && U && alpha_out && beta_out && breakdown_flag && P > 0 && m > 0, -1);
  const uint32_t key[2] = {0, 0};
  return gpxb::gpr_td_iter(reinterpret_cast<Ctx*>(ctx), kind, x, J, nullptr, nullptr, N, nf, nv, nv_first, ell,
                           sigma_targets, sigma_derivs, sigma_derivs2, 0, 0, key, 1, 0.0, 0.0, 0, nullptr, nullptr,
                           nullptr, U, ldu, P, m, alpha_out, beta_out, breakdown_flag, (cudaStream_t)stream);
}

int gpxb_rpcholesky_td(gpxb_ctx* ctx_, int kind, const double* x, const double* J, int N, int nf, int nv, int nv_first,
                       double ell, uint32_t key0, uint32_t key1, int M, int partitionable, double* F_out,
                       long long* pivots_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && J && pivots_out && N > 0 && nf > 0 && nv > 0 && M > 0, -1);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)N + (long long)N * nv;
  GPXB_ARG_CHECK(n < (1LL << 31), -2);
  gpxb::DevBuf fb;
  GPXB_TRY(fb.alloc(sizeof(double) * std::max<size_t>(gpxb::rp_factor_doubles((int)n, M), 8), s));
  const uint32_t key[2] = {key0, key1};
  GPXB_TRY(gpxb::td_rpcholesky(ctx, kind, x, J, N, nf, nv, nv_first, ell, key, M, partitionable, fb.as<double>(),
                               pivots_out, s));
  if (F_out) GPXB_TRY(gpxb::rp_unblock_rowmajor(ctx, fb.as<double>(), (int)n, M, F_out, s));
  return 0;
}

int gpxb_lanczos_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, double ell,
                        double sigma, const double* U, long long ldu, int P, int m, double* alpha_out,
                        double* beta_out, int* breakdown_flag, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && U && alpha_out && beta_out && breakdown_flag && P > 0 && m > 0, -1);
  const uint32_t key[2] = {0, 0};
  return gpxb::gpr_derivs_iter(reinterpret_cast<Ctx*>(ctx), kind, x, J, nullptr, N, nf, nv, ell, sigma, 0, key, 1, 0.0,
                               0.0, 0, nullptr, nullptr, nullptr, U, ldu, P, m, alpha_out, beta_out, breakdown_flag,
                               (cudaStream_t)stream);
}

int gpxb_gpr_fit_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf,
                        int nv, double ell, double sigma, double* c_out, double* jaccoef_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && c_out, -1);
  return gpxb::gpr_derivs(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, N, nf, nv, ell, sigma, nullptr, c_out,
                          jaccoef_out, (cudaStream_t)stream);
}

}  // extern "C"
