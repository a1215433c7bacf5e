// Iterative (matrix-free) SGPR paths on device.
//
// Replaces gpx/models/_sgpr.py:_Ax_lhs_fun (96-143), _fit_iter (343-363), _Hx_fun (204-239), _lml_iter (701-737)
// and _predict_iter (530-546): CG on the M x M normal equations
//     (sigma^2 K_mm + K_mn K_nm + 1e-10 I) c = K_mn (y - mu)
// and, for the LML, CG + stochastic Lanczos quadrature on the N x N Nystrom operator
//     H + (sigma^2 + 1e-10) I,   H = K_nm (K_mm + 1e-10 I)^-1 K_mn.
// No K_mn block is stored: every product is a matrix-free block mat-vec of matvec.cu between the data rows
// and the landmarks.  Single GPU (the dense Woodbury path is the sharded one).
#include <functional>

#include "../../include/gpxb200.h"
#include "chol.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "matvec.cuh"
#include "small_kernels.cuh"

namespace gpxb {

int pcg_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
             const double* b, const double* F, long long ldf, int k, const double* Pkk, double sigma, double tol,
             double atol, int maxiter, double* x, int* iters_out, cudaStream_t s);
int lanczos_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                 const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                 cudaStream_t s);

int lanczos_grad_core(Ctx* ctx, int n_loc, long long N,
                      const std::function<int(const double*, double*, bool)>& matvec, double sigma, const double* U,
                      int P, int PS, int m, double* alpha_out, double* beta_out, double* dalpha_out,
                      double* dbeta_out, int Ptot, int* flag, cudaStream_t s);

namespace {
using namespace sk;

// out[i*stride] = y[i] - *mu
__global__ void center_stride_kernel(const double* __restrict__ y, long long n, const double* __restrict__ mu,
                                     double* __restrict__ out, int stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i * stride] = y[i] - *mu;
}
// out[i*so] = a * x[i*sx] + b * y[i*sy] + c * z[i*sz]   (y / z may be null; out may alias x)
__global__ void lin3_kernel(double* out, int so, double a, const double* x, int sx, double b, const double* y, int sy,
                            double c, const double* z, int sz, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = a * x[i * sx];
  if (y) v += b * y[i * sy];
  if (z) v += c * z[i * sz];
  out[i * so] = v;
}
// W[:, 0..P) += shift * V[:, 0..P) on n x PS blocks
__global__ void block_shift_kernel(double* __restrict__ W, const double* __restrict__ V, double shift, long long n,
                                   int PS, int P) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * PS) return;
  if ((int)(idx % PS) < P) W[idx] += shift * V[idx];
}

// W[i][0..P) += X[i][0..P)
__global__ void add_cols_kernel(double* __restrict__ W, long long ldw, const double* __restrict__ X, long long ldx,
                                long long n, int P) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * P) return;
  const long long i = idx / P;
  const int p = (int)(idx % P);
  W[i * ldw + p] += X[i * ldx + p];
}
// *out = sum_i x[i sx] y[i sy]  (single CTA, fixed order)
__global__ void strided_dot_kernel(const double* __restrict__ x, int sx, const double* __restrict__ y, int sy, long long n,
                                   double* __restrict__ out) {
  __shared__ double sh[32];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) a += x[i * sx] * y[i * sy];
  a = block_sum<1024>(a, sh);
  if (threadIdx.x == 0) *out = a;
}

struct SgprOps {
  Ctx* ctx;
  int kind;
  const double *Z, *Zl;  // pre-scaled data rows (N x S) and landmarks (M x S)
  int N, M;
  ZLayout zl;
  cudaStream_t s;
  double ell = 1.0;  // only read by the d/d lengthscale products
  // out[rows x ldw] = K(rows, cols) V, V: cols x PS padded block with P live columns; dl: d K / d lengthscale instead
  int mv(bool rows_are_data, bool cols_are_data, const double* V, int PS, int P, double* out, long long ldw,
         bool dl = false) const {
    KmvArgs a;
    if (dl) a.dl_ell = ell;
    a.kind = kind;
    a.Zr = rows_are_data ? Z : Zl; a.n_rows = rows_are_data ? N : M;
    a.Za = cols_are_data ? Z : Zl; a.N = cols_are_data ? N : M;
    a.row_offset = (rows_are_data == cols_are_data) ? 0 : -1;  // exact diagonal for K_mm
    a.zl = zl; a.V = V; a.PS = PS; a.P = P; a.W = out; a.ldw = ldw;
    return kmatvec_launch(ctx, a, s);
  }
};
}  // namespace

// c[M] = cg(sigma^2 K_mm + K_mn K_nm + 1e-10 I, K_mn (y - mu)); tol / atol / maxiter as jax.scipy.sparse.linalg.cg
// (maxiter <= 0: 10 M).  mu: device scalar.
int sgpr_fit_iter(Ctx* ctx, int kind, const double* x, const double* y, int N, int D, const double* x_locs, int M,
                  double ell, double sigma, const double* mu, double tol, double atol, int maxiter, double* c_out,
                  int* iters_out, cudaStream_t s) {
  GPXB_ARG_CHECK(ctx->world == 1, -61);
  GPXB_ARG_CHECK(N > 0 && D > 0 && M > 0, -1);
  const ZLayout zl = z_layout(D);
  DevBuf bZ, bZl, bR, bB, bT, bU, bW;
  GPXB_TRY(bZ.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(bZl.alloc(sizeof(double) * ((size_t)M * zl.S + 16), s));
  GPXB_TRY(bR.alloc(sizeof(double) * ((size_t)N * 2 + 16), s));
  GPXB_TRY(bB.alloc(sizeof(double) * (size_t)M, s));
  GPXB_TRY(bT.alloc(sizeof(double) * ((size_t)N * 2 + 16), s));
  GPXB_TRY(bU.alloc(sizeof(double) * (size_t)M, s));
  GPXB_TRY(bW.alloc(sizeof(double) * (size_t)M, s));
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, bZ.as<double>(), s));
  GPXB_TRY(prep_z(ctx, x_locs, D, M, zl, ell, bZl.as<double>(), s));
  const SgprOps op{ctx, kind, bZ.as<double>(), bZl.as<double>(), N, M, zl, s};
  double *r2 = bR.as<double>(), *b = bB.as<double>(), *t2 = bT.as<double>(), *u = bU.as<double>(), *w = bW.as<double>();
  GPXB_CUDA_CHECK(cudaMemsetAsync(r2, 0, sizeof(double) * ((size_t)N * 2 + 16), s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(t2, 0, sizeof(double) * ((size_t)N * 2 + 16), s));
  center_stride_kernel<<<ceil_div(N, 256), 256, 0, s>>>(y, N, mu, r2, 2);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(op.mv(false, true, r2, 2, 1, b, 1));  // rhs = K_mn (y - mu)
  const double s2 = sigma * sigma;
  auto apply = [&](const double* p, double* Ap) -> int {
    GPXB_TRY(op.mv(true, false, p, 2, 1, t2, 2));    // t = K_nm p
    GPXB_TRY(op.mv(false, true, t2, 2, 1, u, 1));    // u = K_mn t
    GPXB_TRY(op.mv(false, false, p, 2, 1, w, 1));    // w = K_mm p
    lin3_kernel<<<ceil_div(M, 256), 256, 0, s>>>(Ap, 1, s2, w, 1, 1.0, u, 1, 1e-10, p, 2, M);
    GPXB_LAUNCHED(ctx);
    return 0;
  };
  return pcg_core(ctx, M, M, apply, b, nullptr, 0, 0, nullptr, sigma, tol, atol,
                  maxiter > 0 ? maxiter : 10 * M, c_out, iters_out, s);
}

// The N x N Nystrom operator of _Hx_fun: quad_out (device) = r^T cg(H + shift, r) with r = y - mu, and/or the
// m-step block Lanczos tridiagonals on it for the unit start vectors U (N x ldu, P columns).
// Gradient mode (gquad_out / dalpha_out / dbeta_out non-null; what jit(value_and_grad) of _lml_iter :700-737 needs):
//   gquad_out[3] (device) = { (dK_mn c) . g,  g . (dK_mm g),  c . c },  c = cg(A, r), g = K_mm^-1 K_mn c, so that
//   c^T (dA/d ell) c = 2 gquad[0] - gquad[1] and c^T (dA/d sigma) c = 2 sigma gquad[2]   (d(r.c) = -c^T dA c);
//   dalpha_out / dbeta_out [2][P][m]: tangents of the Lanczos tridiagonals w.r.t. (lengthscale, sigma), with
//   (dA/d ell) v = dK_nm g_v + K_nm K_mm^-1 (dK_mn v - dK_mm g_v),  g_v = K_mm^-1 K_mn v.
int sgpr_hx_iter(Ctx* ctx, int kind, const double* x, const double* y, int N, int D, const double* x_locs, int M,
                 double ell, double sigma, const double* mu, double tol, double atol, int maxiter, double* quad_out,
                 int* iters_out, const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                 int* flag, double* gquad_out, double* dalpha_out, double* dbeta_out, cudaStream_t s) {
  GPXB_ARG_CHECK(ctx->world == 1, -61);
  GPXB_ARG_CHECK(N > 0 && D > 0 && M > 0, -1);
  const ZLayout zl = z_layout(D);
  const double shift = sigma * sigma + 1e-10;
  const size_t mm = (size_t)M * M;
  DevBuf bZ, bZl, bA, bInv, bWk, bTs, bT1, bT2;
  GPXB_TRY(bZ.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(bZl.alloc(sizeof(double) * ((size_t)M * zl.S + 16), s));
  GPXB_TRY(bA.alloc(sizeof(double) * mm, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  GPXB_TRY(bWk.alloc(sizeof(double) * mm, s));
  GPXB_TRY(bTs.alloc(sizeof(double) * potri_scratch_doubles(M), s));
  GPXB_TRY(bT1.alloc(sizeof(double) * ((size_t)M * 34 + 16), s));
  GPXB_TRY(bT2.alloc(sizeof(double) * ((size_t)M * 34 + 16), s));
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, bZ.as<double>(), s));
  GPXB_TRY(prep_z(ctx, x_locs, D, M, zl, ell, bZl.as<double>(), s));
  double* A = bA.as<double>();
  {
    // K_mm^-1 as an explicit inverse (the reference calls jnp.linalg.inv), here through the Cholesky factor
    GramArgs g;
    g.kind = kind; g.Z1 = bZl.as<double>(); g.n1 = M; g.Z2 = bZl.as<double>(); g.n2 = M; g.zl = zl; g.ell = ell;
    g.diag_add = 1e-10; g.sym = 1; g.lower_only = 1; g.out = A; g.ldo = M;
    GPXB_TRY(gram_launch(ctx, g, s));
    GPXB_TRY(potrf_lower(ctx, A, M, M, bInv.as<double>(), s));
    GPXB_TRY(potri_lower(ctx, A, M, bInv.as<double>(), M, bWk.as<double>(), bTs.as<double>(), A, M, s));
    symmetrize_lower_kernel<<<ceil_div((long long)mm, 256), 256, 0, s>>>(A, M);
    GPXB_LAUNCHED(ctx);
  }
  bWk.release();
  const SgprOps op{ctx, kind, bZ.as<double>(), bZl.as<double>(), N, M, zl, s, ell};
  double *t1 = bT1.as<double>(), *t2 = bT2.as<double>();
  GPXB_CUDA_CHECK(cudaMemsetAsync(t1, 0, sizeof(double) * ((size_t)M * 34 + 16), s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(t2, 0, sizeof(double) * ((size_t)M * 34 + 16), s));
  // w = K_nm Kmm^-1 K_mn v + shift v on an N x PS block
  auto hx = [&](const double* v, int PS, int Pn, double* w, long long ldw) -> int {
    GPXB_TRY(op.mv(false, true, v, PS, Pn, t1, PS));  // t1 = K_mn v          (M x PS)
    GemmArgs g;                                        // t2 = Kmm^-1 t1
    g.M = M; g.N = Pn; g.K = M;
    g.A = A; g.lda = M; g.a_kmajor = true;
    g.B = t1; g.ldb = PS; g.b_kmajor = false;
    g.C = t2; g.ldc = PS;
    GPXB_TRY(gemm_f64(ctx, g, s));
    return op.mv(true, false, t2, PS, Pn, w, ldw);     // w = K_nm t2
  };
  // w = (dA / d lengthscale) v on an N x PS block (no shift)
  DevBuf bT3, bT4, bWn;
  const bool want_grad = gquad_out || dalpha_out;
  if (want_grad) {
    GPXB_TRY(bT3.alloc(sizeof(double) * ((size_t)M * 34 + 16), s));
    GPXB_TRY(bT4.alloc(sizeof(double) * ((size_t)M * 34 + 16), s));
    GPXB_TRY(bWn.alloc(sizeof(double) * ((size_t)N * 34 + 16), s));
    GPXB_CUDA_CHECK(cudaMemsetAsync(bT3.as<double>(), 0, sizeof(double) * ((size_t)M * 34 + 16), s));
    GPXB_CUDA_CHECK(cudaMemsetAsync(bT4.as<double>(), 0, sizeof(double) * ((size_t)M * 34 + 16), s));
  }
  double *t3 = bT3.as<double>(), *t4 = bT4.as<double>(), *wn = bWn.as<double>();
  auto kmm_inv = [&](const double* in, double* out, int PS, int Pn) -> int {
    GemmArgs g;
    g.M = M; g.N = Pn; g.K = M;
    g.A = A; g.lda = M; g.a_kmajor = true;
    g.B = in; g.ldb = PS; g.b_kmajor = false;
    g.C = out; g.ldc = PS;
    return gemm_f64(ctx, g, s);
  };
  auto hx_dl = [&](const double* v, int PS, int Pn, double* w, long long ldw) -> int {
    GPXB_TRY(op.mv(false, true, v, PS, Pn, t1, PS));            // t1 = K_mn v
    GPXB_TRY(kmm_inv(t1, t2, PS, Pn));                          // t2 = g_v
    GPXB_TRY(op.mv(true, false, t2, PS, Pn, w, ldw, true));     // w  = dK_nm g_v
    GPXB_TRY(op.mv(false, true, v, PS, Pn, t3, PS, true));      // t3 = dK_mn v
    GPXB_TRY(op.mv(false, false, t2, PS, Pn, t4, PS, true));    // t4 = dK_mm g_v
    lin3_kernel<<<ceil_div((long long)M * PS, 256), 256, 0, s>>>(t3, 1, 1.0, t3, 1, -1.0, t4, 1, 0.0, nullptr, 0,
                                                                   (long long)M * PS);
    GPXB_LAUNCHED(ctx);
    GPXB_TRY(kmm_inv(t3, t4, PS, Pn));                          // t4 = K_mm^-1 (dK_mn v - dK_mm g_v)
    GPXB_TRY(op.mv(true, false, t4, PS, Pn, wn, PS));           // wn = K_nm t4
    add_cols_kernel<<<ceil_div((long long)N * Pn, 256), 256, 0, s>>>(w, ldw, wn, PS, N, Pn);
    GPXB_LAUNCHED(ctx);
    return 0;
  };
  if (quad_out) {
    DevBuf bR, bC, bPart;
    GPXB_TRY(bR.alloc(sizeof(double) * (size_t)N, s));
    GPXB_TRY(bC.alloc(sizeof(double) * (size_t)N, s));
    center_stride_kernel<<<ceil_div(N, 256), 256, 0, s>>>(y, N, mu, bR.as<double>(), 1);
    GPXB_LAUNCHED(ctx);
    auto apply = [&](const double* p, double* Ap) -> int {
      GPXB_TRY(hx(p, 2, 1, Ap, 1));
      lin3_kernel<<<ceil_div(N, 256), 256, 0, s>>>(Ap, 1, 1.0, Ap, 1, shift, p, 2, 0.0, nullptr, 0, N);
      GPXB_LAUNCHED(ctx);
      return 0;
    };
    GPXB_TRY(pcg_core(ctx, N, N, apply, bR.as<double>(), nullptr, 0, 0, nullptr, sigma, tol, atol,
                      maxiter > 0 ? maxiter : (int)std::min<long long>(10LL * N, 2000000000LL), bC.as<double>(),
                      iters_out, s));
    dot_kernel<<<1, 1024, 0, s>>>(bR.as<double>(), bC.as<double>(), N, quad_out, 0);
    GPXB_LAUNCHED(ctx);
    if (gquad_out) {
      DevBuf bC2;
      GPXB_TRY(bC2.alloc(sizeof(double) * ((size_t)N * 2 + 16), s));
      GPXB_CUDA_CHECK(cudaMemsetAsync(bC2.as<double>(), 0, sizeof(double) * ((size_t)N * 2 + 16), s));
      lin3_kernel<<<ceil_div(N, 256), 256, 0, s>>>(bC2.as<double>(), 2, 1.0, bC.as<double>(), 1, 0.0, nullptr, 0, 0.0,
                                                    nullptr, 0, N);
      GPXB_LAUNCHED(ctx);
      GPXB_TRY(op.mv(false, true, bC2.as<double>(), 2, 1, t1, 2));        // t1 = K_mn c
      GPXB_TRY(kmm_inv(t1, t2, 2, 1));                                    // t2 = g
      GPXB_TRY(op.mv(false, true, bC2.as<double>(), 2, 1, t3, 2, true));  // t3 = dK_mn c
      GPXB_TRY(op.mv(false, false, t2, 2, 1, t4, 2, true));               // t4 = dK_mm g
      strided_dot_kernel<<<1, 1024, 0, s>>>(t3, 2, t2, 2, M, gquad_out + 0);
      strided_dot_kernel<<<1, 1024, 0, s>>>(t2, 2, t4, 2, M, gquad_out + 1);
      strided_dot_kernel<<<1, 1024, 0, s>>>(bC.as<double>(), 1, bC.as<double>(), 1, N, gquad_out + 2);
      GPXB_LAUNCH_CHECK();
      ctx->launches += 3;
    }
  }
  if (U && alpha_out) {
    GPXB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
    DevBuf up;
    for (int c0 = 0; c0 < P; c0 += 32) {
      const int pb = std::min(32, P - c0);
      const int PS = v_padded_stride(pb);
      GPXB_TRY(up.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
      GPXB_TRY(pack_v(ctx, U, ldu, N, c0, pb, PS, up.as<double>(), s));
      auto apply = [&](const double* v, double* w) -> int {
        GPXB_TRY(hx(v, PS, pb, w, PS));
        block_shift_kernel<<<ceil_div((long long)N * PS, 256), 256, 0, s>>>(w, v, shift, N, PS, pb);
        GPXB_LAUNCHED(ctx);
        return 0;
      };
      if (dalpha_out) {
        auto matvec = [&](const double* v, double* w, bool dl) -> int {
          return dl ? hx_dl(v, PS, pb, w, PS) : apply(v, w);
        };
        GPXB_TRY(lanczos_grad_core(ctx, N, N, matvec, sigma, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                                   beta_out + (size_t)c0 * m, dalpha_out + (size_t)c0 * m, dbeta_out + (size_t)c0 * m,
                                   P, flag, s));
      } else {
        GPXB_TRY(lanczos_core(ctx, N, N, apply, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                              beta_out + (size_t)c0 * m, flag, s));
      }
      up.release();
    }
  }
  return 0;
}

}  // namespace gpxb

using gpxb::Ctx;

extern "C" {

int gpxb_sgpr_fit_iter(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D, const double* x_locs,
                       int M, double ell, double sigma, int mean_mode, double tol, double atol, int maxiter,
                       double* c_out, double* mu_out, int* iters_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && x_locs && c_out && mu_out, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  gpxb::sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu_out);
  GPXB_LAUNCHED(ctx);
  return gpxb::sgpr_fit_iter(ctx, kind, x, y, N, D, x_locs, M, ell, sigma, mu_out, tol, atol, maxiter, c_out,
                             iters_out, s);
}

int gpxb_sgpr_lml_iter_parts(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D,
                             const double* x_locs, int M, double ell, double sigma, int mean_mode, double tol,
                             double atol, int maxiter, double* quad_out, int* iters_out, const double* U,
                             long long ldu, int P, int m, double* alpha_out, double* beta_out, int* breakdown_flag,
                             gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && x_locs && quad_out && U && alpha_out && beta_out && breakdown_flag, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && P > 0 && m > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  gpxb::DevBuf mu;
  GPXB_TRY(mu.alloc(sizeof(double), s));
  gpxb::sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu.as<double>());
  GPXB_LAUNCHED(ctx);
  return gpxb::sgpr_hx_iter(ctx, kind, x, y, N, D, x_locs, M, ell, sigma, mu.as<double>(), tol, atol, maxiter,
                            quad_out, iters_out, U, ldu, P, m, alpha_out, beta_out, breakdown_flag, nullptr, nullptr,
                            nullptr, s);
}

int gpxb_sgpr_lml_iter_grad_parts(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D,
                                  const double* x_locs, int M, double ell, double sigma, int mean_mode, double tol,
                                  double atol, int maxiter, double* quad_out, double* gquad_out, int* iters_out,
                                  const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                                  double* dalpha_out, double* dbeta_out, int* breakdown_flag, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && x_locs && quad_out && gquad_out && U && alpha_out && beta_out && dalpha_out &&
                     dbeta_out && breakdown_flag, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && P > 0 && m > 0 && gpxb::kmatvec_supports(D), -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  gpxb::DevBuf mu;
  GPXB_TRY(mu.alloc(sizeof(double), s));
  gpxb::sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu.as<double>());
  GPXB_LAUNCHED(ctx);
  return gpxb::sgpr_hx_iter(ctx, kind, x, y, N, D, x_locs, M, ell, sigma, mu.as<double>(), tol, atol, maxiter,
                            quad_out, iters_out, U, ldu, P, m, alpha_out, beta_out, breakdown_flag, gquad_out,
                            dalpha_out, dbeta_out, s);
}

}  // extern "C"
