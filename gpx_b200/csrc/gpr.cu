// Exact GPR on device: fused LML (+ analytic gradient), fit and predict.
//
// Replaces gpx/models/_gpr.py:_lml_dense (737-769), _fit_dense (262-282),
// _predict_dense (434-463) and the gradient jit(value_and_grad(loss)) yields
// (gpx/optimizers/scipy_optimize.py:72-78) in analytic form:
//   alpha = A^-1 r,  W = alpha alpha^T - A^-1,
//   d lml/d ell = (1/2 sum_ij W_ij dK_ij/d ell)/N,  d lml/d sigma = sigma tr(W)/N.
#include "../../include/gpxb200.h"
#include "chol.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "mma.cuh"
#include "small_kernels.cuh"

namespace gpxb {

namespace {

using namespace sk;

// scal: [0] quad, [1] sum log L_ii, [2] sum W dK, [3] sum alpha^2, [4] tr(Ainv)
__global__ void finalize_lml_kernel(const double* __restrict__ scal, int N, double sigma,
                                    int want_grad, double* __restrict__ out3) {
  const double n = (double)N;
  out3[0] = (-0.5 * scal[0] - scal[1] - n * 0.5 * log(2.0 * 3.141592653589793)) / n;
  if (want_grad) {
    out3[1] = 0.5 * scal[2] / n;
    out3[2] = sigma * (scal[3] - scal[4]) / n;
  } else {
    out3[1] = 0.0;
    out3[2] = 0.0;
  }
}

}  // namespace

// Builds A = K + s2 I (lower) and factors it.  Z, A, inv are caller scratch.
static int build_and_factor(Ctx* ctx, int kind, const double* x, int N, int D, double ell, double s2,
                            double* Z, double* A, double* inv, cudaStream_t s) {
  const ZLayout zl = z_layout(D);
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, Z, s));
  GramArgs g;
  g.kind = kind;
  g.Z1 = Z; g.n1 = N; g.Z2 = Z; g.n2 = N; g.zl = zl;
  g.ell = ell; g.diag_add = s2; g.sym = 1; g.lower_only = 1;
  g.out = A; g.ldo = N;
  GPXB_TRY(gram_launch(ctx, g, s));
  return potrf_lower(ctx, A, N, N, inv, s);
}

int gpr_lml_grad(Ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T,
                 double ell, double sigma, int mean_mode, int want_grad, double* out3,
                 cudaStream_t s) {
  GPXB_ARG_CHECK(N > 0 && D > 0 && T > 0, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3, -2);
  const ZLayout zl = z_layout(D);
  const double s2 = sigma * sigma + 1e-10;
  const size_t nn = (size_t)N * (size_t)N;

  DevBuf bZ, bA, bInv, bR, bScal, bMu, bW, bT, bPart;
  GPXB_TRY(bZ.alloc(sizeof(double) * (size_t)N * zl.S, s));
  GPXB_TRY(bA.alloc(sizeof(double) * nn, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles(N), s));
  GPXB_TRY(bR.alloc(sizeof(double) * (size_t)N * T, s));
  GPXB_TRY(bScal.alloc(sizeof(double) * 8, s));
  GPXB_TRY(bMu.alloc(sizeof(double), s));
  double *Z = bZ.as<double>(), *A = bA.as<double>(), *inv = bInv.as<double>(), *rT = bR.as<double>();
  double *scal = bScal.as<double>(), *mu = bMu.as<double>();
  GPXB_CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double) * 8, s));

  mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N * T, mean_mode == GPXB_MEAN_DATA, mu);
  GPXB_LAUNCHED(ctx);
  center_transpose_kernel<<<ceil_div((long long)N * T, 256), 256, 0, s>>>(y, N, T, mu, rT);
  GPXB_LAUNCHED(ctx);

  GPXB_TRY(build_and_factor(ctx, kind, x, N, D, ell, s2, Z, A, inv, s));
  GPXB_TRY(logdet_lower(ctx, A, N, N, scal + 1, 1.0, s));
  GPXB_TRY(trsm_right_lt(ctx, A, N, inv, rT, N, T, N, s));  // rT <- (L^-1 r)^T
  sumsq_kernel<<<1, 1024, 0, s>>>(rT, (long long)N * T, scal + 0, 0);
  GPXB_LAUNCHED(ctx);

  if (want_grad) {
    GPXB_TRY(trsm_right_l(ctx, A, N, inv, rT, N, T, N, s));  // rT <- alpha^T
    sumsq_kernel<<<1, 1024, 0, s>>>(rT, (long long)N * T, scal + 3, 0);
    GPXB_LAUNCHED(ctx);
    GPXB_TRY(bW.alloc(sizeof(double) * nn, s));
    GPXB_TRY(bT.alloc(sizeof(double) * potri_scratch_doubles(N), s));
    GPXB_TRY(potri_lower(ctx, A, N, inv, N, bW.as<double>(), bT.as<double>(), A, N, s));
    diag_sum_kernel<<<1, 1024, 0, s>>>(A, N, N, scal + 4);
    GPXB_LAUNCHED(ctx);
    GramArgs g;
    g.kind = kind;
    g.Z1 = Z; g.n1 = N; g.Z2 = Z; g.n2 = N; g.zl = zl;
    g.ell = ell; g.sym = 1; g.lower_only = 1;
    g.Ainv = A; g.ldainv = N; g.alpha = rT; g.t = T;
    const long long nparts = gram_num_ctas(g);
    GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)nparts, s));
    g.partials = bPart.as<double>();
    GPXB_TRY(gram_launch(ctx, g, s));
    sum_kernel<<<1, 1024, 0, s>>>(g.partials, nparts, scal + 2);
    GPXB_LAUNCHED(ctx);
  }
  finalize_lml_kernel<<<1, 1, 0, s>>>(scal, N, sigma, want_grad, out3);
  GPXB_LAUNCHED(ctx);
  return 0;
}

int gpr_fit(Ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T, double ell,
            double sigma, int mean_mode, double* c_out, double* mu_out, cudaStream_t s) {
  GPXB_ARG_CHECK(N > 0 && D > 0 && T > 0, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3, -2);
  const ZLayout zl = z_layout(D);
  const double s2 = sigma * sigma + 1e-10;
  DevBuf bZ, bA, bInv, bR;
  GPXB_TRY(bZ.alloc(sizeof(double) * (size_t)N * zl.S, s));
  GPXB_TRY(bA.alloc(sizeof(double) * (size_t)N * N, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles(N), s));
  GPXB_TRY(bR.alloc(sizeof(double) * (size_t)N * T, s));
  double* rT = bR.as<double>();
  mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N * T, mean_mode == GPXB_MEAN_DATA, mu_out);
  GPXB_LAUNCHED(ctx);
  center_transpose_kernel<<<ceil_div((long long)N * T, 256), 256, 0, s>>>(y, N, T, mu_out, rT);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(build_and_factor(ctx, kind, x, N, D, ell, s2, bZ.as<double>(), bA.as<double>(),
                            bInv.as<double>(), s));
  GPXB_TRY(trsm_right_lt(ctx, bA.as<double>(), N, bInv.as<double>(), rT, N, T, N, s));
  GPXB_TRY(trsm_right_l(ctx, bA.as<double>(), N, bInv.as<double>(), rT, N, T, N, s));
  transpose_out_kernel<<<ceil_div((long long)N * T, 256), 256, 0, s>>>(rT, N, T, c_out);
  GPXB_LAUNCHED(ctx);
  return 0;
}

int gpr_predict(Ctx* ctx, int kind, const double* x_train, int N, int D, const double* x, int ns,
                const double* c, int T, const double* mu, double ell, double sigma, double* mean_out,
                double* cov_out, cudaStream_t s) {
  GPXB_ARG_CHECK(N > 0 && D > 0 && T > 0 && ns > 0, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3, -2);
  const ZLayout zl = z_layout(D);
  DevBuf bZt, bZs, bKs;
  GPXB_TRY(bZt.alloc(sizeof(double) * (size_t)N * zl.S, s));
  GPXB_TRY(bZs.alloc(sizeof(double) * (size_t)ns * zl.S, s));
  GPXB_TRY(bKs.alloc(sizeof(double) * (size_t)ns * N, s));
  double *Zt = bZt.as<double>(), *Zs = bZs.as<double>(), *Ks = bKs.as<double>();
  GPXB_TRY(prep_z(ctx, x_train, D, N, zl, ell, Zt, s));
  GPXB_TRY(prep_z(ctx, x, D, ns, zl, ell, Zs, s));
  {
    GramArgs g;  // Ks = K(x, x_train)  [ns x N]
    g.kind = kind;
    g.Z1 = Zs; g.n1 = ns; g.Z2 = Zt; g.n2 = N; g.zl = zl; g.ell = ell;
    g.out = Ks; g.ldo = N;
    GPXB_TRY(gram_launch(ctx, g, s));
  }
  {
    GemmArgs g;  // mean = Ks c
    g.M = ns; g.N = T; g.K = N;
    g.A = Ks; g.lda = N; g.a_kmajor = true;
    g.B = c; g.ldb = T; g.b_kmajor = false;
    g.C = mean_out; g.ldc = T;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  add_scalar_kernel<<<ceil_div((long long)ns * T, 256), 256, 0, s>>>(mean_out, (long long)ns * T, mu);
  GPXB_LAUNCHED(ctx);
  if (cov_out) {
    const double s2 = sigma * sigma + 1e-10;
    DevBuf bA, bInv;
    GPXB_TRY(bA.alloc(sizeof(double) * (size_t)N * N, s));
    GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles(N), s));
    double *A = bA.as<double>(), *inv = bInv.as<double>();
    GramArgs g;
    g.kind = kind;
    g.Z1 = Zt; g.n1 = N; g.Z2 = Zt; g.n2 = N; g.zl = zl; g.ell = ell;
    g.diag_add = s2; g.sym = 1; g.lower_only = 1;
    g.out = A; g.ldo = N;
    GPXB_TRY(gram_launch(ctx, g, s));
    GPXB_TRY(potrf_lower(ctx, A, N, N, inv, s));
    GPXB_TRY(trsm_right_lt(ctx, A, N, inv, Ks, N, ns, N, s));  // Ks <- G^T = Ks L^-T
    GramArgs h;
    h.kind = kind;
    h.Z1 = Zs; h.n1 = ns; h.Z2 = Zs; h.n2 = ns; h.zl = zl; h.ell = ell;
    h.sym = 1;
    h.out = cov_out; h.ldo = ns;
    GPXB_TRY(gram_launch(ctx, h, s));
    GemmArgs m;  // cov -= G^T G
    m.M = ns; m.N = ns; m.K = N;
    m.A = Ks; m.lda = N; m.a_kmajor = true;
    m.B = Ks; m.ldb = N; m.b_kmajor = true;
    m.C = cov_out; m.ldc = ns;
    m.alpha = -1.0; m.beta = 1.0;
    GPXB_TRY(gemm_f64(ctx, m, s));
  }
  return 0;
}

}  // namespace gpxb

using gpxb::Ctx;

extern "C" {

int gpxb_gpr_lml_grad(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T,
                      double ell, double sigma, int mean_mode, int want_grad, double* out3,
                      gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && y && out3, -1);
  return gpxb::gpr_lml_grad(reinterpret_cast<Ctx*>(ctx), kind, x, y, N, D, T, ell, sigma, mean_mode,
                            want_grad, out3, (cudaStream_t)stream);
}

int gpxb_gpr_fit(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T,
                 double ell, double sigma, int mean_mode, double* c_out, double* mu_out,
                 gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && y && c_out && mu_out, -1);
  return gpxb::gpr_fit(reinterpret_cast<Ctx*>(ctx), kind, x, y, N, D, T, ell, sigma, mean_mode,
                       c_out, mu_out, (cudaStream_t)stream);
}

int gpxb_gpr_predict(gpxb_ctx* ctx, int kind, const double* x_train, int N, int D, const double* x,
                     int ns, const double* c, int T, const double* mu, double ell, double sigma,
                     double* mean_out, double* cov_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x_train && x && c && mu && mean_out, -1);
  return gpxb::gpr_predict(reinterpret_cast<Ctx*>(ctx), kind, x_train, N, D, x, ns, c, T, mu, ell,
                           sigma, mean_out, cov_out, (cudaStream_t)stream);
}

}  // extern "C"
