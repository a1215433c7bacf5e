// GPR on targets AND derivatives (GPR_TD): the block kernel matrix
//        / K + s_t I      d1kj        \      rows / columns: N targets, then N nv derivative values
//    A = |                            |
//        \   d0kj      d01kj + s_d I  /
// of gpx/models/gpr_td.py:_A_lhs (43-159), with _lml_dense (507-560), _fit_dense (639-683) and the analytic
// gradient of the LML w.r.t. (lengthscale, sigma_targets, sigma_derivs) that jit(value_and_grad(loss)) yields
// inside scipy_minimize_ol.  The four blocks are written straight into A by the fused Gram kernel (gram.cu),
// the d0kj block kernel and the Hessian kernel (hess.cu); the factorisation, the inverse and the
// gradient contractions are the exact-GPR machinery (chol.cu, gemm.cu) applied block by block:
//   d lml/d ell = 1/2 [ sum W_TT o dK + 2 sum W_DT o d(d0kj) + sum W_DD o d(d01kj) ] / m,  W = alpha alpha^T - A^-1.
// The reference's optional SECOND derivative block (jacobian_2 / y_derivs_2 / sigma_derivs2, gpr_td.py:114-158) is the
// same matrix for the jacobian concatenated along the variable axis, J = [J_1 | J_2] (nv = nv_1 + nv_2), up to a
// symmetric permutation of the derivative rows -- (sample, variable) interleaved here, block after block there --
// which leaves the LML unchanged; the only new ingredient is the per-block noise: variables v >= nv_a carry
// sigma_derivs2.  The facade permutes the coefficient vector to the reference's order.
#include "../../include/gpxb200.h"
#include "chol.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "hess.cuh"
#include "small_kernels.cuh"

namespace gpxb {

namespace {
using namespace sk;

// r = [y - mu ; yd]
__global__ void td_rhs_kernel(const double* __restrict__ y, int N, const double* __restrict__ yd, long long n,
                              const double* __restrict__ mu, double* __restrict__ r) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) r[i] = y[i] - *mu;
  else if (i < N + n) r[i] = yd[i - N];
}
// DD[i][i] += delta for the derivative rows i = s nv + v of the second block (v >= nv_a)
__global__ void td_diag2_kernel(double* __restrict__ DD, long long ld, long long n, int nv, int nv_a, double delta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (int)(i % nv) >= nv_a) DD[i * ld + i] += delta;
}
// out2[0] = sum alpha_D[i]^2, out2[1] = sum Ainv_DD[i][i], both over the rows of the second block (v >= nv_a)
__global__ void td_block2_sums_kernel(const double* __restrict__ aD, const double* __restrict__ DD, long long ld,
                                      long long n, int nv, int nv_a, double* __restrict__ out2) {
  __shared__ double sh[32];
  double s0 = 0.0, s1 = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024)
    if ((int)(i % nv) >= nv_a) {
      s0 += aD[i] * aD[i];
      s1 += DD[i * ld + i];
    }
  s0 = block_sum<1024>(s0, sh);
  __syncthreads();
  s1 = block_sum<1024>(s1, sh);
  if (threadIdx.x == 0) {
    out2[0] = s0;
    out2[1] = s1;
  }
}
// scal: [0] quad, [1] sum log L_ii, [2] TT, [3] DT, [4] DD contractions, [5] |alpha_T|^2, [6] |alpha_D|^2,
//       [7] tr(Ainv_TT), [8] tr(Ainv_DD), [9] / [10]: the second block's share of [6] / [8]
// out5 = {lml, d/d lengthscale, d/d sigma_targets, d/d sigma_derivs, d/d sigma_derivs2}
__global__ void td_finalize_kernel(const double* __restrict__ scal, long long m, double sig_t, double sig_d,
                                   double sig_d2, int want_grad, double* __restrict__ out5) {
  const double mm = (double)m;
  out5[0] = (-0.5 * scal[0] - scal[1] - mm * 0.5 * log(2.0 * 3.141592653589793)) / mm;
  out5[1] = want_grad ? 0.5 * (scal[2] + 2.0 * scal[3] + scal[4]) / mm : 0.0;
  out5[2] = want_grad ? sig_t * (scal[5] - scal[7]) / mm : 0.0;
  out5[3] = want_grad ? sig_d * ((scal[6] - scal[9]) - (scal[8] - scal[10])) / mm : 0.0;
  out5[4] = want_grad ? sig_d2 * (scal[9] - scal[10]) / mm : 0.0;
}
}  // namespace

// out[(N1 + N1 nv1) x (N2 + N2 nv2)] (ld ldo) = the block kernel matrix between (x1, J1) and (x2, J2);
// add_t / add_d: diagonal shifts of the TT / DD blocks (identical operands only).  lower_only (identical
// operands): the strict upper triangle of the TT and DD blocks and the whole TD block are left untouched.
int td_gram(Ctx* ctx, int kind, const double* x1, const double* J1, int N1, int nv1, const double* x2,
            const double* J2, int N2, int nv2, int nf, double ell, double add_t, double add_d, double* out,
            long long ldo, int lower_only, cudaStream_t s) {
  GPXB_ARG_CHECK(kind == KIND_SE || kind == KIND_M52, -50);
  const bool sym = (x1 == x2) && (J1 == J2) && (N1 == N2) && (nv1 == nv2);
  GPXB_ARG_CHECK(sym || (add_t == 0.0 && add_d == 0.0 && !lower_only), -70);
  const ZLayout zl = z_layout(nf);
  GPXB_ARG_CHECK(zl.S <= GRAM_MAX_S, -52);
  DevBuf bZ1, bZ2;
  GPXB_TRY(bZ1.alloc(sizeof(double) * (size_t)N1 * zl.S, s));
  GPXB_TRY(prep_z(ctx, x1, nf, N1, zl, ell, bZ1.as<double>(), s));
  const double* Z2 = bZ1.as<double>();
  if (!sym) {
    GPXB_TRY(bZ2.alloc(sizeof(double) * (size_t)N2 * zl.S, s));
    GPXB_TRY(prep_z(ctx, x2, nf, N2, zl, ell, bZ2.as<double>(), s));
    Z2 = bZ2.as<double>();
  }
  {
    GramArgs g;  // TT
    g.kind = kind; g.Z1 = bZ1.as<double>(); g.n1 = N1; g.Z2 = Z2; g.n2 = N2; g.zl = zl; g.ell = ell;
    g.diag_add = add_t; g.sym = sym ? 1 : 0; g.lower_only = lower_only;
    g.out = out; g.ldo = ldo;
    GPXB_TRY(gram_launch(ctx, g, s));
  }
  // DT: d0kj(x1, x2, J1)
  GPXB_TRY(d0kj_block(ctx, kind, x1, J1, N1, nv1, x2, N2, nf, ell, 0, out + (long long)N1 * ldo, ldo, 1, nullptr, nullptr,
                      nullptr, 0, nullptr, s));
  // TD: d1kj(x1, x2, J2) = d0kj(x2, x1, J2)^T
  if (!lower_only)
    GPXB_TRY(d0kj_block(ctx, kind, x2, J2, N2, nv2, x1, N1, nf, ell, 0, out + N2, 1, ldo, nullptr, nullptr, nullptr, 0,
                        nullptr, s));
  // DD
  return hess_gram(ctx, kind, x1, J1, N1, nv1, x2, J2, N2, nv2, nf, ell, add_d, out + (long long)N1 * ldo + N2, ldo,
                   lower_only, s);
}

// out5 = {lml, d/d lengthscale, d/d sigma_targets, d/d sigma_derivs, d/d sigma_derivs2}; c_out (optional, m) = A^-1 r;
// mu_out (optional).  nv_a: number of variables of the first derivative block (nv_a == nv: one block, sig_d2 unused).
int gpr_td(Ctx* ctx, int kind, const double* x, const double* J, const double* y, const double* yd, int N, int nf,
           int nv, int nv_a, double ell, double sig_t, double sig_d, double sig_d2, int mean_mode, int want_grad,
           double* out5, double* c_out, double* mu_out, cudaStream_t s) {
  const long long n = (long long)N * nv, m = N + n;
  GPXB_ARG_CHECK(N > 0 && nv > 0 && nf > 0 && m < (1LL << 31), -1);
  GPXB_ARG_CHECK(nv_a > 0 && nv_a <= nv, -2);
  const bool two = nv_a < nv;
  double* out4 = out5;
  DevBuf bA, bInv, bR, bSc, bMu, bW, bT;
  GPXB_TRY(bA.alloc(sizeof(double) * (size_t)m * m, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles((int)m), s));
  GPXB_TRY(bR.alloc(sizeof(double) * (size_t)m, s));
  GPXB_TRY(bSc.alloc(sizeof(double) * 16, s));
  GPXB_TRY(bMu.alloc(sizeof(double), s));
  double *A = bA.as<double>(), *inv = bInv.as<double>(), *r = bR.as<double>(), *scal = bSc.as<double>();
  double* mu = mu_out ? mu_out : bMu.as<double>();
  GPXB_CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double) * 16, s));
  mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu);
  GPXB_LAUNCHED(ctx);
  td_rhs_kernel<<<ceil_div(m, 256), 256, 0, s>>>(y, N, yd, n, mu, r);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(td_gram(ctx, kind, x, J, N, nv, x, J, N, nv, nf, ell, sig_t * sig_t + 1e-10, sig_d * sig_d + 1e-10, A, m, 1, s));
  if (two) {
    td_diag2_kernel<<<ceil_div(n, 256), 256, 0, s>>>(A + (long long)N * m + N, m, n, nv, nv_a,
                                                      sig_d2 * sig_d2 - sig_d * sig_d);
    GPXB_LAUNCHED(ctx);
  }
  GPXB_TRY(potrf_lower(ctx, A, m, (int)m, inv, s));
  if (out4) GPXB_TRY(logdet_lower(ctx, A, m, (int)m, scal + 1, 1.0, s));
  GPXB_TRY(trsm_right_lt(ctx, A, m, inv, r, m, 1, (int)m, s));
  sumsq_kernel<<<1, 1024, 0, s>>>(r, m, scal + 0, 0);
  GPXB_LAUNCHED(ctx);
  if (c_out || want_grad) GPXB_TRY(trsm_right_l(ctx, A, m, inv, r, m, 1, (int)m, s));  // r <- alpha
  if (c_out) GPXB_CUDA_CHECK(cudaMemcpyAsync(c_out, r, sizeof(double) * (size_t)m, cudaMemcpyDeviceToDevice, s));
  if (out4 && want_grad) {
    sumsq_kernel<<<1, 1024, 0, s>>>(r, N, scal + 5, 0);
    GPXB_LAUNCHED(ctx);
    sumsq_kernel<<<1, 1024, 0, s>>>(r + N, n, scal + 6, 0);
    GPXB_LAUNCHED(ctx);
    GPXB_TRY(bW.alloc(sizeof(double) * (size_t)m * m, s));
    GPXB_TRY(bT.alloc(sizeof(double) * potri_scratch_doubles((int)m), s));
    GPXB_TRY(potri_lower(ctx, A, m, inv, (int)m, bW.as<double>(), bT.as<double>(), A, m, s));  // A <- A^-1 (lower)
    bW.release();
    diag_sum_kernel<<<1, 1024, 0, s>>>(A, m, N, scal + 7);
    GPXB_LAUNCHED(ctx);
    diag_sum_kernel<<<1, 1024, 0, s>>>(A + (long long)N * m + N, m, (int)n, scal + 8);
    GPXB_LAUNCHED(ctx);
    if (two) {
      td_block2_sums_kernel<<<1, 1024, 0, s>>>(r + N, A + (long long)N * m + N, m, n, nv, nv_a, scal + 9);
      GPXB_LAUNCHED(ctx);
    }
    // TT block: the exact-GPR contraction epilogue of the Gram kernel on the leading N x N block
    const ZLayout zl = z_layout(nf);
    DevBuf bZ, bPart;
    GPXB_TRY(bZ.alloc(sizeof(double) * (size_t)N * zl.S, s));
    GPXB_TRY(prep_z(ctx, x, nf, N, zl, ell, bZ.as<double>(), s));
    GramArgs g;
    g.kind = kind; g.Z1 = bZ.as<double>(); g.n1 = N; g.Z2 = bZ.as<double>(); g.n2 = N; g.zl = zl; g.ell = ell;
    g.sym = 1; g.lower_only = 1; g.Ainv = A; g.ldainv = m; g.alpha = r; g.t = 1;
    const long long nparts = gram_num_ctas(g);
    GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)nparts, s));
    g.partials = bPart.as<double>();
    GPXB_TRY(gram_launch(ctx, g, s));
    sum_kernel<<<1, 1024, 0, s>>>(g.partials, nparts, scal + 2);
    GPXB_LAUNCHED(ctx);
    // DT block (counted twice in the symmetric sum): rows N.., columns 0..N of A^-1, weights alpha_D alpha_T^T
    GPXB_TRY(d0kj_block(ctx, kind, x, J, N, nv, x, N, nf, ell, 1, nullptr, 0, 0, r + N, r, A + (long long)N * m, m,
                        scal + 3, s));
    // DD block
    GPXB_TRY(hess_contract_dl(ctx, kind, x, J, N, nv, nf, ell, A + (long long)N * m + N, m, r + N, scal + 4, s));
  }
  if (out4) {
    td_finalize_kernel<<<1, 1, 0, s>>>(scal, m, sig_t, sig_d, two ? sig_d2 : 0.0, want_grad, out4);
    GPXB_LAUNCHED(ctx);
  }
  return 0;
}

}  // namespace gpxb

using gpxb::Ctx;

extern "C" {

int gpxb_td_gram(gpxb_ctx* ctx, int kind, const double* x1, const double* J1, int n1, int nv1, const double* x2,
                 const double* J2, int n2, int nv2, int nf, double ell, double add_targets, double add_derivs,
                 double* out, long long ldo, int lower_only, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x1 && J1 && x2 && J2 && out, -1);
  return gpxb::td_gram(reinterpret_cast<Ctx*>(ctx), kind, x1, J1, n1, nv1, x2, J2, n2, nv2, nf, ell, add_targets,
                       add_derivs, out, ldo, lower_only, (cudaStream_t)stream);
}

int gpxb_gpr_td_lml_grad(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y,
                         const double* y_derivs, int N, int nf, int nv, double ell, double sigma_targets,
                         double sigma_derivs, int mean_mode, int want_grad, double* out4, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && y_derivs && out4, -1);
  gpxb::DevBuf b5;  // the kernel writes five values; the one-block entry point exposes the first four
  cudaStream_t s = (cudaStream_t)stream;
  GPXB_TRY(b5.alloc(sizeof(double) * 5, s));
  GPXB_TRY(gpxb::gpr_td(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, y_derivs, N, nf, nv, nv, ell, sigma_targets,
                        sigma_derivs, 0.0, mean_mode, want_grad, b5.as<double>(), nullptr, nullptr, s));
  GPXB_CUDA_CHECK(cudaMemcpyAsync(out4, b5.as<double>(), sizeof(double) * 4, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int gpxb_gpr_td2_lml_grad(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y,
                          const double* y_derivs, int N, int nf, int nv, int nv_first, double ell, double sigma_targets,
                          double sigma_derivs, double sigma_derivs2, int mean_mode, int want_grad, double* out5,
                          gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && y_derivs && out5, -1);
  return gpxb::gpr_td(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, y_derivs, N, nf, nv, nv_first, ell, sigma_targets,
                      sigma_derivs, sigma_derivs2, mean_mode, want_grad, out5, nullptr, nullptr, (cudaStream_t)stream);
}

int gpxb_gpr_td2_fit(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, const double* y_derivs,
                     int N, int nf, int nv, int nv_first, double ell, double sigma_targets, double sigma_derivs,
                     double sigma_derivs2, int mean_mode, double* c_out, double* mu_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && y_derivs && c_out && mu_out, -1);
  return gpxb::gpr_td(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, y_derivs, N, nf, nv, nv_first, ell, sigma_targets,
                      sigma_derivs, sigma_derivs2, mean_mode, 0, nullptr, c_out, mu_out, (cudaStream_t)stream);
}

int gpxb_gpr_td_fit(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, const double* y_derivs,
                    int N, int nf, int nv, double ell, double sigma_targets, double sigma_derivs, int mean_mode,
                    double* c_out, double* mu_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && x && J && y && y_derivs && c_out && mu_out, -1);
  return gpxb::gpr_td(reinterpret_cast<Ctx*>(ctx), kind, x, J, y, y_derivs, N, nf, nv, nv, ell, sigma_targets,
                      sigma_derivs, 0.0, mean_mode, 0, nullptr, c_out, mu_out, (cudaStream_t)stream);
}

}  // extern "C"
