// Sparse GPR (Projected Processes) on device: Woodbury LML, fit, predict.
//
// Replaces gpx/models/_sgpr.py:_lml_dense (659-697), _A_lhs (39-61), _fit_dense (319-339),
// _predict_dense (434-467).  The reference's LML factors a dense N x N matrix
// C = G^T G + (sigma^2 + 1e-10) I; the identical value is obtained here in O(N M^2) with
//   B = s2 I_M + G G^T,  L_B = chol(B),  t = L_B^-1 G r,
//   quad = (r.r - t.t)/s2,  logdet = (N - M) log s2 + 2 sum log L_B,ii        
// K_mn is never stored: rows of x are streamed in chunks; each chunk's kernel block is built
// by the fused Gram kernel, solved against L_m and accumulated into B on the FP64 tensor pipe.
// Row-sharded: every rank streams its own rows; B, G r and r.r are all-reduced once.
#include "../../include/gpxb200.h"
#include "chol.cuh"
#include "comm.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "matvec.cuh"
#include "small_kernels.cuh"

namespace gpxb {

namespace {
using namespace sk;

constexpr int CHUNK = 32768;

// scal: [0] r.r, [4] t.t, [5] sum log L_B,ii
__global__ void finalize_sgpr_kernel(const double* __restrict__ scal, long long N, int M, double s2,
                                     double* __restrict__ out) {
  const double n = (double)N;
  const double quad = (scal[0] - scal[4]) / s2;
  const double logdet = (n - (double)M) * log(s2) + 2.0 * scal[5];
  out[0] = (-0.5 * quad - 0.5 * logdet - n * 0.5 * log(2.0 * 3.141592653589793)) / n;
}
}  // namespace

// x: this rank's rows (n_loc x D), y: (n_loc x T); N: global row count; mu: device scalar with
// the GLOBAL mean (computed by the caller).  out: device scalar lml / N.
// The row chunks are independent until the final reduction, so they are pipelined round-robin over
// three streams with private buffers and private B / v / r.r accumulators (summed in a fixed order at
// the end): the tails of one chunk's kernels (528 lower tiles = 3.6 waves of the SYRK) overlap the next
// chunk's work.
int sgpr_lml(Ctx* ctx, int kind, const double* x, const double* y, int n_loc, long long N, int D, int T,
             const double* x_locs, int M, double ell, double sigma, const double* mu, double* out,
             cudaStream_t s) {
  GPXB_ARG_CHECK(n_loc >= 0 && D > 0 && T > 0 && M > 0, -1);
  const ZLayout zl = z_layout(D);
  const double s2 = sigma * sigma + 1e-10;
  const int nc_max = std::min(CHUNK, std::max(n_loc, 1));
  const int nchunks = ceil_div(std::max(n_loc, 1), CHUNK);
  const int NS = nchunks >= 3 ? 3 : 1;  // concurrent chunk pipelines
  const size_t mm = (size_t)M * M;
  DevBuf bZl, bKmm, bInv, bB, bV, bZc, bKc, bRc, bSc, bInvB;
  GPXB_TRY(bZl.alloc(sizeof(double) * (size_t)M * zl.S, s));
  GPXB_TRY(bKmm.alloc(sizeof(double) * mm, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  GPXB_TRY(bB.alloc(sizeof(double) * mm * NS, s));
  GPXB_TRY(bV.alloc(sizeof(double) * (size_t)M * T * NS, s));
  GPXB_TRY(bZc.alloc(sizeof(double) * (size_t)nc_max * zl.S * NS, s));
  GPXB_TRY(bKc.alloc(sizeof(double) * (size_t)nc_max * M * NS, s));
  GPXB_TRY(bRc.alloc(sizeof(double) * (size_t)nc_max * T * NS, s));
  GPXB_TRY(bSc.alloc(sizeof(double) * 16, s));
  GPXB_TRY(bInvB.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  double *Zl = bZl.as<double>(), *Kmm = bKmm.as<double>(), *inv = bInv.as<double>(), *B = bB.as<double>();
  double *v = bV.as<double>(), *scal = bSc.as<double>();  // scal[0..2]: r.r per pipeline, [4] t.t, [5] sum log
  GPXB_CUDA_CHECK(cudaMemsetAsync(B, 0, sizeof(double) * mm * NS, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * (size_t)M * T * NS, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double) * 16, s));

  // L_m = chol(K_mm + 1e-10 I)
  {
    PhaseScope ph(ctx, PH_POTRF_REPL, s);
    GPXB_TRY(prep_z(ctx, x_locs, D, M, zl, ell, Zl, s));
    GramArgs g;
    g.kind = kind; g.Z1 = Zl; g.n1 = M; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell;
    g.diag_add = 1e-10; g.sym = 1; g.lower_only = 1; g.out = Kmm; g.ldo = M;
    GPXB_TRY(gram_launch(ctx, g, s));
    GPXB_TRY(potrf_lower(ctx, Kmm, M, M, inv, s));
  }

  cudaStream_t st[3] = {s, ctx->aux[0], ctx->aux[1]};
  if (NS > 1) {
    cudaEvent_t e;
    GPXB_TRY(ctx->next_event(&e));
    GPXB_CUDA_CHECK(cudaEventRecord(e, s));
    for (int k = 1; k < NS; ++k) GPXB_CUDA_CHECK(cudaStreamWaitEvent(st[k], e, 0));
  }
  int ci = 0;
  for (int r0 = 0; r0 < n_loc; r0 += CHUNK, ++ci) {
    const int nc = std::min(CHUNK, n_loc - r0);
    const int k = ci % NS;
    cudaStream_t sk = st[k];
    double* Zc = bZc.as<double>() + (size_t)k * nc_max * zl.S;
    double* Kc = bKc.as<double>() + (size_t)k * nc_max * M;
    double* rc = bRc.as<double>() + (size_t)k * nc_max * T;
    PhaseScope ph(ctx, PH_SGPR_CHUNK, sk);
    GPXB_TRY(prep_z(ctx, x + (long long)r0 * D, D, nc, zl, ell, Zc, sk));
    {
      GramArgs g;  // Kc = K(x_chunk, x_locs)  [nc x M]
      g.kind = kind; g.Z1 = Zc; g.n1 = nc; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell;
      g.out = Kc; g.ldo = M;
      GPXB_TRY(gram_launch(ctx, g, sk));
    }
    GPXB_TRY(trsm_right_lt(ctx, Kmm, M, inv, Kc, M, nc, M, sk));  // Kc <- G_c^T = Kc L_m^-T
    center_kernel<<<ceil_div((long long)nc * T, 256), 256, 0, sk>>>(y + (long long)r0 * T, (long long)nc * T, mu, rc);
    GPXB_LAUNCHED(ctx);
    sumsq_kernel<<<1, 1024, 0, sk>>>(rc, (long long)nc * T, scal + k, 1);
    GPXB_LAUNCHED(ctx);
    {
      GemmArgs g;  // B_k += G_c G_c^T  (A, B stored [K = nc][M])
      g.M = M; g.N = M; g.K = nc;
      g.A = Kc; g.lda = M; g.a_kmajor = false;
      g.B = Kc; g.ldb = M; g.b_kmajor = false;
      g.C = B + (size_t)k * mm; g.ldc = M; g.beta = 1.0; g.lower_only = 1;
      GPXB_TRY(gemm_f64(ctx, g, sk));
    }
    {
      GemmArgs g;  // v_k += G_c r_c   [M x T]
      g.M = M; g.N = T; g.K = nc;
      g.A = Kc; g.lda = M; g.a_kmajor = false;
      g.B = rc; g.ldb = T; g.b_kmajor = false;
      g.C = v + (size_t)k * M * T; g.ldc = T; g.beta = 1.0;
      GPXB_TRY(gemm_f64(ctx, g, sk));
    }
  }
  for (int k = 1; k < NS; ++k) {  // join, then reduce the private accumulators in a fixed order
    cudaEvent_t e;
    GPXB_TRY(ctx->next_event(&e));
    GPXB_CUDA_CHECK(cudaEventRecord(e, st[k]));
    GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, e, 0));
  }
  for (int k = 1; k < NS; ++k) {
    add_vec_kernel<<<ceil_div((long long)mm, 256), 256, 0, s>>>(B, B + (size_t)k * mm, (long long)mm);
    GPXB_LAUNCHED(ctx);
    add_vec_kernel<<<ceil_div((long long)M * T, 256), 256, 0, s>>>(v, v + (size_t)k * M * T, (long long)M * T);
    GPXB_LAUNCHED(ctx);
    add_vec_kernel<<<1, 1, 0, s>>>(scal, scal + k, 1);
    GPXB_LAUNCHED(ctx);
  }
  if (ctx->world > 1) {
    // upper triangle of B is untouched zero, so the full buffer can be summed
    GPXB_TRY(comm_allreduce_sum(ctx, B, mm, s));
    GPXB_TRY(comm_allreduce_sum(ctx, v, (size_t)M * T, s));
    GPXB_TRY(comm_allreduce_sum(ctx, scal, 1, s));
  }
  PhaseScope ph(ctx, PH_POTRF_REPL, s);
  add_diag_kernel<<<ceil_div(M, 256), 256, 0, s>>>(B, M, M, s2);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(potrf_lower(ctx, B, M, M, bInvB.as<double>(), s));
  GPXB_TRY(logdet_lower(ctx, B, M, M, scal + 5, 1.0, s));
  // t^T = v^T L_B^-T  (v is M x T; its transpose T x M is needed)
  DevBuf bVt;
  GPXB_TRY(bVt.alloc(sizeof(double) * (size_t)M * T, s));
  transpose_out_kernel<<<ceil_div((long long)M * T, 256), 256, 0, s>>>(v, T, M, bVt.as<double>());
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(trsm_right_lt(ctx, B, M, bInvB.as<double>(), bVt.as<double>(), M, T, M, s));
  sumsq_kernel<<<1, 1024, 0, s>>>(bVt.as<double>(), (long long)M * T, scal + 4, 0);
  GPXB_LAUNCHED(ctx);
  finalize_sgpr_kernel<<<1, 1, 0, s>>>(scal, N, M, s2, out);
  GPXB_LAUNCHED(ctx);
  return 0;
}

namespace {
// scal: [0] r.r  [1] u.what  [2] sum log diag L_B  [3] tr(B^-1)  [4] alpha.alpha
//       [5] g1 = sum dK_nm (alpha_n w_m - Y_nm)   [6] g2 = sum dA_mm' (w_m w_m' - F_mm')
__global__ void finalize_sgpr_grad_kernel(const double* __restrict__ scal, long long N, int M, double s2,
                                          double sigma, int want_grad, double* __restrict__ out3) {
  const double n = (double)N;
  const double quad = (scal[0] - scal[1]) / s2;
  const double logdet = (n - (double)M) * log(s2) + 2.0 * scal[2];
  out3[0] = (-0.5 * quad - 0.5 * logdet - n * 0.5 * log(2.0 * 3.141592653589793)) / n;
  if (want_grad) {
    const double trCinv = (n - (double)M + s2 * scal[3]) / s2;
    out3[1] = (scal[5] - 0.5 * scal[6]) / n;
    out3[2] = sigma * (scal[4] - trCinv) / n;
  } else {
    out3[1] = out3[2] = 0.0;
  }
}
// alpha_c = (r_c - t_c) / s2 in place on t_c
__global__ void alpha_chunk_kernel(const double* __restrict__ r, double* __restrict__ t, long long n, double s2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = (r[i] - t[i]) / s2;
}
// H = I - s2 * Binv (full M x M)
__global__ void i_minus_scaled_kernel(const double* __restrict__ Binv, int M, double s2, double* __restrict__ H) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * M) return;
  const int i = (int)(idx / M), j = (int)(idx % M);
  H[idx] = (i == j ? 1.0 : 0.0) - s2 * Binv[idx];
}
__global__ void transpose_sq_kernel(const double* __restrict__ in, int M, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + threadIdx.x;
    tile[k][threadIdx.x] = (r < M && c < M) ? in[(long long)r * M + c] : 0.0;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = c0 + k, c = r0 + threadIdx.x;
    if (r < M && c < M) out[(long long)r * M + c] = tile[threadIdx.x][k];
  }
}
// V[i] = (1, x[i][0..D))
__global__ void ones_x_kernel(const double* __restrict__ x, int n, int D, double* __restrict__ V) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * (D + 1)) return;
  const int i = (int)(idx / (D + 1)), c = (int)(idx % (D + 1));
  V[idx] = c == 0 ? 1.0 : x[(long long)i * D + c - 1];
}
// grad[m][f] = scale (-(z_mf acc1[m][0] - acc1[m][1+f]) + (z_mf acc2[m][0] - acc2[m][1+f]))
__global__ void grad_locs_kernel(const double* __restrict__ acc1, const double* __restrict__ acc2,
                                 const double* __restrict__ z, int M, int D, double scale,
                                 double* __restrict__ grad) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * D) return;
  const int m = (int)(idx / D), f = (int)(idx % D);
  const double zf = z[idx];
  const double* a1 = acc1 + (long long)m * (D + 1);
  const double* a2 = acc2 + (long long)m * (D + 1);
  grad[idx] = scale * (-(zf * a1[0] - a1[1 + f]) + (zf * a2[0] - a2[1 + f]));
}
}  // namespace

// SGPR LML and its gradient w.r.t. (lengthscale, sigma) with fixed landmarks -- what
// jit(value_and_grad) of _sgpr._lml_dense yields, in O(N M^2).  Everything is
// expressed through the well-conditioned B = s2 I + G G^T (eigenvalues >= s2), G = L_A^-1 K_mn,
// A = K_mm + 1e-10 I = L_A L_A^T:
//   what = B^-1 G r,  w = L_A^-T what,  alpha = (r - G^T what)/s2
//   lml N = -1/2 (r.r - (G r).what)/s2 - 1/2 [(N-M) log s2 + log|B|] - N/2 log 2pi
//   d/d ell  = sum_nm dK_nm (alpha_n w_m - Y_nm) - 1/2 sum_mm' dA_mm' (w_m w_m' - F_mm'),
//              Y = G^T B^-1 L_A^-1 (per chunk),  F = L_A^-T (I - s2 B^-1) L_A^-1
//   d/d sigma = sigma (alpha.alpha - tr C^-1),  tr C^-1 = (N - M + s2 tr(B^-1))/s2
// Single output column (T = 1).  out3: device [lml/N, d/d ell, d/d sigma].
//
// grad_locs (optional, M x D): d(lml/N)/d x_locs for trainable landmark coordinates.  With
// d k(x, z)/d z_f = -c g(x, z) (z_f - x_f) (g: radial derivative profile, matvec.cu KMODE 3) the same weights give
//   grad[m][f] = -c sum_n g_nm (alpha_n w_m - Y_nm) (z_mf - x_nf) + c sum_m' g_mm' (w_m w_m' - F_mm') (z_mf - z_m'f)
// = two weighted profile products H [1 | X] of the matrix-free kernel (no M x N x D temporary).
int sgpr_lml_grad(Ctx* ctx, int kind, const double* x, const double* y, int n_loc, long long N, int D,
                  const double* x_locs, int M, double ell, double sigma, const double* mu, int want_grad,
                  double* out3, double* grad_locs, cudaStream_t s) {
  GPXB_ARG_CHECK(n_loc >= 0 && D > 0 && M > 0, -1);
  GPXB_ARG_CHECK(!grad_locs || (want_grad && kmatvec_supports(D)), -4);
  const ZLayout zl = z_layout(D);
  const double s2 = sigma * sigma + 1e-10;
  const size_t mm = (size_t)M * M;
  const int nc_max = std::min(CHUNK, std::max(n_loc, 1));
  const int nchunks = ceil_div(std::max(n_loc, 1), CHUNK);
  const int NS = nchunks >= 3 ? 3 : 1;  // concurrent chunk pipelines, as in sgpr_lml
  const int P1 = D + 1;
  DevBuf bZl, bA, bInvA, bB, bInvB, bU, bWh, bW, bZc, bKc, bRc, bTc, bY, bSc, bWk, bTs, bH, bHt;
  DevBuf bVx, bAcc1, bAcc2, bTmp;
  GPXB_TRY(bZl.alloc(sizeof(double) * (size_t)M * zl.S, s));
  GPXB_TRY(bA.alloc(sizeof(double) * mm, s));
  GPXB_TRY(bInvA.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  GPXB_TRY(bB.alloc(sizeof(double) * mm * NS, s));
  GPXB_TRY(bInvB.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  GPXB_TRY(bU.alloc(sizeof(double) * M * NS, s));
  GPXB_TRY(bWh.alloc(sizeof(double) * M, s));
  GPXB_TRY(bW.alloc(sizeof(double) * M, s));
  GPXB_TRY(bZc.alloc(sizeof(double) * (size_t)nc_max * zl.S * NS, s));
  GPXB_TRY(bKc.alloc(sizeof(double) * (size_t)nc_max * M * NS, s));
  GPXB_TRY(bRc.alloc(sizeof(double) * (size_t)nc_max * NS, s));
  GPXB_TRY(bTc.alloc(sizeof(double) * (size_t)nc_max * NS, s));
  GPXB_TRY(bSc.alloc(sizeof(double) * 8 * (NS + 1), s));
  double *Zl = bZl.as<double>(), *A = bA.as<double>(), *B = bB.as<double>(), *invA = bInvA.as<double>();
  double *u = bU.as<double>(), *wh = bWh.as<double>(), *w = bW.as<double>();
  // scal: the combined scalars (layout of finalize_sgpr_grad_kernel); pscal[k]: pipeline k's partial sums
  double* scal = bSc.as<double>();
  auto pscal = [&](int k) { return scal + 8 * (k + 1); };
  GPXB_CUDA_CHECK(cudaMemsetAsync(B, 0, sizeof(double) * mm * NS, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(u, 0, sizeof(double) * M * NS, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double) * 8 * (NS + 1), s));
  const int eb = ceil_div((long long)mm, 256);

  // L_A = chol(K_mm + 1e-10 I)
  GPXB_TRY(prep_z(ctx, x_locs, D, M, zl, ell, Zl, s));
  {
    GramArgs g;
    g.kind = kind; g.Z1 = Zl; g.n1 = M; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell;
    g.diag_add = 1e-10; g.sym = 1; g.lower_only = 1; g.out = A; g.ldo = M;
    GPXB_TRY(gram_launch(ctx, g, s));
  }
  GPXB_TRY(potrf_lower(ctx, A, M, M, invA, s));

  cudaStream_t st[3] = {s, ctx->aux[0], ctx->aux[1]};
  auto fork = [&]() -> int {
    if (NS == 1) return 0;
    cudaEvent_t e;
    GPXB_TRY(ctx->next_event(&e));
    GPXB_CUDA_CHECK(cudaEventRecord(e, s));
    for (int k = 1; k < NS; ++k) GPXB_CUDA_CHECK(cudaStreamWaitEvent(st[k], e, 0));
    return 0;
  };
  auto join = [&]() -> int {
    for (int k = 1; k < NS; ++k) {
      cudaEvent_t e;
      GPXB_TRY(ctx->next_event(&e));
      GPXB_CUDA_CHECK(cudaEventRecord(e, st[k]));
      GPXB_CUDA_CHECK(cudaStreamWaitEvent(s, e, 0));
    }
    return 0;
  };
  struct Pipe {
    double *Zc, *Kc, *rc, *tc;
    cudaStream_t sk;
  };
  auto pipe = [&](int k) {
    Pipe p;
    p.Zc = bZc.as<double>() + (size_t)k * nc_max * zl.S;
    p.Kc = bKc.as<double>() + (size_t)k * nc_max * M;
    p.rc = bRc.as<double>() + (size_t)k * nc_max;
    p.tc = bTc.as<double>() + (size_t)k * nc_max;
    p.sk = st[k];
    return p;
  };
  auto chunk_G = [&](const Pipe& p, int r0, int nc) -> int {  // Kc <- G_c^T = K(x_chunk, locs) L_A^-T ; rc <- r_c
    GPXB_TRY(prep_z(ctx, x + (long long)r0 * D, D, nc, zl, ell, p.Zc, p.sk));
    GramArgs g;
    g.kind = kind; g.Z1 = p.Zc; g.n1 = nc; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell; g.out = p.Kc; g.ldo = M;
    GPXB_TRY(gram_launch(ctx, g, p.sk));
    GPXB_TRY(trsm_right_lt(ctx, A, M, invA, p.Kc, M, nc, M, p.sk));
    center_kernel<<<ceil_div(nc, 256), 256, 0, p.sk>>>(y + r0, nc, mu, p.rc);
    GPXB_LAUNCHED(ctx);
    return 0;
  };

  // pass 1: B_k += G_c G_c^T, u_k += G_c r_c, r.r (private accumulators per pipeline, summed in a fixed order)
  GPXB_TRY(fork());
  int ci = 0;
  for (int r0 = 0; r0 < n_loc; r0 += CHUNK, ++ci) {
    const int nc = std::min(CHUNK, n_loc - r0);
    const int k = ci % NS;
    const Pipe p = pipe(k);
    GPXB_TRY(chunk_G(p, r0, nc));
    sumsq_kernel<<<1, 1024, 0, p.sk>>>(p.rc, nc, pscal(k) + 0, 1);
    GPXB_LAUNCHED(ctx);
    GemmArgs a;
    a.M = M; a.N = M; a.K = nc;
    a.A = p.Kc; a.lda = M; a.a_kmajor = false;
    a.B = p.Kc; a.ldb = M; a.b_kmajor = false;
    a.C = B + (size_t)k * mm; a.ldc = M; a.beta = 1.0; a.lower_only = 1;
    GPXB_TRY(gemm_f64(ctx, a, p.sk));
    GemmArgs b;
    b.M = M; b.N = 1; b.K = nc;
    b.A = p.Kc; b.lda = M; b.a_kmajor = false;
    b.B = p.rc; b.ldb = 1; b.b_kmajor = false;
    b.C = u + (size_t)k * M; b.ldc = 1; b.beta = 1.0;
    GPXB_TRY(gemm_f64(ctx, b, p.sk));
  }
  GPXB_TRY(join());
  for (int k = 1; k < NS; ++k) {
    add_vec_kernel<<<eb, 256, 0, s>>>(B, B + (size_t)k * mm, (long long)mm);
    GPXB_LAUNCHED(ctx);
    add_vec_kernel<<<ceil_div(M, 256), 256, 0, s>>>(u, u + (size_t)k * M, M);
    GPXB_LAUNCHED(ctx);
  }
  for (int k = 0; k < NS; ++k) {
    add_vec_kernel<<<1, 1, 0, s>>>(scal + 0, pscal(k) + 0, 1);
    GPXB_LAUNCHED(ctx);
  }
  if (ctx->world > 1) {
    GPXB_TRY(comm_allreduce_sum(ctx, B, mm, s));
    GPXB_TRY(comm_allreduce_sum(ctx, u, M, s));
    GPXB_TRY(comm_allreduce_sum(ctx, scal, 1, s));
  }
  add_diag_kernel<<<ceil_div(M, 256), 256, 0, s>>>(B, M, M, s2);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(potrf_lower(ctx, B, M, M, bInvB.as<double>(), s));
  GPXB_TRY(logdet_lower(ctx, B, M, M, scal + 2, 1.0, s));
  GPXB_CUDA_CHECK(cudaMemcpyAsync(wh, u, sizeof(double) * M, cudaMemcpyDeviceToDevice, s));
  GPXB_TRY(trsm_right_lt(ctx, B, M, bInvB.as<double>(), wh, M, 1, M, s));
  GPXB_TRY(trsm_right_l(ctx, B, M, bInvB.as<double>(), wh, M, 1, M, s));  // what = B^-1 u
  dot_kernel<<<1, 1024, 0, s>>>(u, wh, M, scal + 1, 0);
  GPXB_LAUNCHED(ctx);

  if (want_grad) {
    GPXB_TRY(bWk.alloc(sizeof(double) * mm, s));
    GPXB_TRY(bTs.alloc(sizeof(double) * potri_scratch_doubles(M), s));
    GPXB_TRY(bY.alloc(sizeof(double) * (size_t)nc_max * M * NS, s));
    GPXB_TRY(bH.alloc(sizeof(double) * mm, s));
    GPXB_TRY(bHt.alloc(sizeof(double) * mm, s));
    double *H = bH.as<double>(), *Ht = bHt.as<double>();
    if (grad_locs) {
      GPXB_TRY(bVx.alloc(sizeof(double) * (size_t)std::max(nc_max, M) * P1 * NS, s));
      GPXB_TRY(bAcc1.alloc(sizeof(double) * (size_t)M * P1 * NS, s));
      GPXB_TRY(bAcc2.alloc(sizeof(double) * (size_t)M * P1, s));
      GPXB_TRY(bTmp.alloc(sizeof(double) * (size_t)M * P1 * NS, s));
      GPXB_CUDA_CHECK(cudaMemsetAsync(bAcc1.as<double>(), 0, sizeof(double) * (size_t)M * P1 * NS, s));
    }
    // w = L_A^-T what (landmark weights)
    GPXB_CUDA_CHECK(cudaMemcpyAsync(w, wh, sizeof(double) * M, cudaMemcpyDeviceToDevice, s));
    GPXB_TRY(trsm_right_l(ctx, A, M, invA, w, M, 1, M, s));
    // B <- B^-1 (full symmetric)
    GPXB_TRY(potri_lower(ctx, B, M, bInvB.as<double>(), M, bWk.as<double>(), bTs.as<double>(), B, M, s));
    diag_sum_kernel<<<1, 1024, 0, s>>>(B, M, M, scal + 3);  // tr(B^-1)
    GPXB_LAUNCHED(ctx);
    symmetrize_lower_kernel<<<eb, 256, 0, s>>>(B, M);
    GPXB_LAUNCHED(ctx);
    // pass 2: per chunk Y = G_c^T B^-1 L_A^-1, alpha_c = (r_c - G_c^T what)/s2, g1 += sum dK (alpha w^T - Y)
    GPXB_TRY(fork());
    ci = 0;
    for (int r0 = 0; r0 < n_loc; r0 += CHUNK, ++ci) {
      const int nc = std::min(CHUNK, n_loc - r0);
      const int k = ci % NS;
      const Pipe p = pipe(k);
      double* Y = bY.as<double>() + (size_t)k * nc_max * M;
      GPXB_TRY(chunk_G(p, r0, nc));
      GemmArgs b;  // t_c = G_c^T what
      b.M = nc; b.N = 1; b.K = M;
      b.A = p.Kc; b.lda = M; b.a_kmajor = true;
      b.B = wh; b.ldb = 1; b.b_kmajor = false;
      b.C = p.tc; b.ldc = 1;
      GPXB_TRY(gemm_f64(ctx, b, p.sk));
      alpha_chunk_kernel<<<ceil_div(nc, 256), 256, 0, p.sk>>>(p.rc, p.tc, nc, s2);
      GPXB_LAUNCHED(ctx);
      sumsq_kernel<<<1, 1024, 0, p.sk>>>(p.tc, nc, pscal(k) + 4, 1);
      GPXB_LAUNCHED(ctx);
      GemmArgs a;  // Y = G_c^T B^-1
      a.M = nc; a.N = M; a.K = M;
      a.A = p.Kc; a.lda = M; a.a_kmajor = true;
      a.B = B; a.ldb = M; a.b_kmajor = false;
      a.C = Y; a.ldc = M;
      GPXB_TRY(gemm_f64(ctx, a, p.sk));
      GPXB_TRY(trsm_right_l(ctx, A, M, invA, Y, M, nc, M, p.sk));  // Y <- Y L_A^-1
      GramArgs c;
      c.kind = kind; c.Z1 = p.Zc; c.n1 = nc; c.Z2 = Zl; c.n2 = M; c.zl = zl; c.ell = ell;
      c.Ainv = Y; c.ldainv = M; c.alpha = p.tc; c.alpha2 = w; c.t = 1;
      const long long nparts = gram_num_ctas(c);
      DevBuf bPart;
      GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)(nparts + 1), p.sk));
      c.partials = bPart.as<double>();
      GPXB_TRY(gram_launch(ctx, c, p.sk));
      sum_kernel<<<1, 1024, 0, p.sk>>>(c.partials, nparts, c.partials + nparts);
      GPXB_LAUNCHED(ctx);
      add_vec_kernel<<<1, 1, 0, p.sk>>>(pscal(k) + 5, c.partials + nparts, 1);
      GPXB_LAUNCHED(ctx);
      if (grad_locs) {  // acc1_k += (g o (w alpha_c^T - Y^T)) [1 | x_c]   (rows: landmarks, columns: chunk points)
        double* Vx = bVx.as<double>() + (size_t)k * std::max(nc_max, M) * P1;
        double* tmp = bTmp.as<double>() + (size_t)k * M * P1;
        ones_x_kernel<<<ceil_div((long long)nc * P1, 256), 256, 0, p.sk>>>(x + (long long)r0 * D, nc, D, Vx);
        GPXB_LAUNCHED(ctx);
        GPXB_TRY(kprofile_weighted(ctx, kind, Zl, M, -1, p.Zc, nc, zl, Vx, P1, P1, w, p.tc, Y, M, 1, tmp, P1, p.sk));
        add_vec_kernel<<<ceil_div((long long)M * P1, 256), 256, 0, p.sk>>>(bAcc1.as<double>() + (size_t)k * M * P1, tmp,
                                                                          (long long)M * P1);
        GPXB_LAUNCHED(ctx);
      }
    }
    GPXB_TRY(join());
    for (int k = 0; k < NS; ++k) {
      add_vec_kernel<<<1, 2, 0, s>>>(scal + 4, pscal(k) + 4, 2);  // alpha.alpha and g1
      GPXB_LAUNCHED(ctx);
      if (grad_locs && k > 0) {
        add_vec_kernel<<<ceil_div((long long)M * P1, 256), 256, 0, s>>>(bAcc1.as<double>(),
                                                                         bAcc1.as<double>() + (size_t)k * M * P1,
                                                                         (long long)M * P1);
        GPXB_LAUNCHED(ctx);
      }
    }
    if (ctx->world > 1) GPXB_TRY(comm_allreduce_sum(ctx, scal + 4, 2, s));
    if (grad_locs && ctx->world > 1) GPXB_TRY(comm_allreduce_sum(ctx, bAcc1.as<double>(), (size_t)M * P1, s));
    // F = L_A^-T (I - s2 B^-1) L_A^-1:  H = I - s2 B^-1;  H <- H L_A^-1;  F = (H^T L_A^-1)^T = H^T L_A^-1
    i_minus_scaled_kernel<<<eb, 256, 0, s>>>(B, M, s2, H);
    GPXB_LAUNCHED(ctx);
    GPXB_TRY(trsm_right_l(ctx, A, M, invA, H, M, M, M, s));
    {
      dim3 grid(ceil_div(M, 32), ceil_div(M, 32)), blk(32, 8);
      transpose_sq_kernel<<<grid, blk, 0, s>>>(H, M, Ht);
      GPXB_LAUNCHED(ctx);
    }
    GPXB_TRY(trsm_right_l(ctx, A, M, invA, Ht, M, M, M, s));  // Ht = F (symmetric, full)
    GramArgs c;
    c.kind = kind; c.Z1 = Zl; c.n1 = M; c.Z2 = Zl; c.n2 = M; c.zl = zl; c.ell = ell; c.sym = 1; c.lower_only = 1;
    c.Ainv = Ht; c.ldainv = M; c.alpha = w; c.t = 1;
    const long long nparts = gram_num_ctas(c);
    DevBuf bPart;
    GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)nparts, s));
    c.partials = bPart.as<double>();
    GPXB_TRY(gram_launch(ctx, c, s));
    sum_kernel<<<1, 1024, 0, s>>>(c.partials, nparts, scal + 6);
    GPXB_LAUNCHED(ctx);
    if (grad_locs) {  // acc2 = (g o (w w^T - F)) [1 | z]   (rows = columns = landmarks)
      ones_x_kernel<<<ceil_div((long long)M * P1, 256), 256, 0, s>>>(x_locs, M, D, bVx.as<double>());
      GPXB_LAUNCHED(ctx);
      GPXB_TRY(kprofile_weighted(ctx, kind, Zl, M, 0, Zl, M, zl, bVx.as<double>(), P1, P1, w, w, Ht, M, 0,
                                 bAcc2.as<double>(), P1, s));
      const double cprof = kind == KIND_SE ? 2.0 : (kind == KIND_M12 ? 1.0 : (kind == KIND_M32 ? 3.0 : 5.0 / 3.0));
      grad_locs_kernel<<<ceil_div((long long)M * D, 256), 256, 0, s>>>(bAcc1.as<double>(), bAcc2.as<double>(), x_locs,
                                                                      M, D, cprof / (ell * ell) / (double)N,
                                                                      grad_locs);
      GPXB_LAUNCHED(ctx);
    }
  }
  finalize_sgpr_grad_kernel<<<1, 1, 0, s>>>(scal, N, M, s2, sigma, want_grad, out3);
  GPXB_LAUNCHED(ctx);
  return 0;
}

// c[M x T] = (sigma^2 K_mm + K_mn K_nm + 1e-10 I)^-1 K_mn (y - mu)      (_sgpr.py:39-61, 319-339)
int sgpr_fit(Ctx* ctx, int kind, const double* x, const double* y, int n_loc, int D, int T,
             const double* x_locs, int M, double ell, double sigma, const double* mu, double* c_out,
             cudaStream_t s) {
  GPXB_ARG_CHECK(n_loc >= 0 && D > 0 && T > 0 && M > 0, -1);
  const ZLayout zl = z_layout(D);
  const int nc_max = std::min(CHUNK, std::max(n_loc, 1));
  DevBuf bZl, bC, bInv, bV, bZc, bKc, bRc, bVt;
  GPXB_TRY(bZl.alloc(sizeof(double) * (size_t)M * zl.S, s));
  GPXB_TRY(bC.alloc(sizeof(double) * (size_t)M * M, s));
  GPXB_TRY(bInv.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  GPXB_TRY(bV.alloc(sizeof(double) * (size_t)M * T, s));
  GPXB_TRY(bVt.alloc(sizeof(double) * (size_t)M * T, s));
  GPXB_TRY(bZc.alloc(sizeof(double) * (size_t)nc_max * zl.S, s));
  GPXB_TRY(bKc.alloc(sizeof(double) * (size_t)nc_max * M, s));
  GPXB_TRY(bRc.alloc(sizeof(double) * (size_t)nc_max * T, s));
  double *Zl = bZl.as<double>(), *C = bC.as<double>(), *v = bV.as<double>();
  double *Zc = bZc.as<double>(), *Kc = bKc.as<double>(), *rc = bRc.as<double>();
  GPXB_CUDA_CHECK(cudaMemsetAsync(C, 0, sizeof(double) * (size_t)M * M, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * (size_t)M * T, s));
  GPXB_TRY(prep_z(ctx, x_locs, D, M, zl, ell, Zl, s));
  for (int r0 = 0; r0 < n_loc; r0 += CHUNK) {
    const int nc = std::min(CHUNK, n_loc - r0);
    GPXB_TRY(prep_z(ctx, x + (long long)r0 * D, D, nc, zl, ell, Zc, s));
    GramArgs g;
    g.kind = kind; g.Z1 = Zc; g.n1 = nc; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell;
    g.out = Kc; g.ldo = M;
    GPXB_TRY(gram_launch(ctx, g, s));
    center_kernel<<<ceil_div((long long)nc * T, 256), 256, 0, s>>>(y + (long long)r0 * T, (long long)nc * T, mu, rc);
    GPXB_LAUNCHED(ctx);
    GemmArgs a;  // C += K_c^T K_c
    a.M = M; a.N = M; a.K = nc;
    a.A = Kc; a.lda = M; a.a_kmajor = false;
    a.B = Kc; a.ldb = M; a.b_kmajor = false;
    a.C = C; a.ldc = M; a.beta = 1.0; a.lower_only = 1;
    GPXB_TRY(gemm_f64(ctx, a, s));
    GemmArgs b;  // v += K_c^T r_c
    b.M = M; b.N = T; b.K = nc;
    b.A = Kc; b.lda = M; b.a_kmajor = false;
    b.B = rc; b.ldb = T; b.b_kmajor = false;
    b.C = v; b.ldc = T; b.beta = 1.0;
    GPXB_TRY(gemm_f64(ctx, b, s));
  }
  if (ctx->world > 1) {
    GPXB_TRY(comm_allreduce_sum(ctx, C, (size_t)M * M, s));
    GPXB_TRY(comm_allreduce_sum(ctx, v, (size_t)M * T, s));
  }
  {
    // C += sigma^2 K_mm + 1e-10 I : build K_mm into scratch (reuse Kc: nc_max*M >= ? not
    // guaranteed) -> dedicated buffer
    DevBuf bK;
    GPXB_TRY(bK.alloc(sizeof(double) * (size_t)M * M, s));
    GPXB_CUDA_CHECK(cudaMemsetAsync(bK.as<double>(), 0, sizeof(double) * (size_t)M * M, s));
    GramArgs g;
    g.kind = kind; g.Z1 = Zl; g.n1 = M; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell;
    g.sym = 1; g.lower_only = 1; g.out = bK.as<double>(); g.ldo = M;
    GPXB_TRY(gram_launch(ctx, g, s));
    // C = C + sigma^2 * K_mm  via GEMM-free axpy: scale then add with a tiny kernel pair
    scale_kernel<<<ceil_div((long long)M * M, 256), 256, 0, s>>>(bK.as<double>(), (long long)M * M, sigma * sigma);
    GPXB_LAUNCHED(ctx);
    add_vec_kernel<<<ceil_div((long long)M * M, 256), 256, 0, s>>>(C, bK.as<double>(), (long long)M * M);
    GPXB_LAUNCHED(ctx);
    add_diag_kernel<<<ceil_div(M, 256), 256, 0, s>>>(C, M, M, 1e-10);
    GPXB_LAUNCHED(ctx);
  }
  GPXB_TRY(potrf_lower(ctx, C, M, M, bInv.as<double>(), s));
  transpose_out_kernel<<<ceil_div((long long)M * T, 256), 256, 0, s>>>(v, T, M, bVt.as<double>());
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(trsm_right_lt(ctx, C, M, bInv.as<double>(), bVt.as<double>(), M, T, M, s));
  GPXB_TRY(trsm_right_l(ctx, C, M, bInv.as<double>(), bVt.as<double>(), M, T, M, s));
  transpose_out_kernel<<<ceil_div((long long)M * T, 256), 256, 0, s>>>(bVt.as<double>(), M, T, c_out);
  GPXB_LAUNCHED(ctx);
  return 0;
}

// mean[ns x T] = mu + K(x, x_locs) c; optional covariance exactly as the reference builds it
// (from the TEST points: C_mm = sigma^2 K_mm + K_m* K_*m + 1e-10 I, _sgpr.py:452-465).
int sgpr_predict(Ctx* ctx, int kind, const double* x_locs, int M, int D, const double* x, int ns,
                 const double* c, int T, const double* mu, double ell, double sigma, double* mean_out,
                 double* cov_out, cudaStream_t s) {
  GPXB_ARG_CHECK(ns > 0 && D > 0 && T > 0 && M > 0, -1);
  const ZLayout zl = z_layout(D);
  DevBuf bZl, bZs, bKs;
  GPXB_TRY(bZl.alloc(sizeof(double) * (size_t)M * zl.S, s));
  GPXB_TRY(bZs.alloc(sizeof(double) * (size_t)ns * zl.S, s));
  GPXB_TRY(bKs.alloc(sizeof(double) * (size_t)ns * M, s));
  double *Zl = bZl.as<double>(), *Zs = bZs.as<double>(), *Ks = bKs.as<double>();
  GPXB_TRY(prep_z(ctx, x_locs, D, M, zl, ell, Zl, s));
  GPXB_TRY(prep_z(ctx, x, D, ns, zl, ell, Zs, s));
  {
    GramArgs g;  // Ks = K(x, x_locs) [ns x M]
    g.kind = kind; g.Z1 = Zs; g.n1 = ns; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell;
    g.out = Ks; g.ldo = M;
    GPXB_TRY(gram_launch(ctx, g, s));
  }
  {
    GemmArgs g;
    g.M = ns; g.N = T; g.K = M;
    g.A = Ks; g.lda = M; g.a_kmajor = true;
    g.B = c; g.ldb = T; g.b_kmajor = false;
    g.C = mean_out; g.ldc = T;
    GPXB_TRY(gemm_f64(ctx, g, s));
  }
  add_scalar_kernel<<<ceil_div((long long)ns * T, 256), 256, 0, s>>>(mean_out, (long long)ns * T, mu);
  GPXB_LAUNCHED(ctx);
  if (!cov_out) return 0;

  DevBuf bKmm, bC, bInv1, bInv2, bGt, bHt;
  GPXB_TRY(bKmm.alloc(sizeof(double) * (size_t)M * M, s));
  GPXB_TRY(bC.alloc(sizeof(double) * (size_t)M * M, s));
  GPXB_TRY(bInv1.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  GPXB_TRY(bInv2.alloc(sizeof(double) * leaf_inv_doubles(M), s));
  GPXB_TRY(bGt.alloc(sizeof(double) * (size_t)ns * M, s));
  GPXB_TRY(bHt.alloc(sizeof(double) * (size_t)ns * M, s));
  double *Kmm = bKmm.as<double>(), *C = bC.as<double>(), *Gt = bGt.as<double>(), *Ht = bHt.as<double>();
  GramArgs g;
  g.kind = kind; g.Z1 = Zl; g.n1 = M; g.Z2 = Zl; g.n2 = M; g.zl = zl; g.ell = ell;
  g.sym = 1; g.lower_only = 1; g.out = Kmm; g.ldo = M;
  GPXB_CUDA_CHECK(cudaMemsetAsync(Kmm, 0, sizeof(double) * (size_t)M * M, s));
  GPXB_TRY(gram_launch(ctx, g, s));
  // C = sigma^2 K_mm + Ks^T Ks + 1e-10 I
  GPXB_CUDA_CHECK(cudaMemcpyAsync(C, Kmm, sizeof(double) * (size_t)M * M, cudaMemcpyDeviceToDevice, s));
  scale_kernel<<<ceil_div((long long)M * M, 256), 256, 0, s>>>(C, (long long)M * M, sigma * sigma);
  GPXB_LAUNCHED(ctx);
  {
    GemmArgs a;
    a.M = M; a.N = M; a.K = ns;
    a.A = Ks; a.lda = M; a.a_kmajor = false;
    a.B = Ks; a.ldb = M; a.b_kmajor = false;
    a.C = C; a.ldc = M; a.beta = 1.0; a.lower_only = 1;
    GPXB_TRY(gemm_f64(ctx, a, s));
  }
  add_diag_kernel<<<ceil_div(M, 256), 256, 0, s>>>(C, M, M, 1e-10);
  GPXB_LAUNCHED(ctx);
  add_diag_kernel<<<ceil_div(M, 256), 256, 0, s>>>(Kmm, M, M, 1e-10);
  GPXB_LAUNCHED(ctx);
  GPXB_TRY(potrf_lower(ctx, Kmm, M, M, bInv1.as<double>(), s));
  GPXB_TRY(potrf_lower(ctx, C, M, M, bInv2.as<double>(), s));
  GPXB_CUDA_CHECK(cudaMemcpyAsync(Gt, Ks, sizeof(double) * (size_t)ns * M, cudaMemcpyDeviceToDevice, s));
  GPXB_CUDA_CHECK(cudaMemcpyAsync(Ht, Ks, sizeof(double) * (size_t)ns * M, cudaMemcpyDeviceToDevice, s));
  GPXB_TRY(trsm_right_lt(ctx, Kmm, M, bInv1.as<double>(), Gt, M, ns, M, s));
  GPXB_TRY(trsm_right_lt(ctx, C, M, bInv2.as<double>(), Ht, M, ns, M, s));
  GramArgs h;
  h.kind = kind; h.Z1 = Zs; h.n1 = ns; h.Z2 = Zs; h.n2 = ns; h.zl = zl; h.ell = ell; h.sym = 1;
  h.out = cov_out; h.ldo = ns;
  GPXB_TRY(gram_launch(ctx, h, s));
  GemmArgs m1;  // cov -= Gt Gt^T
  m1.M = ns; m1.N = ns; m1.K = M;
  m1.A = Gt; m1.lda = M; m1.a_kmajor = true;
  m1.B = Gt; m1.ldb = M; m1.b_kmajor = true;
  m1.C = cov_out; m1.ldc = ns; m1.alpha = -1.0; m1.beta = 1.0;
  GPXB_TRY(gemm_f64(ctx, m1, s));
  GemmArgs m2 = m1;  // cov += sigma^2 Ht Ht^T
  m2.A = Ht; m2.B = Ht; m2.alpha = sigma * sigma;
  GPXB_TRY(gemm_f64(ctx, m2, s));
  return 0;
}

}  // namespace gpxb
