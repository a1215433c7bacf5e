// C-ABI surface of Family 3 (matrix-free mat-vec, Lanczos, PCG, RPCholesky), SGPR and the
// multi-GPU plumbing.  Arrays are passed whole (replicated) on every rank; the row range a
// rank works on is derived from (rank, world) inside, so the calls look the same on 1 or 8 GPUs.
#include "../../include/gpxb200.h"
#include "chol.cuh"
#include "comm.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "matvec.cuh"
#include "rpchol.cuh"
#include "small_kernels.cuh"

namespace gpxb {
int lanczos_block(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N,
                  ZLayout zl, double diag_add, const double* U, int P, int PS, int m, double* alpha_out,
                  double* beta_out, int* flag, cudaStream_t s);
int rademacher_probes(Ctx* ctx, const uint32_t key[2], long long Ntot, int P, long long row0, int n_rows,
                      int PS, int partitionable, double* U, cudaStream_t s);
int precond_build(Ctx* ctx, const double* F, long long ldf, int k, int n_loc, double sigma, double* Pkk,
                  cudaStream_t s);
int pcg_solve(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N,
              ZLayout zl, double diag_add, const double* b, const double* F, long long ldf, int k,
              const double* Pkk, double sigma, double tol, double atol, int maxiter, double* x, int* iters_out,
              cudaStream_t s);
int lml_iter_fused(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N, ZLayout zl,
                   double diag_add, const double* b, const double* F, long long ldf, int k, const double* Pkk,
                   double sigma, double tol, double atol, int maxiter, double* x, int* iters_out, const double* U,
                   int P, int PS, int m, double* alpha_out, double* beta_out, int* flag, cudaStream_t s);
int lanczos_block_grad(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N,
                       ZLayout zl, double ell, double sigma, const double* U, int P, int PS, int m,
                       double* alpha_out, double* beta_out, double* dalpha_out, double* dbeta_out, int Ptot,
                       int* flag, cudaStream_t s);
int sgpr_lml(Ctx* ctx, int kind, const double* x, const double* y, int n_loc, long long N, int D, int T,
             const double* x_locs, int M, double ell, double sigma, const double* mu, double* out,
             cudaStream_t s);
int sgpr_fit(Ctx* ctx, int kind, const double* x, const double* y, int n_loc, int D, int T,
             const double* x_locs, int M, double ell, double sigma, const double* mu, double* c_out,
             cudaStream_t s);
int sgpr_lml_grad(Ctx* ctx, int kind, const double* x, const double* y, int n_loc, long long N, int D,
                  const double* x_locs, int M, double ell, double sigma, const double* mu, int want_grad,
                  double* out3, double* grad_locs, cudaStream_t s);
int sgpr_predict(Ctx* ctx, int kind, const double* x_locs, int M, int D, const double* x, int ns,
                 const double* c, int T, const double* mu, double ell, double sigma, double* mean_out,
                 double* cov_out, cudaStream_t s);
}  // namespace gpxb

using namespace gpxb;

namespace {
struct Shard {
  long long b, e;
  int n() const { return (int)(e - b); }
};
Shard shard_of(Ctx* c, long long N) {
  return Shard{shard_begin(N, c->world, c->rank), shard_begin(N, c->world, c->rank + 1)};
}
}  // namespace

extern "C" {

int gpxb_comm_unique_id(void* out128) { return comm_unique_id(out128); }

int gpxb_comm_init(gpxb_ctx* ctx, const void* id128, int rank, int world) {
  GPXB_ARG_CHECK(ctx != nullptr, -1);
  return comm_init(reinterpret_cast<Ctx*>(ctx), id128, rank, world);
}

int gpxb_shard_range(long long N, int world, int rank, long long* begin, long long* end) {
  GPXB_ARG_CHECK(begin && end && N >= 0 && world >= 1 && rank >= 0 && rank < world, -1);
  *begin = shard_begin(N, world, rank);
  *end = shard_begin(N, world, rank + 1);
  return 0;
}

int gpxb_gather_rows(gpxb_ctx* ctx_, const double* x, int D, const long long* idx, int M, double* out,
                     gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && idx && out && D > 0 && M >= 0, -1);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  if (M == 0) return 0;
  sk::gather_rows_kernel<<<ceil_div((long long)M * D, 256), 256, 0, s>>>(x, D, idx, M, out);
  GPXB_LAUNCHED(ctx);
  return 0;
}

int gpxb_kmatvec(gpxb_ctx* ctx_, int kind, const double* x_rows, int n_rows, long long row_offset,
                 const double* x_all, int N, int D, double ell, double diag_add, const double* V,
                 long long ldv, int P, double* W, long long ldw, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x_rows && x_all && V && W, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3, -2);
  return kmatvec(reinterpret_cast<Ctx*>(ctx_), kind, x_rows, n_rows, row_offset, x_all, N, D, ell, diag_add,
                 V, ldv, P, W, ldw, (cudaStream_t)stream);
}

int gpxb_rpcholesky(gpxb_ctx* ctx_, int kind, const double* x, int N, int D, double ell, uint32_t key0,
                    uint32_t key1, int M, int partitionable, double* F_out, long long* pivots_out,
                    gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && pivots_out && N > 0 && D > 0 && M > 0, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const ZLayout zl = z_layout(D);
  const Shard sh = shard_of(ctx, N);  // rows are sharded over the ranks (1024-row aligned)
  DevBuf z, ft;
  GPXB_TRY(z.alloc(sizeof(double) * (size_t)std::max(sh.n(), 1) * zl.S, s));
  GPXB_TRY(prep_z(ctx, x + sh.b * D, D, sh.n(), zl, ell, z.as<double>(), s));
  GPXB_TRY(ft.alloc(sizeof(double) * std::max<size_t>(rp_factor_doubles(sh.n(), M), 8), s));
  const uint32_t key[2] = {key0, key1};
  GPXB_TRY(rpcholesky_z(ctx, kind, z.as<double>(), sh.n(), sh.b, N, zl, key, M, partitionable, ft.as<double>(),
                        pivots_out, s));
  if (F_out) {
    GPXB_TRY(rp_unblock_rowmajor(ctx, ft.as<double>(), sh.n(), M, F_out + sh.b * M, s));
    if (ctx->world > 1) {
      ctx->shard_N = N;
      GPXB_TRY(comm_allgather_rows(ctx, F_out + sh.b * M, F_out, sh.n(), M, s));
    }
  }
  return 0;
}

int gpxb_rademacher_probes(gpxb_ctx* ctx_, uint32_t key0, uint32_t key1, int N, int P, int partitionable,
                           double* U_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && U_out && N > 0 && P > 0, -1);
  const uint32_t key[2] = {key0, key1};
  return rademacher_probes(reinterpret_cast<Ctx*>(ctx_), key, N, P, 0, N, P, partitionable, U_out,
                           (cudaStream_t)stream);
}

int gpxb_lanczos_set_restart(gpxb_ctx* ctx_, int enabled, uint32_t key0, uint32_t key1, int partitionable) {
  GPXB_ARG_CHECK(ctx_, -1);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  ctx->restart_on = enabled != 0;
  ctx->restart_key[0] = key0;
  ctx->restart_key[1] = key1;
  ctx->restart_partitionable = partitionable != 0;
  return 0;
}

int gpxb_lanczos(gpxb_ctx* ctx_, int kind, const double* x, int N, int D, double ell, double diag_add,
                 const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                 int* breakdown_flag, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && U && alpha_out && beta_out && breakdown_flag, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0 && D > 0 && P > 0 && m > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const ZLayout zl = z_layout(D);
  const Shard sh = shard_of(ctx, N);
  DevBuf za, up;
  GPXB_TRY(za.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, za.as<double>(), s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(breakdown_flag, 0, sizeof(int), s));
  for (int c0 = 0; c0 < P; c0 += 32) {  // blocks of 32 probes
    const int pb = std::min(32, P - c0);
    const int PS = v_padded_stride(pb);
    GPXB_TRY(up.alloc(sizeof(double) * ((size_t)std::max(sh.n(), 1) * PS + 16), s));
    GPXB_TRY(pack_v(ctx, U + sh.b * ldu, ldu, sh.n(), c0, pb, PS, up.as<double>(), s));
    GPXB_TRY(lanczos_block(ctx, kind, za.as<double>() + sh.b * zl.S, sh.n(), sh.b, za.as<double>(), N, zl,
                           diag_add, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                           beta_out + (size_t)c0 * m, breakdown_flag, s));
    up.release();
  }
  return 0;
}

int gpxb_lanczos_grad(gpxb_ctx* ctx_, int kind, const double* x, int N, int D, double ell, double sigma,
                      const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                      double* dalpha_out, double* dbeta_out, int* breakdown_flag, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && U && alpha_out && beta_out && dalpha_out && dbeta_out && breakdown_flag, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0 && D > 0 && P > 0 && m > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const ZLayout zl = z_layout(D);
  const Shard sh = shard_of(ctx, N);
  DevBuf za, up;
  GPXB_TRY(za.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, za.as<double>(), s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(breakdown_flag, 0, sizeof(int), s));
  for (int c0 = 0; c0 < P; c0 += 32) {
    const int pb = std::min(32, P - c0);
    const int PS = v_padded_stride(pb);
    GPXB_TRY(up.alloc(sizeof(double) * ((size_t)std::max(sh.n(), 1) * PS + 16), s));
    GPXB_TRY(pack_v(ctx, U + sh.b * ldu, ldu, sh.n(), c0, pb, PS, up.as<double>(), s));
    GPXB_TRY(lanczos_block_grad(ctx, kind, za.as<double>() + sh.b * zl.S, sh.n(), sh.b, za.as<double>(), N, zl, ell,
                                sigma, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                                beta_out + (size_t)c0 * m, dalpha_out + (size_t)c0 * m, dbeta_out + (size_t)c0 * m, P,
                                breakdown_flag, s));
    up.release();
  }
  return 0;
}

int gpxb_kmatvec_dl(gpxb_ctx* ctx_, int kind, const double* x, int N, int D, double ell, const double* V,
                    long long ldv, int P, double* W, long long ldw, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && V && W, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0 && D > 0 && P > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const ZLayout zl = z_layout(D);
  DevBuf za, vp;
  GPXB_TRY(za.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, za.as<double>(), s));
  for (int c0 = 0; c0 < P; c0 += 32) {
    const int pb = std::min(32, P - c0);
    const int PS = v_padded_stride(pb);
    GPXB_TRY(vp.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
    GPXB_TRY(pack_v(ctx, V, ldv, N, c0, pb, PS, vp.as<double>(), s));
    KmvArgs a;
    a.kind = kind; a.Zr = za.as<double>(); a.n_rows = N; a.row_offset = 0; a.Za = za.as<double>(); a.N = N; a.zl = zl;
    a.V = vp.as<double>(); a.PS = PS; a.P = pb; a.W = W + c0; a.ldw = ldw; a.dl_ell = ell;
    GPXB_TRY(kmatvec_launch(ctx, a, s));
    vp.release();
  }
  return 0;
}

int gpxb_gpr_fit_iter(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D, double ell,
                      double sigma, int mean_mode, int n_pivots, uint32_t key0, uint32_t key1,
                      int partitionable, double tol, double atol, int maxiter, double* c_out, double* mu_out,
                      int* iters_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && c_out && mu_out, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0 && D > 0 && n_pivots >= 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const ZLayout zl = z_layout(D);
  const Shard sh = shard_of(ctx, N);
  const double s2 = sigma * sigma + 1e-10;
  DevBuf za, r, ft, piv, pkk, xl;
  GPXB_TRY(za.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, za.as<double>(), s));
  GPXB_TRY(r.alloc(sizeof(double) * N, s));
  sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu_out);
  GPXB_LAUNCHED(ctx);
  sk::center_kernel<<<ceil_div(N, 256), 256, 0, s>>>(y, N, mu_out, r.as<double>());
  GPXB_LAUNCHED(ctx);
  const long long ldf = round_up(N, 2);
  if (n_pivots > 0) {
    // preconditioner: row-sharded RPCholesky, P_kk = inv(sigma^2 I + F^T F) (all-reduced)
    DevBuf fb;
    GPXB_TRY(fb.alloc(sizeof(double) * std::max<size_t>(rp_factor_doubles(sh.n(), n_pivots), 8), s));
    GPXB_TRY(ft.alloc(sizeof(double) * (size_t)n_pivots * ldf, s));
    GPXB_TRY(piv.alloc(sizeof(long long) * n_pivots, s));
    GPXB_TRY(pkk.alloc(sizeof(double) * (size_t)n_pivots * n_pivots, s));
    const uint32_t key[2] = {key0, key1};
    GPXB_TRY(rpcholesky_z(ctx, kind, za.as<double>() + sh.b * zl.S, sh.n(), sh.b, N, zl, key, n_pivots,
                          partitionable, fb.as<double>(), piv.as<long long>(), s));
    // local rows of F, transposed: ft[c][0 .. n_loc)
    GPXB_TRY(rp_unblock_transposed(ctx, fb.as<double>(), sh.n(), n_pivots, ft.as<double>(), ldf, s));
    GPXB_TRY(precond_build(ctx, ft.as<double>(), ldf, n_pivots, sh.n(), sigma, pkk.as<double>(), s));
  }
  GPXB_TRY(xl.alloc(sizeof(double) * std::max(sh.n(), 1), s));
  if (maxiter <= 0) maxiter = 10 * N;
  GPXB_TRY(pcg_solve(ctx, kind, za.as<double>() + sh.b * zl.S, sh.n(), sh.b, za.as<double>(), N, zl, s2,
                     r.as<double>() + sh.b, n_pivots > 0 ? ft.as<double>() : nullptr, ldf, n_pivots,
                     pkk.as<double>(), sigma, tol, atol, maxiter, xl.as<double>(), iters_out, s));
  if (ctx->world > 1) {
    ctx->shard_N = N;
    GPXB_TRY(comm_allgather_rows(ctx, xl.as<double>(), c_out, sh.n(), 1, s));
  } else {
    GPXB_CUDA_CHECK(cudaMemcpyAsync(c_out, xl.as<double>(), sizeof(double) * N, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

int gpxb_gpr_lml_iter(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D, double ell,
                      double sigma, int mean_mode, int n_pivots, uint32_t key0, uint32_t key1, int partitionable,
                      double tol, double atol, int maxiter, const double* U, long long ldu, int P, int m,
                      double* c_out, double* mu_out, int* iters_out, double* alpha_out, double* beta_out,
                      int* breakdown_flag, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && c_out && mu_out && U && alpha_out && beta_out && breakdown_flag, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0 && D > 0 && n_pivots >= 0 && P > 0 && m > 0, -2);
  if (P > 31) {  // the PCG vector rides as column P of ONE 32-column block: larger probe sets run the two parts in turn
    GPXB_TRY(gpxb_gpr_fit_iter(ctx_, kind, x, y, N, D, ell, sigma, mean_mode, n_pivots, key0, key1, partitionable, tol,
                               atol, maxiter, c_out, mu_out, iters_out, stream));
    return gpxb_lanczos(ctx_, kind, x, N, D, ell, sigma * sigma + 1e-10, U, ldu, P, m, alpha_out, beta_out,
                        breakdown_flag, stream);
  }
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const ZLayout zl = z_layout(D);
  const Shard sh = shard_of(ctx, N);
  const double s2 = sigma * sigma + 1e-10;
  DevBuf za, r, ft, piv, pkk, xl, up;
  GPXB_TRY(za.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(prep_z(ctx, x, D, N, zl, ell, za.as<double>(), s));
  GPXB_TRY(r.alloc(sizeof(double) * N, s));
  sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu_out);
  GPXB_LAUNCHED(ctx);
  sk::center_kernel<<<ceil_div(N, 256), 256, 0, s>>>(y, N, mu_out, r.as<double>());
  GPXB_LAUNCHED(ctx);
  const long long ldf = round_up(N, 2);
  if (n_pivots > 0) {
    DevBuf fb;
    GPXB_TRY(fb.alloc(sizeof(double) * std::max<size_t>(rp_factor_doubles(sh.n(), n_pivots), 8), s));
    GPXB_TRY(ft.alloc(sizeof(double) * (size_t)n_pivots * ldf, s));
    GPXB_TRY(piv.alloc(sizeof(long long) * n_pivots, s));
    GPXB_TRY(pkk.alloc(sizeof(double) * (size_t)n_pivots * n_pivots, s));
    const uint32_t key[2] = {key0, key1};
    GPXB_TRY(rpcholesky_z(ctx, kind, za.as<double>() + sh.b * zl.S, sh.n(), sh.b, N, zl, key, n_pivots,
                          partitionable, fb.as<double>(), piv.as<long long>(), s));
    GPXB_TRY(rp_unblock_transposed(ctx, fb.as<double>(), sh.n(), n_pivots, ft.as<double>(), ldf, s));
    GPXB_TRY(precond_build(ctx, ft.as<double>(), ldf, n_pivots, sh.n(), sigma, pkk.as<double>(), s));
  }
  GPXB_TRY(xl.alloc(sizeof(double) * std::max(sh.n(), 1), s));
  if (maxiter <= 0) maxiter = 10 * N;
  GPXB_CUDA_CHECK(cudaMemsetAsync(breakdown_flag, 0, sizeof(int), s));
  const int PS = v_padded_stride(P + 1);
  GPXB_TRY(up.alloc(sizeof(double) * ((size_t)std::max(sh.n(), 1) * PS + 16), s));
  GPXB_TRY(pack_v(ctx, U + sh.b * ldu, ldu, sh.n(), 0, P, PS, up.as<double>(), s));
  GPXB_TRY(lml_iter_fused(ctx, kind, za.as<double>() + sh.b * zl.S, sh.n(), sh.b, za.as<double>(), N, zl, s2,
                          r.as<double>() + sh.b, n_pivots > 0 ? ft.as<double>() : nullptr, ldf, n_pivots,
                          pkk.as<double>(), sigma, tol, atol, maxiter, xl.as<double>(), iters_out, up.as<double>(), P,
                          PS, m, alpha_out, beta_out, breakdown_flag, s));
  if (ctx->world > 1) {
    ctx->shard_N = N;
    GPXB_TRY(comm_allgather_rows(ctx, xl.as<double>(), c_out, sh.n(), 1, s));
  } else {
    GPXB_CUDA_CHECK(cudaMemcpyAsync(c_out, xl.as<double>(), sizeof(double) * N, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

int gpxb_sgpr_lml(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D, int T,
                  const double* x_locs, int M, double ell, double sigma, int mean_mode, double* out,
                  gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && x_locs && out, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const Shard sh = shard_of(ctx, N);
  DevBuf mu;
  GPXB_TRY(mu.alloc(sizeof(double), s));
  sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N * T, mean_mode == GPXB_MEAN_DATA, mu.as<double>());
  GPXB_LAUNCHED(ctx);
  return sgpr_lml(ctx, kind, x + sh.b * D, y + sh.b * T, sh.n(), N, D, T, x_locs, M, ell, sigma,
                  mu.as<double>(), out, s);
}

int gpxb_sgpr_lml_grad(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D,
                       const double* x_locs, int M, double ell, double sigma, int mean_mode, int want_grad,
                       double* out3, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && x_locs && out3, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const Shard sh = shard_of(ctx, N);
  DevBuf mu;
  GPXB_TRY(mu.alloc(sizeof(double), s));
  sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu.as<double>());
  GPXB_LAUNCHED(ctx);
  return sgpr_lml_grad(ctx, kind, x + sh.b * D, y + sh.b, sh.n(), N, D, x_locs, M, ell, sigma, mu.as<double>(),
                       want_grad, out3, nullptr, s);
}

int gpxb_sgpr_lml_grad_locs(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D,
                            const double* x_locs, int M, double ell, double sigma, int mean_mode, double* out3,
                            double* grad_locs, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && x_locs && out3 && grad_locs, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const Shard sh = shard_of(ctx, N);
  DevBuf mu;
  GPXB_TRY(mu.alloc(sizeof(double), s));
  sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu.as<double>());
  GPXB_LAUNCHED(ctx);
  return sgpr_lml_grad(ctx, kind, x + sh.b * D, y + sh.b, sh.n(), N, D, x_locs, M, ell, sigma, mu.as<double>(), 1,
                       out3, grad_locs, s);
}

int gpxb_sgpr_fit(gpxb_ctx* ctx_, int kind, const double* x, const double* y, int N, int D, int T,
                  const double* x_locs, int M, double ell, double sigma, int mean_mode, double* c_out,
                  double* mu_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x && y && x_locs && c_out && mu_out, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && N > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const Shard sh = shard_of(ctx, N);
  sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N * T, mean_mode == GPXB_MEAN_DATA, mu_out);
  GPXB_LAUNCHED(ctx);
  return sgpr_fit(ctx, kind, x + sh.b * D, y + sh.b * T, sh.n(), D, T, x_locs, M, ell, sigma, mu_out, c_out, s);
}

int gpxb_sgpr_predict(gpxb_ctx* ctx_, int kind, const double* x_locs, int M, int D, const double* x, int ns,
                      const double* c, int T, const double* mu, double ell, double sigma, double* mean_out,
                      double* cov_out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x_locs && x && c && mu && mean_out, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3, -2);
  return sgpr_predict(reinterpret_cast<Ctx*>(ctx_), kind, x_locs, M, D, x, ns, c, T, mu, ell, sigma, mean_out,
                      cov_out, (cudaStream_t)stream);
}

}  // extern "C"
