// Family 3b: block Lanczos tridiagonalisation (for stochastic Lanczos quadrature) and
// preconditioned conjugate gradients over the matrix-free mat-vec (matvec.cu).
//
// Replaces gpx/lanczos.py:29-183 (lanczos_tridiagonal with full modified-Gram-Schmidt
// re-orthogonalisation, lanczos_logdet's Rademacher probes), jax.scipy.sparse.linalg.cg as
// called by gpx/models/_gpr.py:285-310 and the RPCholesky preconditioner of
// gpx/models/_gpr.py:180-216.  The reference runs the probes one after the other; here all
// P probes advance together as the columns of an N x P block, so one pass over the
// recomputed kernel tiles serves every probe.
#include <cmath>
#include <functional>

#include "../../include/gpxb200.h"
#include "chol.cuh"
#include "comm.cuh"
#include "gemm.cuh"
#include "matvec.cuh"
#include "mma.cuh"
#include "small_kernels.cuh"
#include <vector>

#include "threefry.cuh"

namespace gpxb {

namespace {

constexpr int RB = 256;   // rows per CTA in the column-wise reductions
constexpr int MAXP = 32;  // columns per block

// part[blockIdx.x][p] = sum over this CTA's rows of X[i][p] * Y[i][p]   (fixed order)
__global__ void __launch_bounds__(256) coldot_partial_kernel(const double* __restrict__ X,
                                                             const double* __restrict__ Y, int n, int PS,
                                                             int P, double* __restrict__ part) {
  __shared__ double sh[8][MAXP];
  const int p = threadIdx.x & 31, g = threadIdx.x >> 5;  // 8 row groups x 32 columns
  const long long r0 = (long long)blockIdx.x * RB;
  double s = 0.0;
  if (p < P) {
    for (int i = g; i < RB; i += 8) {
      const long long r = r0 + i;
      if (r < n) s += X[r * PS + p] * Y[r * PS + p];
    }
  }
  sh[g][p] = s;
  __syncthreads();
  if (g == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][p];
    part[(long long)blockIdx.x * MAXP + p] = t;
  }
}

// out[p] = sum_b part[b][p] in block order (one CTA, 32 columns x 32 lanes of blocks)
__global__ void __launch_bounds__(1024) coldot_final_kernel(const double* __restrict__ part, int nblk,
                                                            double* __restrict__ out) {
  __shared__ double sh[32][MAXP + 1];
  const int p = threadIdx.x & 31, g = threadIdx.x >> 5;
  double s = 0.0;
  for (int b = g; b < nblk; b += 32) s += part[(long long)b * MAXP + p];
  sh[g][p] = s;
  __syncthreads();
  if (g == 0) {
    double t = 0.0;
    for (int k = 0; k < 32; ++k) t += sh[k][p];
    out[p] = t;
  }
}

// W[:,p] -= a[p] * X[:,p] (+ b[p] * Y[:,p] if Y)
__global__ void colaxpy_kernel(double* __restrict__ W, const double* __restrict__ a,
                               const double* __restrict__ X, const double* __restrict__ b,
                               const double* __restrict__ Y, int n, int PS, int P) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * PS) return;
  const int p = (int)(idx % PS);
  if (p >= P) return;
  double w = W[idx] - a[p] * X[idx];
  if (Y) w -= b[p] * Y[idx];
  W[idx] = w;
}

// beta[p] = sqrt(nrm2[p]); Vn[:,p] = W[:,p] / beta[p]; flags breakdown (beta <= 1e-12)
__global__ void colnormalize_kernel(const double* __restrict__ W, const double* __restrict__ nrm2,
                                    double* __restrict__ Vn, int n, int PS, int P, int* __restrict__ flag) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * PS) return;
  const int p = (int)(idx % PS);
  if (p >= P) {
    Vn[idx] = 0.0;
    return;
  }
  const double beta = sqrt(nrm2[p]);
  if (!(beta > 1e-12) && idx < PS) atomicOr(flag, 1);
  Vn[idx] = W[idx] / beta;
}

// T bookkeeping: alpha_out[p][j] = alpha[p];  beta_out[p][j] = sqrt(nrm2[p])
__global__ void store_ab_kernel(const double* __restrict__ alpha, const double* __restrict__ nrm2, int P,
                                int m, int j, double* __restrict__ alpha_out, double* __restrict__ beta_out,
                                double* __restrict__ beta_cur) {
  const int p = threadIdx.x;
  if (p >= P) return;
  alpha_out[p * m + j] = alpha[p];
  const double b = sqrt(nrm2[p]);
  beta_out[p * m + j] = b;
  beta_cur[p] = b;
}

// Rademacher probes normalised to unit length: jax.random.rademacher(subkey, (N, P)) / sqrt(N)
// (gpx/lanczos.py:168-170).  row0: global index of the first local row.
__global__ void rademacher_kernel(uint32_t k0, uint32_t k1, long long Ntot, int P, long long row0, int n_rows,
                                  int PS, int partitionable, double* __restrict__ U) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_rows * PS) return;
  const long long i = idx / PS;
  const int p = (int)(idx % PS);
  if (p >= P) {
    U[idx] = 0.0;
    return;
  }
  const unsigned long long e = (unsigned long long)(row0 + i) * P + p;
  const double u = jax_uniform_f64(k0, k1, e, (unsigned long long)Ntot * P, partitionable);
  const double sgn = (u < 0.5) ? 1.0 : -1.0;
  U[idx] = sgn / sqrt((double)Ntot);
}

int coldot(Ctx* ctx, const double* X, const double* Y, int n, int PS, int P, double* part, double* out,
           cudaStream_t s) {
  const int nblk = ceil_div(n, RB);
  {
    PhaseScope ph(ctx, PH_DOTS, s);
    coldot_partial_kernel<<<nblk, 256, 0, s>>>(X, Y, n, PS, P, part);
    GPXB_LAUNCH_CHECK();
    coldot_final_kernel<<<1, 1024, 0, s>>>(part, nblk, out);
    GPXB_LAUNCH_CHECK();
    ctx->launches += 2;
  }
  return comm_allreduce_sum(ctx, out, MAXP, s);  // no-op on one GPU
}

}  // namespace

// Block Lanczos on A = K(x,x) + diag_add*I restricted to this rank's rows.
//   Zr: this rank's pre-scaled rows (n_loc x S), Za: all rows (N x S), row0: first local row.
//   U: normalised start block, local rows, padded stride PS.
//   alpha_out / beta_out: [P][m] device.  T_p = tridiag(alpha_out[p], beta_out[p][0..m-2]).
//   flag: device int set to 1 on breakdown (beta <= 1e-12).
// apply(v, w): w = A v for this rank's rows; v, w are n_loc x PS blocks with P live columns.
// A "rider" shares the Lanczos block mat-vecs: while it is active, its vector travels as column P of the block
// (put), the operator is applied to P + 1 columns in the same pass over the recomputed kernel tiles, and the
// rider's product is handed back (take).  Used to run the first m PCG iterations of the iterative LML inside the
// SLQ mat-vecs (a 31st column costs nothing: the tensor-core tile is 32 columns wide).
struct LanczosRider {
  std::function<bool()> active;
  std::function<int(double* block, int col, int PS)> put;
  std::function<int(const double* W, int col, int PS)> take;
};

int lanczos_core_rider(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                       const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                       LanczosRider* rider, cudaStream_t s);

// Breakdown branch of lanczos_tridiagonal (gpx/lanczos.py:92-108): column p of the next Lanczos block becomes
// orthonormalize(random.normal(subkey, (N,)), vecs) for every probe p in `bad` (defined after the vector helpers).
int lanczos_restart_columns(Ctx* ctx, cudaStream_t s, double* vecs, size_t blk, int n_loc, long long N, int PS, int j,
                            const std::vector<int>& bad, uint32_t sk0, uint32_t sk1, int partitionable);

int lanczos_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                 const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                 cudaStream_t s) {
  return lanczos_core_rider(ctx, n_loc, N, apply, U, P, PS, m, alpha_out, beta_out, flag, nullptr, s);
}

namespace {
int lanczos_pass(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                 const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                 LanczosRider* rider, bool restarts, cudaStream_t s);
}  // namespace

// Fast pass first: enqueue-only, a breakdown (beta <= 1e-12) only raises the device flag.  With the restart key set
// on the context (gpxb_lanczos_set_restart) the flag is read back once at the end; in the rare case that it is set
// the recursion is redone with the reference's breakdown branch, which needs one host read of beta per step.
int lanczos_core_rider(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                       const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                       LanczosRider* rider, cudaStream_t s) {
  GPXB_TRY(lanczos_pass(ctx, n_loc, N, apply, U, P, PS, m, alpha_out, beta_out, flag, rider, false, s));
  if (!ctx->restart_on || m < 2) return 0;
  int h = 0;
  GPXB_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  GPXB_CUDA_CHECK(cudaStreamSynchronize(s));
  if (h == 0) return 0;
  GPXB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
  return lanczos_pass(ctx, n_loc, N, apply, U, P, PS, m, alpha_out, beta_out, flag, nullptr, true, s);
}

namespace {
int lanczos_pass(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                 const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                 LanczosRider* rider, bool restarts, cudaStream_t s) {
  GPXB_ARG_CHECK(P >= 1 && P <= MAXP && m >= 1, -30);
  uint32_t rk0 = ctx->restart_key[0], rk1 = ctx->restart_key[1];  // the key lanczos_tridiagonal carries
  ctx->shard_N = N;
  const size_t blk = (size_t)n_loc * PS;
  DevBuf bVecs, bW, bPart, bSc;
  GPXB_TRY(bVecs.alloc(sizeof(double) * blk * m, s));
  GPXB_TRY(bW.alloc(sizeof(double) * blk, s));
  GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)ceil_div(n_loc, RB) * MAXP, s));
  GPXB_TRY(bSc.alloc(sizeof(double) * MAXP * 4, s));
  double *vecs = bVecs.as<double>(), *W = bW.as<double>(), *part = bPart.as<double>();
  double *alpha = bSc.as<double>(), *nrm2 = alpha + MAXP, *coef = alpha + 2 * MAXP, *beta = alpha + 3 * MAXP;
  GPXB_CUDA_CHECK(cudaMemcpyAsync(vecs, U, sizeof(double) * blk, cudaMemcpyDeviceToDevice, s));
  const int nthr = 256;
  const int nblk_e = ceil_div((long long)blk, nthr);

  for (int j = 0; j < m; ++j) {
    double* vj = vecs + blk * j;
    const bool ride = rider && rider->active();
    if (ride) GPXB_TRY(rider->put(vj, P, PS));
    GPXB_TRY(apply(vj, W));                                      // w = A v_j
    if (ride) GPXB_TRY(rider->take(W, P, PS));
    GPXB_TRY(coldot(ctx, W, vj, n_loc, PS, P, part, alpha, s));  // alpha = w . v_j
    {
      PhaseScope ph(ctx, PH_VECOPS, s);
      colaxpy_kernel<<<nblk_e, nthr, 0, s>>>(W, alpha, vj, beta, j > 0 ? vecs + blk * (j - 1) : nullptr, n_loc, PS,
                                             P);                  // w -= alpha v_j + beta v_{j-1}
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
    }
    if (j > 0) {
      // full re-orthogonalisation, modified Gram-Schmidt in column order (lanczos.py:29-37);
      // columns beyond j are still zero in the reference and are skipped
      for (int c = 0; c <= j; ++c) {
        const double* vc = vecs + blk * c;
        GPXB_TRY(coldot(ctx, W, vc, n_loc, PS, P, part, coef, s));
        PhaseScope ph(ctx, PH_VECOPS, s);
        colaxpy_kernel<<<nblk_e, nthr, 0, s>>>(W, coef, vc, nullptr, nullptr, n_loc, PS, P);
        GPXB_LAUNCH_CHECK();
        ctx->launches++;
      }
    }
    GPXB_TRY(coldot(ctx, W, W, n_loc, PS, P, part, nrm2, s));
    PhaseScope ph(ctx, PH_VECOPS, s);
    store_ab_kernel<<<1, MAXP, 0, s>>>(alpha, nrm2, P, m, j, alpha_out, beta_out, beta);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
    if (j + 1 < m) {
      colnormalize_kernel<<<nblk_e, nthr, 0, s>>>(W, nrm2, vecs + blk * (j + 1), n_loc, PS, P, flag);
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
      if (restarts) {
        // subkey, key = random.split(key) at every step, used or not (lanczos.py:99)
        uint32_t sk0, sk1, nk0, nk1;
        jax_split2(rk0, rk1, 0, ctx->restart_partitionable, sk0, sk1);
        jax_split2(rk0, rk1, 1, ctx->restart_partitionable, nk0, nk1);
        rk0 = nk0; rk1 = nk1;
        double hb[MAXP];
        GPXB_CUDA_CHECK(cudaMemcpyAsync(hb, beta, sizeof(double) * P, cudaMemcpyDeviceToHost, s));
        GPXB_CUDA_CHECK(cudaStreamSynchronize(s));
        std::vector<int> bad;
        for (int p = 0; p < P; ++p)
          if (!(hb[p] > 1e-12)) bad.push_back(p);
        if (!bad.empty())
          GPXB_TRY(lanczos_restart_columns(ctx, s, vecs, blk, n_loc, N, PS, j, bad, sk0, sk1,
                                           ctx->restart_partitionable));
      }
    }
  }
  if (restarts) GPXB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));  // every breakdown was handled
  return 0;
}
}  // namespace

int lanczos_block(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N,
                  ZLayout zl, double diag_add, const double* U, int P, int PS, int m, double* alpha_out,
                  double* beta_out, int* flag, cudaStream_t s) {
  // the mat-vec needs the full block V on every rank: all-gather when sharded
  DevBuf bVall;
  double* Vall = nullptr;
  if (ctx->world > 1) {
    GPXB_TRY(bVall.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
    Vall = bVall.as<double>();
  }
  auto apply = [&](const double* v, double* w) -> int {
    const double* vfull = v;
    if (ctx->world > 1) {
      GPXB_TRY(comm_allgather_rows(ctx, v, Vall, n_loc, PS, s));
      vfull = Vall;
    }
    KmvArgs a;
    a.kind = kind; a.Zr = Zr; a.n_rows = n_loc; a.row_offset = row0; a.Za = Za; a.N = N; a.zl = zl;
    a.V = vfull; a.PS = PS; a.P = P; a.W = w; a.ldw = PS; a.diag_add = diag_add;
    return kmatvec_launch(ctx, a, s);
  };
  return lanczos_core(ctx, n_loc, N, apply, U, P, PS, m, alpha_out, beta_out, flag, s);
}

namespace {
// out[p] = a[p] + b[p]
__global__ void padd_kernel(const double* __restrict__ a, const double* __restrict__ b, int P, double* __restrict__ out) {
  const int p = threadIdx.x;
  if (p < P) out[p] = a[p] + b[p];
}
// Y = c * X (all PS columns)
__global__ void colscale_copy_kernel(const double* __restrict__ X, double c, long long n, double* __restrict__ Y) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) Y[idx] = c * X[idx];
}
__global__ void vadd_kernel(double* __restrict__ Y, const double* __restrict__ X, long long n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) Y[idx] += X[idx];
}
// d beta[p] = (w . dw)[p] / beta[p];  d alpha_out[p][j] = d alpha[p];  d beta_out[p][j] = d beta[p]
__global__ void store_dab_kernel(const double* __restrict__ dalpha, const double* __restrict__ wdw,
                                 const double* __restrict__ beta, int P, int m, int j, double* __restrict__ dalpha_out,
                                 double* __restrict__ dbeta_out, double* __restrict__ dbeta_cur) {
  const int p = threadIdx.x;
  if (p >= P) return;
  dalpha_out[p * m + j] = dalpha[p];
  const double db = wdw[p] / beta[p];
  dbeta_out[p * m + j] = db;
  dbeta_cur[p] = db;
}
// dVn[:,p] = (dW[:,p] - Vn[:,p] dbeta[p]) / beta[p]
__global__ void coltangent_kernel(const double* __restrict__ dW, const double* __restrict__ Vn,
                                  const double* __restrict__ dbeta, const double* __restrict__ beta,
                                  double* __restrict__ dVn, int n, int PS, int P) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * PS) return;
  const int p = (int)(idx % PS);
  dVn[idx] = p < P ? (dW[idx] - Vn[idx] * dbeta[p]) / beta[p] : 0.0;
}
}  // namespace

// Block Lanczos together with its forward-mode derivative w.r.t. (lengthscale, sigma) of an operator
// A(ell) + (sigma^2 + 1e-10) I given by matvec(v, out, dl) (dl: apply dA/d ell instead of A + shift):
// every statement of gpx/lanczos.py:46-133 is differentiated
// (tangent of the recursion jit(value_and_grad) differentiates in reverse through the fori_loop;
// SURVEY.md 8f N1).  Per step: A [v | dv_l | dv_s] (three block mat-vecs) and (dK/d ell) v (one block
// mat-vec with the d/d ell kernel profile); dA/d sigma = 2 sigma I.
//   dalpha_out / dbeta_out: [2][P][m] (q = 0: lengthscale, q = 1: sigma).
int lanczos_grad_core(Ctx* ctx, int n_loc, long long N,
                      const std::function<int(const double*, double*, bool)>& matvec, double sigma, const double* U,
                      int P, int PS, int m, double* alpha_out, double* beta_out, double* dalpha_out,
                      double* dbeta_out, int Ptot, int* flag, cudaStream_t s) {
  GPXB_ARG_CHECK(P >= 1 && P <= MAXP && m >= 1, -30);
  ctx->shard_N = N;
  const size_t blk = (size_t)n_loc * PS;
  DevBuf bVecs, bDv[2], bW, bT[2], bX, bPart, bSc;
  GPXB_TRY(bVecs.alloc(sizeof(double) * blk * m, s));
  for (int q = 0; q < 2; ++q) {
    GPXB_TRY(bDv[q].alloc(sizeof(double) * blk * m, s));
    GPXB_TRY(bT[q].alloc(sizeof(double) * blk, s));
    GPXB_CUDA_CHECK(cudaMemsetAsync(bDv[q].as<double>(), 0, sizeof(double) * blk, s));  // d v_0 = 0
  }
  GPXB_TRY(bW.alloc(sizeof(double) * blk, s));
  GPXB_TRY(bX.alloc(sizeof(double) * blk, s));
  GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)ceil_div(n_loc, RB) * MAXP, s));
  GPXB_TRY(bSc.alloc(sizeof(double) * MAXP * 16, s));
  double *vecs = bVecs.as<double>(), *W = bW.as<double>(), *X = bX.as<double>(), *part = bPart.as<double>();
  double* dvecs[2] = {bDv[0].as<double>(), bDv[1].as<double>()};
  double* T[2] = {bT[0].as<double>(), bT[1].as<double>()};
  double* sc = bSc.as<double>();
  double *alpha = sc, *nrm2 = sc + MAXP, *coef = sc + 2 * MAXP, *beta = sc + 3 * MAXP, *t1 = sc + 4 * MAXP,
         *t2 = sc + 5 * MAXP;
  double* dalpha[2] = {sc + 6 * MAXP, sc + 7 * MAXP};
  double* dcoef[2] = {sc + 8 * MAXP, sc + 9 * MAXP};
  double* dbeta[2] = {sc + 10 * MAXP, sc + 11 * MAXP};
  double* wdw[2] = {sc + 12 * MAXP, sc + 13 * MAXP};
  GPXB_CUDA_CHECK(cudaMemsetAsync(sc, 0, sizeof(double) * MAXP * 16, s));
  GPXB_CUDA_CHECK(cudaMemcpyAsync(vecs, U, sizeof(double) * blk, cudaMemcpyDeviceToDevice, s));
  const int nthr = 256;
  const int nblk_e = ceil_div((long long)blk, nthr);

  auto axpy2 = [&](double* Wd, const double* a, const double* Xv, const double* b, const double* Yv) -> int {
    colaxpy_kernel<<<nblk_e, nthr, 0, s>>>(Wd, a, Xv, b, Yv, n_loc, PS, P);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
    return 0;
  };

  for (int j = 0; j < m; ++j) {
    double* vj = vecs + blk * j;
    double* dvj[2] = {dvecs[0] + blk * j, dvecs[1] + blk * j};
    const double* vjm = j > 0 ? vecs + blk * (j - 1) : nullptr;
    GPXB_TRY(matvec(vj, W, false));      // w = A v
    GPXB_TRY(matvec(vj, T[0], true));    // dw_l = (dK/d ell) v
    colscale_copy_kernel<<<nblk_e, nthr, 0, s>>>(vj, 2.0 * sigma, (long long)blk, T[1]);  // dw_s = 2 sigma v
    GPXB_LAUNCHED(ctx);
    if (j > 0) {
      for (int q = 0; q < 2; ++q) {      // + A dv
        GPXB_TRY(matvec(dvj[q], X, false));
        vadd_kernel<<<nblk_e, nthr, 0, s>>>(T[q], X, (long long)blk);
        GPXB_LAUNCHED(ctx);
      }
    }
    GPXB_TRY(coldot(ctx, W, vj, n_loc, PS, P, part, alpha, s));
    for (int q = 0; q < 2; ++q) {        // d alpha = dw . v + w . dv
      GPXB_TRY(coldot(ctx, T[q], vj, n_loc, PS, P, part, t1, s));
      GPXB_TRY(coldot(ctx, W, dvj[q], n_loc, PS, P, part, t2, s));
      padd_kernel<<<1, MAXP, 0, s>>>(t1, t2, P, dalpha[q]);
      GPXB_LAUNCHED(ctx);
      // dw -= d alpha v + alpha dv + d beta v_{j-1} + beta dv_{j-1}
      GPXB_TRY(axpy2(T[q], dalpha[q], vj, alpha, dvj[q]));
      if (j > 0) GPXB_TRY(axpy2(T[q], dbeta[q], vjm, beta, dvecs[q] + blk * (j - 1)));
    }
    GPXB_TRY(axpy2(W, alpha, vj, beta, vjm));  // w -= alpha v + beta v_{j-1}
    if (j > 0) {
      for (int c = 0; c <= j; ++c) {  // modified Gram-Schmidt and its tangent, column order
        const double* vc = vecs + blk * c;
        GPXB_TRY(coldot(ctx, W, vc, n_loc, PS, P, part, coef, s));
        for (int q = 0; q < 2; ++q) {
          const double* dvc = dvecs[q] + blk * c;
          GPXB_TRY(coldot(ctx, T[q], vc, n_loc, PS, P, part, t1, s));
          GPXB_TRY(coldot(ctx, W, dvc, n_loc, PS, P, part, t2, s));
          padd_kernel<<<1, MAXP, 0, s>>>(t1, t2, P, dcoef[q]);
          GPXB_LAUNCHED(ctx);
          GPXB_TRY(axpy2(T[q], dcoef[q], vc, coef, dvc));
        }
        GPXB_TRY(axpy2(W, coef, vc, nullptr, nullptr));
      }
    }
    GPXB_TRY(coldot(ctx, W, W, n_loc, PS, P, part, nrm2, s));
    for (int q = 0; q < 2; ++q) GPXB_TRY(coldot(ctx, W, T[q], n_loc, PS, P, part, wdw[q], s));
    store_ab_kernel<<<1, MAXP, 0, s>>>(alpha, nrm2, P, m, j, alpha_out, beta_out, beta);
    GPXB_LAUNCHED(ctx);
    for (int q = 0; q < 2; ++q) {
      store_dab_kernel<<<1, MAXP, 0, s>>>(dalpha[q], wdw[q], beta, P, m, j, dalpha_out + (size_t)q * Ptot * m,
                                          dbeta_out + (size_t)q * Ptot * m, dbeta[q]);
      GPXB_LAUNCHED(ctx);
    }
    if (j + 1 < m) {
      double* vn = vecs + blk * (j + 1);
      colnormalize_kernel<<<nblk_e, nthr, 0, s>>>(W, nrm2, vn, n_loc, PS, P, flag);
      GPXB_LAUNCHED(ctx);
      for (int q = 0; q < 2; ++q) {
        coltangent_kernel<<<nblk_e, nthr, 0, s>>>(T[q], vn, dbeta[q], beta, dvecs[q] + blk * (j + 1), n_loc, PS, P);
        GPXB_LAUNCHED(ctx);
      }
    }
  }
  return 0;
}

// Block Lanczos together with its forward-mode derivative w.r.t. (lengthscale, sigma) of
// A = K(x, x; ell) + (sigma^2 + 1e-10) I on the matrix-free stationary-kernel operator (row-sharded).
int lanczos_block_grad(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N,
                       ZLayout zl, double ell, double sigma, const double* U, int P, int PS, int m,
                       double* alpha_out, double* beta_out, double* dalpha_out, double* dbeta_out, int Ptot,
                       int* flag, cudaStream_t s) {
  const double diag_add = sigma * sigma + 1e-10;
  DevBuf bVall;
  double* Vall = nullptr;
  if (ctx->world > 1) {
    GPXB_TRY(bVall.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
    Vall = bVall.as<double>();
  }
  // out = Op * v for this rank's rows; dl: the d K / d ell operator instead of A
  auto matvec = [&](const double* v, double* out, bool dl) -> int {
    const double* vfull = v;
    if (ctx->world > 1) {
      GPXB_TRY(comm_allgather_rows(ctx, v, Vall, n_loc, PS, s));
      vfull = Vall;
    }
    KmvArgs a;
    a.kind = kind; a.Zr = Zr; a.n_rows = n_loc; a.row_offset = row0; a.Za = Za; a.N = N; a.zl = zl;
    a.V = vfull; a.PS = PS; a.P = P; a.W = out; a.ldw = PS;
    a.diag_add = dl ? 0.0 : diag_add;
    a.dl_ell = dl ? ell : 0.0;
    return kmatvec_launch(ctx, a, s);
  };
  return lanczos_grad_core(ctx, n_loc, N, matvec, sigma, U, P, PS, m, alpha_out, beta_out, dalpha_out, dbeta_out, Ptot,
                           flag, s);
}

int rademacher_probes(Ctx* ctx, const uint32_t key[2], long long Ntot, int P, long long row0, int n_rows,
                      int PS, int partitionable, double* U, cudaStream_t s) {
  const long long tot = (long long)n_rows * PS;
  rademacher_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(key[0], key[1], Ntot, P, row0, n_rows, PS,
                                                        partitionable, U);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Preconditioned CG (jax.scipy.sparse.linalg.cg semantics: x0 = 0, tol = 1e-5, atol given,
// stop when r.r <= max(tol^2 b.b, atol^2) or k == maxiter).
// ------------------------------------------------------------------------------------------
namespace {

__global__ void vdot_partial_kernel(const double* __restrict__ x, int sx, const double* __restrict__ y, int sy,
                                    int n, double* __restrict__ part) {
  __shared__ double sh[8];
  double s = 0.0;
  const long long r0 = (long long)blockIdx.x * 1024;
  for (int i = threadIdx.x; i < 1024; i += 256) {
    const long long r = r0 + i;
    if (r < n) s += x[r * sx] * y[r * sy];
  }
  s = block_sum<256>(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void vsum_final_kernel(const double* __restrict__ part, int nblk, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 1024) s += part[b];
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = s;
}
// y[i*sy] = a*x[i*sx] + b*y[i*sy] with device scalars a = num/den * sign (den may be null)
__global__ void vaxpby_kernel(double* __restrict__ y, int sy, const double* __restrict__ x, int sx, int n,
                              const double* __restrict__ num, const double* __restrict__ den, double sign,
                              double b) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = sign * (den ? (*num / *den) : *num);
  y[i * sy] = a * x[i * sx] + b * y[i * sy];
}
// p = z + (gn/g) p
__global__ void vxpby_kernel(double* __restrict__ p, int sp, const double* __restrict__ z, int n,
                             const double* __restrict__ gn, const double* __restrict__ g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i * sp] = z[i] + (*gn / *g) * p[i * sp];
}
// t1[c] = sum_r F[c][r] z[r]   (one CTA per pivot column; F stored [k][ldf])
__global__ void ftz_kernel(const double* __restrict__ F, long long ldf, int n, const double* __restrict__ z,
                           double* __restrict__ t1) {
  __shared__ double sh[32];
  const double* f = F + (long long)blockIdx.x * ldf;
  double s = 0.0;
  for (int r = threadIdx.x; r < n; r += 1024) s += f[r] * z[r];
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) t1[blockIdx.x] = s;
}
// t2 = Pkk t1 (k x k full symmetric, tiny)
__global__ void small_matvec_kernel(const double* __restrict__ Pkk, int k, const double* __restrict__ t1,
                                    double* __restrict__ t2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  double s = 0.0;
  for (int c = 0; c < k; ++c) s += Pkk[(long long)i * k + c] * t1[c];
  t2[i] = s;
}
// out[r] = -(sigma^-2) * sum_c F[c][r] t2[c] + sigma^-2 z[r]     (gpx/models/_gpr.py:208-214)
__global__ void precond_apply_kernel(const double* __restrict__ F, long long ldf, int k, int n,
                                     const double* __restrict__ t2, const double* __restrict__ z,
                                     double isig2, double* __restrict__ out) {
  extern __shared__ double st2[];
  for (int c = threadIdx.x; c < k; c += blockDim.x) st2[c] = t2[c];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double s = 0.0;
  for (int c = 0; c < k; ++c) s += F[(long long)c * ldf + r] * st2[c];
  out[r] = -isig2 * s + isig2 * z[r];
}
// symmetrise a lower-stored k x k matrix in place and add nothing else
__global__ void symmetrize_kernel(double* __restrict__ A, int k) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k * k) return;
  const int i = idx / k, j = idx % k;
  if (j > i) A[idx] = A[(long long)j * k + i];
}
__global__ void add_diag_kernel(double* __restrict__ A, int k, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) A[(long long)i * k + i] += v;
}
__global__ void pad_copy_kernel(const double* __restrict__ x, int n, double* __restrict__ xp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xp[2 * i] = x[i];
  xp[2 * i + 1] = 0.0;
}

struct VecOps {
  Ctx* ctx;
  cudaStream_t s;
  double* part;
  int n;
  int dot(const double* x, int sx, const double* y, int sy, double* out) {
    const int nblk = ceil_div(n, 1024);
    {
      PhaseScope ph(ctx, PH_DOTS, s);
      vdot_partial_kernel<<<nblk, 256, 0, s>>>(x, sx, y, sy, n, part);
      GPXB_LAUNCH_CHECK();
      vsum_final_kernel<<<1, 1024, 0, s>>>(part, nblk, out);
      GPXB_LAUNCH_CHECK();
      ctx->launches += 2;
    }
    return comm_allreduce_sum(ctx, out, 1, s);
  }
};

// out[i] = element (row0 + i) of jax.random.normal(key, (N,), float64): sqrt(2) erfinv(u), u uniform on
// [nextafter(-1, 0), 1) (jax._src.random._normal_real); erfinv is CUDA's (a few ulp from SciPy's / XLA's)
__global__ void jax_normal_kernel(uint32_t k0, uint32_t k1, long long N, long long row0, int n, int partitionable,
                                  double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lo = -0.99999999999999989;  // nextafter(-1, 0)
  double u = jax_uniform_f64(k0, k1, (unsigned long long)(row0 + i), (unsigned long long)N, partitionable);
  u = fmax(lo, u * (1.0 - lo) + lo);
  out[i] = 1.4142135623730951 * erfinv(u);
}
// t -= coef * v (v with stride sv)
__global__ void restart_axpy_kernel(double* __restrict__ t, const double* __restrict__ v, int sv, int n,
                                    const double* __restrict__ coef) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] -= coef[0] * v[i * sv];
}
// dst (stride sd) = t / sqrt(nrm2)
__global__ void restart_store_kernel(const double* __restrict__ t, const double* __restrict__ nrm2,
                                     double* __restrict__ dst, int sd, int n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i * sd] = t[i] / sqrt(nrm2[0]);
}

}  // namespace

int lanczos_restart_columns(Ctx* ctx, cudaStream_t s, double* vecs, size_t blk, int n_loc, long long N, int PS, int j,
                            const std::vector<int>& bad, uint32_t sk0, uint32_t sk1, int partitionable) {
  DevBuf bR, bT, bPart, bSc;
  GPXB_TRY(bR.alloc(sizeof(double) * std::max(n_loc, 1), s));
  GPXB_TRY(bT.alloc(sizeof(double) * std::max(n_loc, 1), s));
  GPXB_TRY(bPart.alloc(sizeof(double) * (ceil_div(n_loc, 1024) + 1), s));
  GPXB_TRY(bSc.alloc(sizeof(double) * 2, s));
  double *rv = bR.as<double>(), *t = bT.as<double>(), *sc = bSc.as<double>();
  const long long row0 = (n_loc == N || ctx->world <= 1) ? 0 : shard_begin(N, ctx->world, ctx->rank);
  const int nb = std::max(ceil_div(n_loc, 256), 1);
  jax_normal_kernel<<<nb, 256, 0, s>>>(sk0, sk1, N, row0, n_loc, partitionable, rv);  // the same draw for every probe
  GPXB_LAUNCHED(ctx);
  VecOps vo{ctx, s, bPart.as<double>(), n_loc};
  for (int p : bad) {
    GPXB_CUDA_CHECK(cudaMemcpyAsync(t, rv, sizeof(double) * n_loc, cudaMemcpyDeviceToDevice, s));
    // orthonormalize(randv, vecs): modified Gram-Schmidt over the columns of vecs in order; the columns beyond
    // j are still zero in the reference (vecs.at[:, j + 1] is set after this call) and change nothing
    for (int c = 0; c <= j; ++c) {
      const double* vc = vecs + blk * c + p;
      GPXB_TRY(vo.dot(t, 1, vc, PS, sc));
      restart_axpy_kernel<<<nb, 256, 0, s>>>(t, vc, PS, n_loc, sc);
      GPXB_LAUNCHED(ctx);
    }
    GPXB_TRY(vo.dot(t, 1, t, 1, sc + 1));
    restart_store_kernel<<<nb, 256, 0, s>>>(t, sc + 1, vecs + blk * (j + 1) + p, PS, n_loc);
    GPXB_LAUNCHED(ctx);
  }
  return 0;
}

// Builds the RPCholesky preconditioner pieces from F (stored [k][ldf], local rows):
// Pkk = inv(sigma^2 I + F^T F) (full symmetric k x k, device).
int precond_build(Ctx* ctx, const double* F, long long ldf, int k, int n_loc, double sigma, double* Pkk,
                  cudaStream_t s) {
  GemmArgs g;  // C = F F^T over the local rows  (A stored [k][n]: K-major both)
  g.M = k; g.N = k; g.K = n_loc;
  g.A = F; g.lda = ldf; g.a_kmajor = true;
  g.B = F; g.ldb = ldf; g.b_kmajor = true;
  g.C = Pkk; g.ldc = k;
  g.lower_only = 1;
  GPXB_CUDA_CHECK(cudaMemsetAsync(Pkk, 0, sizeof(double) * (size_t)k * k, s));
  GPXB_TRY(gemm_f64(ctx, g, s));
  symmetrize_kernel<<<ceil_div((long long)k * k, 256), 256, 0, s>>>(Pkk, k);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  GPXB_TRY(comm_allreduce_sum(ctx, Pkk, (size_t)k * k, s));
  add_diag_kernel<<<ceil_div(k, 256), 256, 0, s>>>(Pkk, k, sigma * sigma);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  DevBuf inv, W, T;
  GPXB_TRY(inv.alloc(sizeof(double) * leaf_inv_doubles(k), s));
  GPXB_TRY(W.alloc(sizeof(double) * (size_t)k * k, s));
  GPXB_TRY(T.alloc(sizeof(double) * potri_scratch_doubles(k), s));
  GPXB_TRY(potrf_lower(ctx, Pkk, k, k, inv.as<double>(), s));
  GPXB_TRY(potri_lower(ctx, Pkk, k, inv.as<double>(), k, W.as<double>(), T.as<double>(), Pkk, k, s));
  symmetrize_kernel<<<ceil_div((long long)k * k, 256), 256, 0, s>>>(Pkk, k);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

// Preconditioned CG as a resumable object: init() -> while (active()) { Ap = A p(); step(Ap); }.
// (K + diag_add I) x = b with b the local rows (length n_loc); F / Pkk: preconditioner (k may be 0: identity).
// Host-synchronous: one scalar read per iteration for the stopping test.  p() is the local search direction
// stored with stride 2 (the padded single-column layout of the mat-vec).
struct PcgRun {
  Ctx* ctx = nullptr;
  cudaStream_t s = nullptr;
  int n_loc = 0, k = 0, maxiter = 0, it = 0, nb = 0;
  const double* F = nullptr;
  long long ldf = 0;
  const double* Pkk = nullptr;
  double isig2 = 0.0, atol2 = 0.0, rs = 0.0;
  double* x = nullptr;
  DevBuf bR, bZ, bP, bAp, bPart, bSc, bT;
  double *r = nullptr, *z = nullptr, *pv = nullptr, *Ap = nullptr, *sc = nullptr, *t1 = nullptr, *t2 = nullptr;

  int dot(const double* a, int sa, const double* b, int sb, double* out) {
    VecOps vo{ctx, s, bPart.as<double>(), n_loc};
    return vo.dot(a, sa, b, sb, out);
  }
  int precond(const double* in, double* out) {
    if (k <= 0) {
      GPXB_CUDA_CHECK(cudaMemcpyAsync(out, in, sizeof(double) * n_loc, cudaMemcpyDeviceToDevice, s));
      return 0;
    }
    PhaseScope ph(ctx, PH_PRECOND, s);
    ftz_kernel<<<k, 1024, 0, s>>>(F, ldf, n_loc, in, t1);
    GPXB_LAUNCH_CHECK();
    GPXB_TRY(comm_allreduce_sum(ctx, t1, k, s));
    small_matvec_kernel<<<ceil_div(k, 128), 128, 0, s>>>(Pkk, k, t1, t2);
    GPXB_LAUNCH_CHECK();
    precond_apply_kernel<<<nb, 256, sizeof(double) * k, s>>>(F, ldf, k, n_loc, t2, in, isig2, out);
    GPXB_LAUNCH_CHECK();
    ctx->launches += 3;
    return 0;
  }
  int init(Ctx* ctx_, int n_loc_, long long N, const double* b, const double* F_, long long ldf_, int k_,
           const double* Pkk_, double sigma, double tol, double atol, int maxiter_, double* x_, cudaStream_t s_) {
    ctx = ctx_; s = s_; n_loc = n_loc_; k = k_; maxiter = maxiter_; F = F_; ldf = ldf_; Pkk = Pkk_; x = x_;
    ctx->shard_N = N;
    GPXB_TRY(bR.alloc(sizeof(double) * n_loc, s));
    GPXB_TRY(bZ.alloc(sizeof(double) * n_loc, s));
    GPXB_TRY(bP.alloc(sizeof(double) * ((size_t)n_loc * 2 + 16), s));
    GPXB_TRY(bAp.alloc(sizeof(double) * n_loc, s));
    GPXB_TRY(bPart.alloc(sizeof(double) * (ceil_div(n_loc, 1024) + 1), s));
    GPXB_TRY(bSc.alloc(sizeof(double) * 8, s));
    GPXB_TRY(bT.alloc(sizeof(double) * (2 * (size_t)std::max(k, 1)), s));
    r = bR.as<double>(); z = bZ.as<double>(); pv = bP.as<double>(); Ap = bAp.as<double>();
    sc = bSc.as<double>();  // [0] bs, [1] gamma, [2] pAp, [3] gamma_new, [4] rs
    t1 = bT.as<double>(); t2 = t1 + std::max(k, 1);
    nb = ceil_div(n_loc, 256);
    isig2 = 1.0 / (sigma * sigma);
    GPXB_CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * n_loc, s));
    GPXB_CUDA_CHECK(cudaMemcpyAsync(r, b, sizeof(double) * n_loc, cudaMemcpyDeviceToDevice, s));
    GPXB_TRY(dot(b, 1, b, 1, sc + 0));
    GPXB_TRY(precond(r, z));
    {
      PhaseScope ph(ctx, PH_VECOPS, s);
      pad_copy_kernel<<<nb, 256, 0, s>>>(z, n_loc, pv);
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
    }
    GPXB_TRY(dot(r, 1, z, 1, sc + 1));
    double h[2];
    GPXB_CUDA_CHECK(cudaMemcpyAsync(h, sc, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
    GPXB_CUDA_CHECK(cudaStreamSynchronize(s));
    atol2 = std::max(tol * tol * h[0], atol * atol);
    rs = (k <= 0) ? h[1] : h[0];  // r0 = b
    it = 0;
    return 0;
  }
  bool active() const { return rs > atol2 && it < maxiter; }
  const double* p() const { return pv; }
  double* p_mut() { return pv; }
  double* ap() { return Ap; }
  // the updates of one CG iteration once Ap = A p is in ap()
  int step() {
    GPXB_TRY(dot(pv, 2, Ap, 1, sc + 2));
    {
      PhaseScope ph(ctx, PH_VECOPS, s);
      vaxpby_kernel<<<nb, 256, 0, s>>>(x, 1, pv, 2, n_loc, sc + 1, sc + 2, 1.0, 1.0);    // x += alpha p
      vaxpby_kernel<<<nb, 256, 0, s>>>(r, 1, Ap, 1, n_loc, sc + 1, sc + 2, -1.0, 1.0);  // r -= alpha Ap
      GPXB_LAUNCH_CHECK();
      ctx->launches += 2;
    }
    GPXB_TRY(precond(r, z));
    GPXB_TRY(dot(r, 1, z, 1, sc + 3));
    {
      PhaseScope ph(ctx, PH_VECOPS, s);
      vxpby_kernel<<<nb, 256, 0, s>>>(pv, 2, z, n_loc, sc + 3, sc + 1);  // p = z + (gn/g) p
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
    }
    GPXB_TRY(dot(r, 1, r, 1, sc + 4));
    double h[2];
    GPXB_CUDA_CHECK(cudaMemcpyAsync(sc + 1, sc + 3, sizeof(double), cudaMemcpyDeviceToDevice, s));
    GPXB_CUDA_CHECK(cudaMemcpyAsync(h, sc + 3, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
    GPXB_CUDA_CHECK(cudaStreamSynchronize(s));
    rs = (k <= 0) ? h[0] : h[1];
    ++it;
    return 0;
  }
};

// Solves (K + diag_add I) x = b (b: local rows, length n_loc).  F/Pkk: preconditioner (k may be 0:
// identity).  Host-synchronous: one scalar read per iteration for the stopping test.
// apply(p, Ap): Ap[n_loc] = A p for this rank's rows; p is the local search direction stored with stride 2.
int pcg_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
             const double* b, const double* F, long long ldf, int k, const double* Pkk, double sigma, double tol,
             double atol, int maxiter, double* x, int* iters_out, cudaStream_t s) {
  PcgRun run;
  GPXB_TRY(run.init(ctx, n_loc, N, b, F, ldf, k, Pkk, sigma, tol, atol, maxiter, x, s));
  while (run.active()) {
    GPXB_TRY(apply(run.p(), run.ap()));
    GPXB_TRY(run.step());
  }
  if (iters_out) *iters_out = run.it;
  return 0;
}

// PCG on the matrix-free stationary-kernel operator (row-sharded: the search direction is all-gathered)
int pcg_solve(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N,
              ZLayout zl, double diag_add, const double* b, const double* F, long long ldf, int k,
              const double* Pkk, double sigma, double tol, double atol, int maxiter, double* x, int* iters_out,
              cudaStream_t s) {
  DevBuf bPall;
  double* Pall = nullptr;
  if (ctx->world > 1) {
    GPXB_TRY(bPall.alloc(sizeof(double) * ((size_t)N * 2 + 16), s));
    Pall = bPall.as<double>();
  }
  auto apply = [&](const double* p, double* Ap) -> int {
    const double* pfull = p;
    if (ctx->world > 1) {
      GPXB_TRY(comm_allgather_rows(ctx, p, Pall, n_loc, 2, s));
      pfull = Pall;
    }
    KmvArgs a;
    a.kind = kind; a.Zr = Zr; a.n_rows = n_loc; a.row_offset = row0; a.Za = Za; a.N = N; a.zl = zl;
    a.V = pfull; a.PS = 2; a.P = 1; a.W = Ap; a.ldw = 1; a.diag_add = diag_add;
    return kmatvec_launch(ctx, a, s);
  };
  return pcg_core(ctx, n_loc, N, apply, b, F, ldf, k, Pkk, sigma, tol, atol, maxiter, x, iters_out, s);
}


namespace {
// block[i][col] = p[2 i]   /   out[i] = W[i][col]
__global__ void put_col_kernel(double* __restrict__ block, int PS, int col, const double* __restrict__ p, int n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) block[i * PS + col] = p[2 * i];
}
__global__ void take_col_kernel(const double* __restrict__ W, int PS, int col, double* __restrict__ out, int n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = W[i * PS + col];
}
}  // namespace

// The two device parts of the iterative LML (_lml_iter, gpx/models/_gpr.py:772-821) in one sweep: the PCG solve
// c = (K + s2 I)^-1 b (RPCholesky preconditioner F / Pkk) and the m-step block Lanczos on the P start vectors U
// (SLQ log-det).  While both run, the PCG search direction rides as column P of the Lanczos block, so the first m
// CG iterations cost no extra pass over the recomputed kernel tiles; the remaining iterations use the
// single-vector mat-vec.  The arithmetic of each part is the one of pcg_solve / lanczos_block.
int lml_iter_fused(Ctx* ctx, int kind, const double* Zr, int n_loc, long long row0, const double* Za, int N, ZLayout zl,
                   double diag_add, const double* b, const double* F, long long ldf, int k, const double* Pkk,
                   double sigma, double tol, double atol, int maxiter, double* x, int* iters_out, const double* U,
                   int P, int PS, int m, double* alpha_out, double* beta_out, int* flag, cudaStream_t s) {
  GPXB_ARG_CHECK(P >= 1 && P < 32 && PS >= P + 1, -31);
  DevBuf bVall, bPall;
  double *Vall = nullptr, *Pall = nullptr;
  if (ctx->world > 1) {
    GPXB_TRY(bVall.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
    GPXB_TRY(bPall.alloc(sizeof(double) * ((size_t)N * 2 + 16), s));
    Vall = bVall.as<double>();
    Pall = bPall.as<double>();
  }
  PcgRun run;
  GPXB_TRY(run.init(ctx, n_loc, N, b, F, ldf, k, Pkk, sigma, tol, atol, maxiter, x, s));
  int pcols = P;  // columns of the block the operator is applied to (P + 1 while the PCG vector rides)
  auto apply_block = [&](const double* v, double* w) -> int {
    const double* vfull = v;
    if (ctx->world > 1) {
      GPXB_TRY(comm_allgather_rows(ctx, v, Vall, n_loc, PS, s));
      vfull = Vall;
    }
    KmvArgs a;
    a.kind = kind; a.Zr = Zr; a.n_rows = n_loc; a.row_offset = row0; a.Za = Za; a.N = N; a.zl = zl;
    a.V = vfull; a.PS = PS; a.P = pcols; a.W = w; a.ldw = PS; a.diag_add = diag_add;
    return kmatvec_launch(ctx, a, s);
  };
  const int nb = ceil_div(n_loc, 256);
  LanczosRider rider;
  rider.active = [&]() {
    pcols = run.active() ? P + 1 : P;
    return run.active();
  };
  rider.put = [&](double* block, int col, int ps) -> int {
    PhaseScope ph(ctx, PH_VECOPS, s);
    put_col_kernel<<<nb, 256, 0, s>>>(block, ps, col, run.p(), n_loc);
    GPXB_LAUNCHED(ctx);
    return 0;
  };
  rider.take = [&](const double* W, int col, int ps) -> int {
    {
      PhaseScope ph(ctx, PH_VECOPS, s);
      take_col_kernel<<<nb, 256, 0, s>>>(W, ps, col, run.ap(), n_loc);
      GPXB_LAUNCHED(ctx);
    }
    return run.step();
  };
  GPXB_TRY(lanczos_core_rider(ctx, n_loc, N, apply_block, U, P, PS, m, alpha_out, beta_out, flag, &rider, s));
  while (run.active()) {  // the rest of the CG iterations on the single-vector mat-vec
    const double* pfull = run.p();
    if (ctx->world > 1) {
      GPXB_TRY(comm_allgather_rows(ctx, run.p(), Pall, n_loc, 2, s));
      pfull = Pall;
    }
    KmvArgs a;
    a.kind = kind; a.Zr = Zr; a.n_rows = n_loc; a.row_offset = row0; a.Za = Za; a.N = N; a.zl = zl;
    a.V = pfull; a.PS = 2; a.P = 1; a.W = run.ap(); a.ldw = 1; a.diag_add = diag_add;
    GPXB_TRY(kmatvec_launch(ctx, a, s));
    GPXB_TRY(run.step());
  }
  if (iters_out) *iters_out = run.it;
  return 0;
}

}  // namespace gpxb
