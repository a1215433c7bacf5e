/* libgpxb200 -- B200-native (sm_100a) Gaussian-process hot path behind a plain C ABI.
 *
 * The reference (Molecolab-Pisa/GPX) is pure Python on JAX and has no FFI of its own; the seam
 * this library sits behind is the set of jit'd numerical-core functions the gpx.models facade
 * calls.  Each entry point cites the reference function it replaces (paths relative to the
 * GPX repository root).  Under JAX these symbols are what an XLA-FFI (jax.ffi) custom call
 * would bind; here they are bound by ctypes (gpx_b200/_lib.py).
 *
 * Conventions
 *   - every pointer argument is a DEVICE pointer to row-major FP64 (or int64 / uint32 where
 *     stated), owned by the caller; outputs are pre-allocated by the caller;
 *   - every call is enqueue-only on the given CUDA stream (no host synchronisation) and
 *     re-entrant per context; scratch memory is stream-ordered (cudaMallocAsync);
 *   - return value: 0 = OK, < 0 = argument error, > 0 = CUDA / NCCL error code;
 *     gpxb_last_error() returns the message of the calling thread's last failure;
 *   - numerical failure (a non positive-definite matrix) is NOT an error: NaN propagates to
 *     the outputs exactly as jax.scipy.linalg.cholesky does in the reference.
 *   - kind: 0 = SquaredExponential, 1 = Matern12, 2 = Matern32, 3 = Matern52
 *     (gpx/kernels/kernels.py:1623-1754); lengthscale is a scalar.
 */
#ifndef GPXB200_H
#define GPXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpxb_ctx gpxb_ctx;
typedef void* gpxb_stream; /* cudaStream_t */

enum { GPXB_SE = 0, GPXB_M12 = 1, GPXB_M32 = 2, GPXB_M52 = 3 };
enum { GPXB_MEAN_ZERO = 0, GPXB_MEAN_DATA = 1 }; /* gpx/mean_functions.py:25-30 */

/* ---- library / context ------------------------------------------------------------- */
int gpxb_version(void);
int gpxb_last_error(char* buf, size_t buflen);
int gpxb_ctx_create(int device, gpxb_ctx** out);
int gpxb_ctx_destroy(gpxb_ctx* ctx);
/* number of CUDA kernels this library has launched through ctx (bench.py "gpu_launches") */
unsigned long long gpxb_ctx_launches(gpxb_ctx* ctx);
/* Measurement aid (bench.py "roofline"): while enabled, every DMMA GEMM launch is bracketed by CUDA
 * events on the stream it is launched on; collect synchronises the device and returns the summed
 * launch durations, their algorithmic flops and the launch count, then clears the records.  on = 2 also keeps
 * the look-ahead work of the Cholesky factorisation on the caller's stream, so that no two launches overlap and
 * the summed durations are pure kernel time. */
int gpxb_prof_enable(gpxb_ctx* ctx, int on);
int gpxb_prof_collect(gpxb_ctx* ctx, double* total_ms, double* total_flops, int* n_launches);

/* Measurement aid for the row-sharded paths (bench.py "phases"): while enabled, the phases of gpxb_lanczos,
 * gpxb_gpr_fit_iter, gpxb_rpcholesky and gpxb_sgpr_lml are bracketed by CUDA events on the stream they run on;
 * collect synchronises the device and returns, per phase id, the summed durations (ms) and the number of brackets,
 * then clears the records.  Collective phases include the time spent waiting for the slowest rank. */
enum {
  GPXB_PH_MATVEC = 0,     /* matrix-free block mat-vec */
  GPXB_PH_ALLGATHER = 1,  /* all-gather of the N x P vector block */
  GPXB_PH_DOTS = 2,       /* local column dot products / norms */
  GPXB_PH_ALLREDUCE = 3,  /* every all-reduce */
  GPXB_PH_VECOPS = 4,     /* axpy / normalise passes over the local vector blocks */
  GPXB_PH_RP_PIVOT = 5,   /* RPCholesky: draw, inverse-CDF search, pivot row gather */
  GPXB_PH_RP_COLUMN = 6,  /* RPCholesky: fused kernel column / GEMV / diagonal update */
  GPXB_PH_SGPR_CHUNK = 7, /* SGPR: K_nm block, TRSM against L_m, SYRK into B (the chunk streams overlap) */
  GPXB_PH_POTRF_REPL = 8, /* SGPR: replicated M x M factorisations and final solves */
  GPXB_PH_PRECOND = 9,    /* PCG: preconditioner application */
  GPXB_PH_COUNT = 16
};
int gpxb_phase_enable(gpxb_ctx* ctx, int on);
int gpxb_phase_collect(gpxb_ctx* ctx, double* ms_out, int* count_out, int n);

/* ---- Family 1: fused kernel-matrix construction ------------------------------------- */
/* out[n1 x n2] (ld ldo) = k(x1, x2) + diag_add * I.  Replaces Kernel.k -> squared_distances +
 * nonlinearity (gpx/utils.py:55-72, gpx/kernels/kernels.py:751-757, 922-927, 966-971, 1050-1055)
 * and the noise diagonal of _A_lhs (gpx/models/_gpr.py:39-63).  x2 == x1 (same pointer, n2 == n1)
 * selects the symmetric path; lower_only then writes only entries with row >= col. */
int gpxb_gram(gpxb_ctx* ctx, int kind, const double* x1, int n1, const double* x2, int n2, int D,
              double ell, double diag_add, double* out, long long ldo, int lower_only,
              gpxb_stream stream);

/* Gram matrix of a Sum / Prod tree of kernels in ONE tile-level pass (gpx/kernels/kernels.py:1794-1848 Sum, :1851-1942
 * Prod; leaves: kinds 0..3 as in gpxb_gram, 4 = Linear, kernels.py:226-236).  Leaf l sees its own feature slice
 * x1_leaf[l] (n1 x leaf_D[l]) / x2_leaf[l] (n2 x leaf_D[l]) and lengthscale ells[l] (ignored for Linear).  prog: the
 * tree in postfix order, entries >= 0 are leaf indices, -1 = Sum, -2 = Prod (at most 8 leaves, stack depth <= 4).
 * sym: x2_leaf == x1_leaf (exact d2 = 1e-12 on the diagonal of the stationary leaves); diag_add, lower_only as in
 * gpxb_gram.  No component matrix is written.  gpxb_gram_composite_fits: 1 when the leaves' operand tiles fit in
 * shared memory (sum over the leaves of 2 * 128 * S(leaf_D) * 8 bytes <= 227 KB), else the caller combines component
 * matrices with gpxb_elementwise. */
int gpxb_gram_composite_fits(int nleaf, const int* leaf_D);
int gpxb_gram_composite(gpxb_ctx* ctx, int nleaf, const int* kinds, const double* ells, const int* leaf_D,
                        const double* const* x1_leaf, const double* const* x2_leaf, int n1, int n2, const int* prog,
                        int nprog, double diag_add, int sym, int lower_only, double* out, long long ldo,
                        gpxb_stream stream);
/* Composite (Sum / Prod tree) kernels outside the dense paths; leaf / program description as in gpxb_gram_composite,
 * x_leaf[l]: N x leaf_D[l].  Single GPU.
 * gpxb_rpcholesky_composite: randomly pivoted Cholesky (gpx/kernels/approximations.py:30-129) with the composite kernel:
 *   the pivot's column is evaluated pair by pair (external-column mode of the pivot loop); F_out may be NULL.
 * gpxb_gpr_iter_composite: the iterative GPR paths (_fit_iter gpx/models/_gpr.py:285-310 with _precond_rpcholesky
 *   :180-216; the SLQ part of _lml_iter :772-821) on the operator K + (sigma^2+1e-10) I applied in row chunks (fused
 *   composite Gram of chunk x all rows, then one DMMA GEMM with the block).  c_out != NULL: PCG fit (y, mu_out,
 *   iters_out as in gpxb_gpr_fit_iter); U != NULL: m-step block Lanczos as in gpxb_lanczos.  The leaves' operand
 *   tiles must fit in shared memory (gpxb_gram_composite_fits). */
int gpxb_rpcholesky_composite(gpxb_ctx* ctx, int nleaf, const int* kinds, const double* ells, const int* leaf_D,
                              const double* const* x_leaf, int N, const int* prog, int nprog, uint32_t key0,
                              uint32_t key1, int M, int partitionable, double* F_out, long long* pivots_out,
                              gpxb_stream stream);
int gpxb_gpr_iter_composite(gpxb_ctx* ctx, int nleaf, const int* kinds, const double* ells, const int* leaf_D,
                            const double* const* x_leaf, int N, const int* prog, int nprog, const double* y,
                            double sigma, int mean_mode, int n_pivots, uint32_t key0, uint32_t key1, int partitionable,
                            double tol, double atol, int maxiter, double* c_out, double* mu_out, int* iters_out,
                            const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                            int* breakdown_flag, gpxb_stream stream);
/* *out (device) = sum_ij Y_ij d k(x1_i, x2_j)/d lengthscale without forming dK (the contraction epilogue of
 * the Gram kernel): the building block of hyper-parameter gradients of composite kernels
 * (gpx/kernels/kernels.py:1794-1942 Sum / Prod under value_and_grad).  sym_lower (identical operands): Y is the
 * lower triangle of a symmetric matrix and the sum runs over the full matrix. */
int gpxb_gram_dl_contract(gpxb_ctx* ctx, int kind, const double* x1, int n1, const double* x2, int n2, int D,
                          double ell, const double* Y, long long ldy, int sym_lower, double* out,
                          gpxb_stream stream);
/* out[i] = a X[i] + b Y[i] (op = 0) or a X[i] Y[i] (op = 1) on n contiguous doubles; out may alias X or Y
 * (sum_kernels / prod_kernels, gpx/kernels/operations.py:103-426, applied to device matrices). */
int gpxb_elementwise(gpxb_ctx* ctx, int op, long long n, double a, const double* X, double b, const double* Y,
                     double* out, gpxb_stream stream);

/* ---- Family 1b: Hessian-Jacobian kernel (training on derivative values) --------------- */
/* out[(n1*nv1) x (n2*nv2)] = Kernel.d01kj(x1, x2, params, J1, J2) (+ diag_add * I when both operands
 * are the same arrays); row index = sample * nv + variable.  x: n x nf, J: n x nf x nv row-major.
 * kind: GPXB_SE (gpx/kernels/kernels.py:804-831) or GPXB_M52 (:1253-1283); _A_derivs_lhs
 * (gpx/models/_gpr.py:66-91).  lower_only requires identical operands. */
int gpxb_hess_gram(gpxb_ctx* ctx, int kind, const double* x1, const double* J1, int n1, int nv1,
                   const double* x2, const double* J2, int n2, int nv2, int nf, double ell, double diag_add,
                   double* out, long long ldo, int lower_only, gpxb_stream stream);
/* *out (device) = lml / (N nv) of _lml_derivs_dense (gpx/models/_gpr.py:824-870; zero mean as forced by
 * gpx/models/gpr.py:160, 173).  y: N x nv. */
int gpxb_gpr_lml_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf,
                        int nv, double ell, double sigma, double* out, gpxb_stream stream);
/* out3 (device) = {lml, d lml/d lengthscale, d lml/d sigma} of _lml_derivs_dense: the value and gradient
 * jit(value_and_grad(loss)) yields for neg_log_marginal_likelihood_derivs (gpx/models/gpr.py:276-300) inside
 * scipy_minimize_derivs (gpx/optimizers/scipy_optimize.py:92-122), up to the sign and the bijector chain rule
 * the host applies.  Analytic: W = alpha alpha^T - A^-1, d lml/d l = (1/2 sum W o dA/dl)/(N nv) with the
 * Hessian kernel re-evaluated in contraction mode (dA/dl is never stored).  want_grad = 0: out3[1..2] = 0. */
int gpxb_gpr_lml_grad_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N,
                             int nf, int nv, double ell, double sigma, int want_grad, double* out3,
                             gpxb_stream stream);
/* c[N nv] = (d01kj + (sigma^2+1e-10) I)^-1 y and jaccoef[N x nf] = einsum("sv,sfv->sf", c, J)
 * (gpx/models/_gpr.py:313-344, gpx/models/gpr.py:475-476); jaccoef_out may be NULL. */
int gpxb_gpr_fit_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N, int nf,
                        int nv, double ell, double sigma, double* c_out, double* jaccoef_out, gpxb_stream stream);

/* mean_out[ns] = mu + sum_s Kernel.d0kjc(x_train, x, params, jaccoef)[s, :]: target values predicted by a
 * model trained on derivatives (_predict_y_derivs_dense fast path, gpx/models/_gpr.py:630-638; d0kjc =
 * grad_kernelize(argnums=0, with_jaccoef) gpx/kernels/kernels.py:1474-1480).  Evaluated matrix-free as one
 * block mat-vec with nf + 1 columns.  jaccoef: N x nf (from gpxb_gpr_fit_derivs); mu: host scalar. */
int gpxb_gpr_predict_y_derivs(gpxb_ctx* ctx, int kind, const double* x_train, const double* jaccoef, int N, int nf,
                              const double* x, int ns, double ell, double mu, double* mean_out,
                              gpxb_stream stream);

/* W[(n1 nv1) x P] = (d01kj(x1, x2; J1, J2) [+ diag_add I]) V without forming the matrix: _Ax_derivs_lhs_fun
 * (gpx/models/_gpr.py:122-152; rowfun_to_matvec, gpx/operations.py:54-123).  The sum over the (sample,
 * variable) columns collapses onto the contracted jacobian of V, so the cost is three (n1 x n2 x nf) GEMMs
 * per column instead of an (n1 nv1)(n2 nv2) nf product.  diag_add needs identical operands. */
int gpxb_hess_matvec(gpxb_ctx* ctx, int kind, const double* x1, const double* J1, int n1, int nv1, const double* x2,
                     const double* J2, int n2, int nv2, int nf, double ell, double diag_add, const double* V,
                     long long ldv, int P, double* W, long long ldw, gpxb_stream stream);
/* Iterative fit on derivative values: c[N nv] = PCG((d01kj + (sigma^2+1e-10) I), y) with the
 * rpcholesky_derivs preconditioner (_fit_derivs_iter gpx/models/_gpr.py:347-390, _precond_rpcholesky_derivs
 * :219-256, rpcholesky_derivs gpx/kernels/approximations.py:132-242), tol / atol / maxiter as
 * jax.scipy.sparse.linalg.cg (maxiter <= 0: 10 N nv).  jaccoef_out, iters_out (host) may be NULL. */
int gpxb_gpr_fit_derivs_iter(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N,
                             int nf, int nv, double ell, double sigma, int n_pivots, uint32_t key0, uint32_t key1,
                             int partitionable, double tol, double atol, int maxiter, double* c_out,
                             double* jaccoef_out, int* iters_out, gpxb_stream stream);
/* W = (d d01kj(x, x; J, J)/d lengthscale) V, matrix-free (the c^T dA c term of the CG solve's cotangent), and
 * gpxb_lanczos_derivs with the forward-mode derivative of the tridiagonals w.r.t. (lengthscale, sigma)
 * [2 x P x m]: together the gradient jit(value_and_grad) yields for the iterative derivative LML
 * (_lml_derivs_iter, gpx/models/_gpr.py:873-935), as for gpxb_lanczos_grad / gpxb_kmatvec_dl. */
int gpxb_hess_matvec_dl(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nv, int nf, double ell,
                        const double* V, long long ldv, int P, double* W, long long ldw, gpxb_stream stream);
int gpxb_lanczos_grad_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv,
                             double ell, double sigma, const double* U, long long ldu, int P, int m,
                             double* alpha_out, double* beta_out, double* dalpha_out, double* dbeta_out,
                             int* breakdown_flag, gpxb_stream stream);
/* rpcholesky_derivs (gpx/kernels/approximations.py:132-242): randomly pivoted Cholesky of the Hessian kernel
 * over the N nv (sample, variable) rows.  pivots_out: int64[M] device (row = sample * nv + variable);
 * F_out: (N nv) x M row-major device or NULL.  Same sampling stream as gpxb_rpcholesky. */
int gpxb_rpcholesky_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv,
                           double ell, uint32_t key0, uint32_t key1, int M, int partitionable, double* F_out,
                           long long* pivots_out, gpxb_stream stream);
/* SGPR trained on derivative values, CG fit (_fit_derivs_iter gpx/models/_sgpr.py:399-428 with _Ax_derivs_lhs_fun
 * :145-201): c[M nv_locs] = cg(sigma^2 H_mm + H_mn H_nm + 1e-10 I, H_mn y), H = d01kj, zero mean, no preconditioner,
 * jax cg stopping rule (tol, atol, maxiter <= 0: 10 M nv_locs).  All products are matrix-free Hessian operators.
 * Host-synchronous like gpxb_gpr_fit_iter; *iters_out is a HOST int. */
int gpxb_sgpr_fit_derivs_iter(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, int N,
                              int nf, int nv, const double* x_locs, const double* J_locs, int M, int nv_locs, double ell,
                              double sigma, double tol, double atol, int maxiter, double* c_out, int* iters_out,
                              gpxb_stream stream);
/* ---- GPR on targets and derivatives, iterative (gpx/models/gpr_td.py; one or two derivative blocks) ------------ */
/* J: N x nf x nv with nv = nv_first + nv_second: the jacobians of both derivative blocks concatenated along the
 * variable axis (nv_first == nv: one block; sigma_derivs2 is then ignored).  Vectors (c, y_derivs, U) are ordered
 * [targets (N); derivatives (sample, variable of [block 1 | block 2])] -- a symmetric permutation of the reference's
 * [targets; block 1; block 2] (the caller permutes, gpx_b200/models/gpr_td.py).
 * gpxb_gpr_td_fit_iter: c = PCG solve of
 *   [[K + s_t I, d1kj], [d0kj, d01kj + diag(s_d | s_d2)]] c = [y - mu; y_derivs],  s_* = sigma_*^2 + 1e-10
 * with the rpcholesky_TD preconditioner (_fit_iter gpr_td.py:685-747, _Ax_lhs_fun :162-310,
 * _precond_rpcholesky_TD :463-504 -- the preconditioner uses sigma_targets for every row, as the reference does; its
 * pivots are drawn over the reference's row order and the factor's rows are then permuted to the order above).
 * Matrix-free on the configuration level: N x N pair-coefficient matrices, never the (N + N nv)^2 kernel.
 * SquaredExponential and Matern-5/2.  Host-synchronous like gpxb_gpr_fit_iter; *iters_out is a HOST int. */
int gpxb_gpr_td_fit_iter(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y,
                         const double* y_derivs, int N, int nf, int nv, int nv_first, double ell, double sigma_targets,
                         double sigma_derivs, double sigma_derivs2, int mean_mode, int n_pivots, uint32_t key0,
                         uint32_t key1, int partitionable, double tol, double atol, int maxiter, double* c_out,
                         double* mu_out, int* iters_out, gpxb_stream stream);
/* gpxb_lanczos on the same block operator (the SLQ part of _lml_iter, gpr_td.py:562-637); U: (N + N nv) x ldu. */
int gpxb_lanczos_td(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, int nv_first,
                    double ell, double sigma_targets, double sigma_derivs, double sigma_derivs2, const double* U,
                    long long ldu, int P, int m, double* alpha_out, double* beta_out, int* breakdown_flag,
                    gpxb_stream stream);
/* rpcholesky_TD (gpr_td.py:313-460): randomly pivoted Cholesky of the noise-free block kernel; rows of F_out
 * ((N + N nv) x M row-major, may be NULL) and pivot indices in the REFERENCE's order [targets; block 1; block 2]. */
int gpxb_rpcholesky_td(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, int nv_first,
                       double ell, uint32_t key0, uint32_t key1, int M, int partitionable, double* F_out,
                       long long* pivots_out, gpxb_stream stream);
/* gpxb_lanczos on the matrix-free operator d01kj + (sigma^2+1e-10) I (the SLQ part of _lml_derivs_iter,
 * gpx/models/_gpr.py:873-935); U: (N nv) x P unit start vectors. */
int gpxb_lanczos_derivs(gpxb_ctx* ctx, int kind, const double* x, const double* J, int N, int nf, int nv, double ell,
                        double sigma, const double* U, long long ldu, int P, int m, double* alpha_out,
                        double* beta_out, int* breakdown_flag, gpxb_stream stream);
/* Predictions of a derivative-trained model from the uncontracted coefficients c[N nv], with the optional
 * full predictive covariance (cov_out may be NULL):
 *  - derivative values (_predict_derivs_dense, gpx/models/_gpr.py:488-560): mean[ns nvs] = mu + d01kj(x, x_train) c,
 *    cov[(ns nvs)^2] = d01kj(x, x) - G^T G, G = chol(d01kj(x_train) + (sigma^2+1e-10) I)^-1 d01kj(x_train, x);
 *  - target values (_predict_y_derivs_dense, :608-660): mean[ns] = mu + d0kj(x_train, x, J_train)^T c,
 *    cov[ns^2] = k(x, x) - G^T G with G = chol(d01kj(x_train) + sigma^2 I)^-1 d0kj(x_train, x, J_train). */
int gpxb_gpr_predict_derivs_full(gpxb_ctx* ctx, int kind, const double* x_train, const double* J_train, int N, int nf,
                                 int nv, const double* x, const double* J, int ns, int nvs, const double* c,
                                 double mu, double ell, double sigma, double* mean_out, double* cov_out,
                                 gpxb_stream stream);
int gpxb_gpr_predict_y_derivs_full(gpxb_ctx* ctx, int kind, const double* x_train, const double* J_train, int N,
                                   int nf, int nv, const double* x, int ns, const double* c, double mu, double ell,
                                   double sigma, double* mean_out, double* cov_out, gpxb_stream stream);

/* ---- GPR on targets and derivatives (GPR_TD, gpx/models/gpr_td.py) ------------------------------------ */
/* out[(n1 + n1 nv1) x (n2 + n2 nv2)] = [[k + add_targets I, d1kj], [d0kj, d01kj + add_derivs I]] between (x1, J1)
 * and (x2, J2): _A_lhs (gpx/models/gpr_td.py:43-159) with one derivative block.  Rows / columns: the n targets,
 * then the n nv derivative values (sample-major).  The shifts and lower_only need identical operands. */
int gpxb_td_gram(gpxb_ctx* ctx, int kind, const double* x1, const double* J1, int n1, int nv1, const double* x2,
                 const double* J2, int n2, int nv2, int nf, double ell, double add_targets, double add_derivs,
                 double* out, long long ldo, int lower_only, gpxb_stream stream);
/* out4 (device) = {lml, d lml/d lengthscale, d lml/d sigma_targets, d lml/d sigma_derivs} of _lml_dense
 * (gpx/models/gpr_td.py:507-560), lml normalised by m = N + N nv: the (value, grad) of the loss inside
 * scipy_minimize_ol.  y: N targets, y_derivs: N x nv.  mean_mode acts on the targets only. */
int gpxb_gpr_td_lml_grad(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y,
                         const double* y_derivs, int N, int nf, int nv, double ell, double sigma_targets,
                         double sigma_derivs, int mean_mode, int want_grad, double* out4, gpxb_stream stream);
/* c[N + N nv] = A^-1 [y - mu; y_derivs], *mu_out = mean_function(y)  (_fit_dense, gpx/models/gpr_td.py:639-683) */
int gpxb_gpr_td_fit(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, const double* y_derivs,
                    int N, int nf, int nv, double ell, double sigma_targets, double sigma_derivs, int mean_mode,
                    double* c_out, double* mu_out, gpxb_stream stream);
/* The same with the reference's optional SECOND derivative block (jacobian_2, y_derivs_2, sigma_derivs2:
 * gpx/models/gpr_td.py:114-158, 507-560, 639-683).  J (N, nf, nv) and y_derivs (N, nv) hold the two blocks
 * concatenated along the variable axis, the first nv_first variables being block 1; variables >= nv_first carry the
 * noise sigma_derivs2.  The system is the reference's 3 x 3 block matrix up to a symmetric permutation of the
 * derivative rows ((sample, variable) interleaved here), which leaves the LML unchanged; c_out is in this library's
 * row order [N targets; (sample, variable)].  out5 = {lml, d/d lengthscale, d/d sigma_targets, d/d sigma_derivs,
 * d/d sigma_derivs2}, lml normalised by m = N + N nv. */
int gpxb_gpr_td2_lml_grad(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y,
                          const double* y_derivs, int N, int nf, int nv, int nv_first, double ell, double sigma_targets,
                          double sigma_derivs, double sigma_derivs2, int mean_mode, int want_grad, double* out5,
                          gpxb_stream stream);
int gpxb_gpr_td2_fit(gpxb_ctx* ctx, int kind, const double* x, const double* J, const double* y, const double* y_derivs,
                     int N, int nf, int nv, int nv_first, double ell, double sigma_targets, double sigma_derivs,
                     double sigma_derivs2, int mean_mode, double* c_out, double* mu_out, gpxb_stream stream);

/* ---- Family 2: blocked FP64 Cholesky / TRSM / log-det / inverse ---------------------- */
/* C[M x N] = alpha * op(A) * op(B) + beta * C on the FP64 tensor pipe.  transa = 0: A is M x K
 * row-major; 1: A is stored K x M.  transb = 0: B is K x N row-major; 1: B is stored N x K.
 * Replaces the jnp.dot / @ products of the dense paths (gpx/models/_gpr.py:453, 460: K_mn^T c and
 * G^T G; gpx/models/_sgpr.py:59, 338, 454, 464, 686: K_mn K_nm, K_mn y, K_nm c, G^T G, H^T H;
 * gpx/kernels/kernels.py:226-227: the Linear kernel).  lower_only: only entries with row >= col are updated (symmetric results). */
int gpxb_gemm(gpxb_ctx* ctx, int transa, int transb, int M, int N, int K, double alpha,
              const double* A, long long lda, const double* B, long long ldb, double beta, double* C,
              long long ldc, int lower_only, gpxb_stream stream);
/* number of doubles of the diagonal-block-inverse buffer potrf fills for an n x n matrix */
size_t gpxb_potrf_inv_size(int n);
/* In-place lower Cholesky of row-major A (only the lower triangle is read / written); replaces
 * jsp.linalg.cholesky(lower=True) (gpx/models/_gpr.py:759, 457, 860).  inv (device,
 * gpxb_potrf_inv_size(n) doubles) receives the inverses of L's 128x128 diagonal blocks, which
 * gpxb_trsm / gpxb_potri consume. */
int gpxb_potrf(gpxb_ctx* ctx, double* A, long long lda, int n, double* inv, gpxb_stream stream);
/* B[m x n] <- B * L^-T (trans = 1) or B * L^-1 (trans = 0); replaces solve_triangular
 * (gpx/models/_gpr.py:760, 458): L^-1 y is the row form y^T L^-T. */
int gpxb_trsm(gpxb_ctx* ctx, const double* L, long long ldl, const double* inv, int n, double* B,
              long long ldb, int m, int trans, gpxb_stream stream);
/* *out (device) = sum_i log L[i][i]   (gpx/models/_gpr.py:763) */
int gpxb_logdet(gpxb_ctx* ctx, const double* L, long long ldl, int n, double* out,
                gpxb_stream stream);
/* out (lower triangle, n x n) = (L L^T)^-1; out may alias L.  The A^-1 of the analytic LML gradient
 * W = alpha alpha^T - A^-1, i.e. what reverse-mode autodiff through jsp.linalg.cholesky amounts to under
 * jit(value_and_grad(loss)) (gpx/optimizers/scipy_optimize.py:72-78); also the explicit jnp.linalg.inv of the
 * preconditioner (gpx/models/_gpr.py:205, 245). */
int gpxb_potri(gpxb_ctx* ctx, const double* L, long long ldl, const double* inv, int n, double* out,
               long long ldo, gpxb_stream stream);

/* Exact-GPR log marginal likelihood and its gradient, fused on device.
 * out[0] = lml / N (gpx/models/_gpr.py:737-769), and if want_grad:
 * out[1] = d(lml/N)/d lengthscale, out[2] = d(lml/N)/d sigma -- the constrained-space gradient
 * jit(value_and_grad) produces before the bijector chain rule
 * (gpx/optimizers/scipy_optimize.py:72-78).  y is N x T row-major. */
int gpxb_gpr_lml_grad(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T,
                      double ell, double sigma, int mean_mode, int want_grad, double* out3,
                      gpxb_stream stream);
/* c[N x T] = (K + (sigma^2+1e-10) I)^-1 (y - mu), *mu_out = mean_function(y)
 * (gpx/models/_gpr.py:262-282; the reference's LU solve is replaced by the Cholesky solve). */
int gpxb_gpr_fit(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T,
                 double ell, double sigma, int mean_mode, double* c_out, double* mu_out,
                 gpxb_stream stream);
/* mean_out[ns x T] = mu + K(x, x_train) c; if cov_out != NULL also the full predictive
 * covariance [ns x ns] (gpx/models/_gpr.py:434-463).  mu is a device scalar. */
int gpxb_gpr_predict(gpxb_ctx* ctx, int kind, const double* x_train, int N, int D, const double* x,
                     int ns, const double* c, int T, const double* mu, double ell, double sigma,
                     double* mean_out, double* cov_out, gpxb_stream stream);

/* ---- Family 3: matrix-free mat-vec, Lanczos / SLQ, PCG, randomly pivoted Cholesky ------ */
/* W[n_rows x P] = (K(x_rows, x_all) + diag_add * I) V, K never stored.  row_offset is the index of
 * x_rows[0] inside x_all (the shifted diagonal sits at column row_offset + i), or -1 for a plain
 * cross-kernel product.  Replaces rowfun_to_matvec + _Ax_lhs_fun + update_row_diagonal
 * (gpx/operations.py:32-123, gpx/models/_gpr.py:94-119). */
int gpxb_kmatvec(gpxb_ctx* ctx, int kind, const double* x_rows, int n_rows, long long row_offset,
                 const double* x_all, int N, int D, double ell, double diag_add, const double* V,
                 long long ldv, int P, double* W, long long ldw, gpxb_stream stream);
/* Randomly pivoted Cholesky (gpx/kernels/approximations.py:30-129).  key = (key0, key1) as
 * jax.random.PRNGKey words; partitionable = jax_threefry_partitionable of the JAX being matched.
 * pivots_out: int64[M] device.  F_out: N x M row-major device, or NULL when only pivots are wanted. */
int gpxb_rpcholesky(gpxb_ctx* ctx, int kind, const double* x, int N, int D, double ell, uint32_t key0,
                    uint32_t key1, int M, int partitionable, double* F_out, long long* pivots_out,
                    gpxb_stream stream);
/* U_out[N x P] = jax.random.rademacher(key, (N, P)) / sqrt(N)   (gpx/lanczos.py:168-170) */
int gpxb_rademacher_probes(gpxb_ctx* ctx, uint32_t key0, uint32_t key1, int N, int P, int partitionable,
                           double* U_out, gpxb_stream stream);
/* Breakdown branch of lanczos_tridiagonal (gpx/lanczos.py:92-108: when beta <= 1e-12 the next Lanczos vector is
 * orthonormalize(random.normal(subkey, (N,)), vecs), subkey drawn from the key carried through the loop).  While
 * enabled, gpxb_lanczos and gpxb_gpr_lml_iter read the breakdown flag back once at the end of the (enqueue-only)
 * recursion and, in the rare case that it is set, redo the recursion with that branch (one host read of beta per
 * step) and clear the flag.  key = the key lanczos_tridiagonal receives (the second half of the split in
 * lanczos_logdet, :166).  Disabled (default): a breakdown only sets the flag.  The tangent and derivative-kernel
 * variants (gpxb_lanczos_grad, gpxb_lanczos_derivs, ...) always only flag. */
int gpxb_lanczos_set_restart(gpxb_ctx* ctx, int enabled, uint32_t key0, uint32_t key1, int partitionable);
/* m-step Lanczos with full MGS re-orthogonalisation on A = K(x,x) + diag_add I for the P unit
 * start vectors in the columns of U (gpx/lanczos.py:46-133, all probes advanced together).
 * alpha_out[P x m]: diagonals; beta_out[P x m]: beta_out[p][j] = T_p[j][j+1] for j < m-1.
 * *breakdown_flag (device int) is set when some beta <= 1e-12 (the reference would restart with a
 * random vector; this build reports it instead). */
int gpxb_lanczos(gpxb_ctx* ctx, int kind, const double* x, int N, int D, double ell, double diag_add,
                 const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                 int* breakdown_flag, gpxb_stream stream);
/* gpxb_lanczos on A = K(x,x; ell) + (sigma^2+1e-10) I together with its forward-mode derivative w.r.t.
 * (lengthscale, sigma): dalpha_out / dbeta_out [2 x P x m] (index 0: lengthscale, 1: sigma) are the tangents
 * of the tridiagonals -- the derivative jit(value_and_grad(loss)) propagates through lanczos_tridiagonal's
 * fori_loop (gpx/lanczos.py:46-133) when the iterative LML is trained. */
int gpxb_lanczos_grad(gpxb_ctx* ctx, int kind, const double* x, int N, int D, double ell, double sigma,
                      const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                      double* dalpha_out, double* dbeta_out, int* breakdown_flag, gpxb_stream stream);
/* W[N x P] = (d K(x, x)/d lengthscale) V, matrix-free (the cotangent contraction c^T dA c of the CG solve,
 * jax.lax.custom_linear_solve under value_and_grad of gpx/models/_gpr.py:772-821). */
int gpxb_kmatvec_dl(gpxb_ctx* ctx, int kind, const double* x, int N, int D, double ell, const double* V,
                    long long ldv, int P, double* W, long long ldw, gpxb_stream stream);
/* c[N] = PCG solve of (K + (sigma^2+1e-10) I) c = y - mu with the RPCholesky preconditioner
 * (gpx/models/_gpr.py:285-310, 180-216; jax.scipy.sparse.linalg.cg semantics: x0 = 0, stop when
 * r.r <= max(tol^2 b.b, atol^2) or after maxiter (<= 0: 10 N) iterations).  n_pivots = 0: no
 * preconditioner.  Host-synchronous (one scalar read per iteration); *iters_out is a HOST int. */
int gpxb_gpr_fit_iter(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, double ell,
                      double sigma, int mean_mode, int n_pivots, uint32_t key0, uint32_t key1,
                      int partitionable, double tol, double atol, int maxiter, double* c_out, double* mu_out,
                      int* iters_out, gpxb_stream stream);

/* The two device parts of the iterative LML in ONE sweep (gpx/models/_gpr.py:772-821 _lml_iter): the PCG solve of
 * gpxb_gpr_fit_iter and the m-step Lanczos of gpxb_lanczos on the P start vectors U with diag_add = sigma^2+1e-10.
 * While both run, the PCG search direction travels as column P of the Lanczos block, so the first m CG iterations
 * share the block mat-vec's pass over the recomputed kernel tiles (P <= 31; larger P runs the two parts in turn).
 * Outputs as of the two separate calls.  Host-synchronous like gpxb_gpr_fit_iter; *iters_out is a HOST int. */
int gpxb_gpr_lml_iter(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, double ell,
                      double sigma, int mean_mode, int n_pivots, uint32_t key0, uint32_t key1, int partitionable,
                      double tol, double atol, int maxiter, const double* U, long long ldu, int P, int m,
                      double* c_out, double* mu_out, int* iters_out, double* alpha_out, double* beta_out,
                      int* breakdown_flag, gpxb_stream stream);

/* ---- SGPR, Projected Processes --------------------------------------------------------- */
/* *out (device) = lml / N of gpx/models/_sgpr.py:659-697, evaluated in the Woodbury form. */
int gpxb_sgpr_lml(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T,
                  const double* x_locs, int M, double ell, double sigma, int mean_mode, double* out,
                  gpxb_stream stream);
/* out3 (device) = [lml/N, d(lml/N)/d lengthscale, d(lml/N)/d sigma] of the SGPR LML with fixed
 * landmarks: the (value, grad) jit(value_and_grad) of gpx/models/_sgpr.py:659-697 yields, in analytic
 * O(N M^2) form (SGPR_RPChol: the pivots carry no gradient).  Single output column. */
int gpxb_sgpr_lml_grad(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D,
                       const double* x_locs, int M, double ell, double sigma, int mean_mode, int want_grad,
                       double* out3, gpxb_stream stream);
/* As gpxb_sgpr_lml_grad (with the gradient), plus grad_locs[M x D] = d(lml/N)/d x_locs: the entry of
 * jit(value_and_grad) for a trainable `x_locs` Parameter of SGPR (gpx/models/sgpr.py:690-732).  Evaluated
 * as two weighted products of the kernel's radial derivative profile with [1 | X], matrix-free. */
int gpxb_sgpr_lml_grad_locs(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D,
                            const double* x_locs, int M, double ell, double sigma, int mean_mode, double* out3,
                            double* grad_locs, gpxb_stream stream);
/* c[M x T] = (sigma^2 K_mm + K_mn K_nm + 1e-10 I)^-1 K_mn (y - mu)  (gpx/models/_sgpr.py:39-61, 319-339) */
int gpxb_sgpr_fit(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, int T,
                  const double* x_locs, int M, double ell, double sigma, int mean_mode, double* c_out,
                  double* mu_out, gpxb_stream stream);
/* Iterative SGPR (single GPU).  gpxb_sgpr_fit_iter: c[M] = cg(sigma^2 K_mm + K_mn K_nm + 1e-10 I, K_mn (y - mu)),
 * _Ax_lhs_fun / _fit_iter (gpx/models/_sgpr.py:96-143, 343-363), every product a matrix-free block mat-vec.
 * gpxb_sgpr_lml_iter_parts: the two device parts of _lml_iter (:701-737) on the Nystrom operator
 * H + (sigma^2+1e-10) I, H = K_nm (K_mm + 1e-10 I)^-1 K_mn (_Hx_fun :204-239): *quad_out = r^T cg(.., r) and the
 * m-step Lanczos tridiagonals for the P start vectors U.  tol / atol / maxiter as jax.scipy.sparse.linalg.cg. */
int gpxb_sgpr_fit_iter(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D, const double* x_locs,
                       int M, double ell, double sigma, int mean_mode, double tol, double atol, int maxiter,
                       double* c_out, double* mu_out, int* iters_out, gpxb_stream stream);
int gpxb_sgpr_lml_iter_parts(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D,
                             const double* x_locs, int M, double ell, double sigma, int mean_mode, double tol,
                             double atol, int maxiter, double* quad_out, int* iters_out, const double* U,
                             long long ldu, int P, int m, double* alpha_out, double* beta_out, int* breakdown_flag,
                             gpxb_stream stream);
/* gpxb_sgpr_lml_iter_parts plus what jit(value_and_grad) of _lml_iter (gpx/models/_sgpr.py:700-737) needs for fixed
 * landmarks: gquad_out[3] (device) = { (dK_mn c).g, g.(dK_mm g), c.c } with c the CG solution and g = K_mm^-1 K_mn c
 * (c^T dA/d ell c = 2 gquad[0] - gquad[1]; c^T dA/d sigma c = 2 sigma gquad[2]; d(r.c) = -c^T dA c because the CG
 * solve is a lax.custom_linear_solve), and dalpha_out / dbeta_out [2][P][m]: the tangents of the Lanczos
 * tridiagonals w.r.t. (lengthscale, sigma) (forward mode through gpx/lanczos.py:46-133).  D <= 64. */
int gpxb_sgpr_lml_iter_grad_parts(gpxb_ctx* ctx, int kind, const double* x, const double* y, int N, int D,
                                  const double* x_locs, int M, double ell, double sigma, int mean_mode, double tol,
                                  double atol, int maxiter, double* quad_out, double* gquad_out, int* iters_out,
                                  const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                                  double* dalpha_out, double* dbeta_out, int* breakdown_flag, gpxb_stream stream);
/* gpx/models/_sgpr.py:434-467 (covariance built from the test points, as the reference does) */
int gpxb_sgpr_predict(gpxb_ctx* ctx, int kind, const double* x_locs, int M, int D, const double* x, int ns,
                      const double* c, int T, const double* mu, double ell, double sigma, double* mean_out,
                      double* cov_out, gpxb_stream stream);
/* out[k][:] = x[idx[k]][:]  (x_locs = x[pivots], gpx/models/_sgpr_rpchol.py:50, 200) */
int gpxb_gather_rows(gpxb_ctx* ctx, const double* x, int D, const long long* idx, int M, double* out,
                     gpxb_stream stream);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) --------------------------------- */
/* No counterpart in the reference (GPX is single-process; SURVEY.md 8e defines the sharding).  Arrays are
 * passed whole on every rank; each rank works on its own row range of x.  Rows are sharded for gpxb_lanczos,
 * gpxb_gpr_fit_iter, gpxb_sgpr_lml and gpxb_sgpr_fit; the collectives are the all-reduces of the column dot
 * products (gpx/lanczos.py:29-37, 84-118), of the CDF block sums (gpx/kernels/approximations.py:95-105) and of
 * B = G G^T (the Woodbury form of gpx/models/_sgpr.py:684-688). */
int gpxb_comm_unique_id(void* out128);
/* host-only: the row range [begin, end) rank owns out of N rows (1024-row aligned partition: the block size of the RPCholesky CDF sums, so pivots do not depend on the GPU count) */
int gpxb_shard_range(long long N, int world, int rank, long long* begin, long long* end);
int gpxb_comm_init(gpxb_ctx* ctx, const void* id128, int rank, int world);

#ifdef __cplusplus
}
#endif
#endif /* GPXB200_H */
