// C-ABI surface: library/context management and the Family 1/2 primitives.
#include "../../include/gpxb200.h"

#include <cstdarg>

#include "chol.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "small_kernels.cuh"

namespace gpxb {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

}  // namespace gpxb

using gpxb::Ctx;
using gpxb::DevBuf;

extern "C" {

int gpxb_version(void) { return 100; }

int gpxb_last_error(char* buf, size_t buflen) {
  if (!buf || buflen == 0) return -1;
  snprintf(buf, buflen, "%s", gpxb::get_error());
  return 0;
}

int gpxb_ctx_create(int device, gpxb_ctx** out) {
  GPXB_ARG_CHECK(out != nullptr, -1);
  int ndev = 0;
  GPXB_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  GPXB_ARG_CHECK(device >= 0 && device < ndev, -2);
  GPXB_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GPXB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    gpxb::set_error("libgpxb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
                    prop.minor);
    return -3;
  }
  Ctx* c = new Ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  int prio_lo = 0, prio_hi = 0;
  GPXB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  GPXB_CUDA_CHECK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_hi));
  GPXB_CUDA_CHECK(cudaStreamCreateWithPriority(&c->crit, cudaStreamNonBlocking, prio_hi));
  for (int i = 0; i < 3; ++i) GPXB_CUDA_CHECK(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
  GPXB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  GPXB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  // keep stream-ordered scratch cached between calls (multi-GB workspaces are reused)
  cudaMemPool_t pool;
  GPXB_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long thresh = ~0ULL;
  GPXB_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  *out = reinterpret_cast<gpxb_ctx*>(c);
  return 0;
}

int gpxb_ctx_destroy(gpxb_ctx* ctx) {
  if (!ctx) return 0;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->crit) cudaStreamDestroy(c->crit);
  if (c->head_counter) cudaFree(c->head_counter);
  for (int i = 0; i < 3; ++i)
    if (c->aux[i]) cudaStreamDestroy(c->aux[i]);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  delete c;
  return 0;
}

int gpxb_prof_enable(gpxb_ctx* ctx, int on) {
  GPXB_ARG_CHECK(ctx != nullptr, -1);
  reinterpret_cast<Ctx*>(ctx)->prof = (on != 0);
  reinterpret_cast<Ctx*>(ctx)->serialize = (on == 2);
  return 0;
}

int gpxb_prof_collect(gpxb_ctx* ctx_, double* total_ms, double* total_flops, int* n_launches) {
  GPXB_ARG_CHECK(ctx_ && total_ms && total_flops && n_launches, -1);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  GPXB_CUDA_CHECK(cudaDeviceSynchronize());
  double ms = 0.0, fl = 0.0;
  for (auto& r : ctx->prof_recs) {
    float t = 0.f;
    GPXB_CUDA_CHECK(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t;
    fl += r.flops;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  *total_ms = ms;
  *total_flops = fl;
  *n_launches = (int)ctx->prof_recs.size();
  ctx->prof_recs.clear();
  return 0;
}

int gpxb_phase_enable(gpxb_ctx* ctx, int on) {
  GPXB_ARG_CHECK(ctx != nullptr, -1);
  reinterpret_cast<Ctx*>(ctx)->phase_prof = (on != 0);
  return 0;
}

int gpxb_phase_collect(gpxb_ctx* ctx_, double* ms_out, int* count_out, int n) {
  GPXB_ARG_CHECK(ctx_ && ms_out && count_out && n >= 1, -1);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  GPXB_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < n; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
  for (auto& r : ctx->phase_recs) {
    float t = 0.f;
    if (r.id >= 0 && r.id < n && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      ms_out[r.id] += t * r.weight;
      count_out[r.id]++;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->phase_recs.clear();
  (void)cudaGetLastError();
  return 0;
}

unsigned long long gpxb_ctx_launches(gpxb_ctx* ctx) {
  return ctx ? reinterpret_cast<Ctx*>(ctx)->launches : 0ULL;
}

int gpxb_gram(gpxb_ctx* ctx_, int kind, const double* x1, int n1, const double* x2, int n2, int D,
              double ell, double diag_add, double* out, long long ldo, int lower_only,
              gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x1 && x2 && out, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && D > 0 && n1 >= 0 && n2 >= 0, -2);
  if (n1 == 0 || n2 == 0) return 0;
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const gpxb::ZLayout zl = gpxb::z_layout(D);
  const bool sym = (x1 == x2) && (n1 == n2);
  GPXB_ARG_CHECK(!lower_only || sym, -4);
  DevBuf z1, z2;
  GPXB_TRY(z1.alloc(sizeof(double) * (size_t)n1 * zl.S, s));
  GPXB_TRY(gpxb::prep_z(ctx, x1, D, n1, zl, ell, z1.as<double>(), s));
  const double* Z2 = z1.as<double>();
  if (!sym) {
    GPXB_TRY(z2.alloc(sizeof(double) * (size_t)n2 * zl.S, s));
    GPXB_TRY(gpxb::prep_z(ctx, x2, D, n2, zl, ell, z2.as<double>(), s));
    Z2 = z2.as<double>();
  }
  gpxb::GramArgs g;
  g.kind = kind;
  g.Z1 = z1.as<double>(); g.n1 = n1; g.Z2 = Z2; g.n2 = n2; g.zl = zl;
  g.ell = ell; g.diag_add = diag_add; g.sym = sym ? 1 : 0; g.lower_only = lower_only;
  g.out = out; g.ldo = ldo;
  return gpxb::gram_launch(ctx, g, s);
}

int gpxb_gram_dl_contract(gpxb_ctx* ctx_, int kind, const double* x1, int n1, const double* x2, int n2, int D,
                          double ell, const double* Y, long long ldy, int sym_lower, double* out,
                          gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x1 && x2 && Y && out, -1);
  GPXB_ARG_CHECK(kind >= 0 && kind <= 3 && D > 0 && n1 > 0 && n2 > 0, -2);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  const gpxb::ZLayout zl = gpxb::z_layout(D);
  const bool sym = (x1 == x2) && (n1 == n2);
  GPXB_ARG_CHECK(!sym_lower || sym, -4);
  DevBuf z1, z2, part;
  GPXB_TRY(z1.alloc(sizeof(double) * (size_t)n1 * zl.S, s));
  GPXB_TRY(gpxb::prep_z(ctx, x1, D, n1, zl, ell, z1.as<double>(), s));
  const double* Z2 = z1.as<double>();
  if (!sym) {
    GPXB_TRY(z2.alloc(sizeof(double) * (size_t)n2 * zl.S, s));
    GPXB_TRY(gpxb::prep_z(ctx, x2, D, n2, zl, ell, z2.as<double>(), s));
    Z2 = z2.as<double>();
  }
  gpxb::GramArgs g;
  g.kind = kind;
  g.Z1 = z1.as<double>(); g.n1 = n1; g.Z2 = Z2; g.n2 = n2; g.zl = zl;
  g.ell = ell; g.sym = sym ? 1 : 0; g.lower_only = sym_lower;
  g.Ainv = Y; g.ldainv = ldy; g.t = 0;  // weights -Y: the sign is undone below
  const long long nparts = gpxb::gram_num_ctas(g);
  GPXB_TRY(part.alloc(sizeof(double) * (size_t)nparts, s));
  g.partials = part.as<double>();
  GPXB_TRY(gpxb::gram_launch(ctx, g, s));
  gpxb::sk::sum_kernel<<<1, 1024, 0, s>>>(g.partials, nparts, out);
  GPXB_LAUNCHED(ctx);
  gpxb::sk::scale_kernel<<<1, 1, 0, s>>>(out, 1, -1.0);
  GPXB_LAUNCHED(ctx);
  return 0;
}

int gpxb_elementwise(gpxb_ctx* ctx_, int op, long long n, double a, const double* X, double b, const double* Y,
                     double* out, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && X && Y && out && n >= 0 && (op == 0 || op == 1), -1);
  if (n == 0) return 0;
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  if (op == 0) gpxb::sk::axpby_kernel<<<gpxb::ceil_div(n, 256), 256, 0, s>>>(out, a, X, b, Y, n);
  else gpxb::sk::hadamard_kernel<<<gpxb::ceil_div(n, 256), 256, 0, s>>>(out, a, X, Y, n);
  GPXB_LAUNCHED(ctx);
  return 0;
}

int gpxb_gemm(gpxb_ctx* ctx, int transa, int transb, int M, int N, int K, double alpha,
              const double* A, long long lda, const double* B, long long ldb, double beta, double* C,
              long long ldc, int lower_only, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx != nullptr, -1);
  gpxb::GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_kmajor = (transa == 0);
  g.B = B; g.ldb = ldb; g.b_kmajor = (transb != 0);
  g.C = C; g.ldc = ldc;
  g.alpha = alpha; g.beta = beta; g.lower_only = lower_only;
  return gpxb::gemm_f64(reinterpret_cast<Ctx*>(ctx), g, (cudaStream_t)stream);
}

size_t gpxb_potrf_inv_size(int n) { return gpxb::leaf_inv_doubles(n); }

int gpxb_potrf(gpxb_ctx* ctx, double* A, long long lda, int n, double* inv, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && A && inv && n >= 0 && lda >= n, -1);
  return gpxb::potrf_lower(reinterpret_cast<Ctx*>(ctx), A, lda, n, inv, (cudaStream_t)stream);
}

int gpxb_trsm(gpxb_ctx* ctx, const double* L, long long ldl, const double* inv, int n, double* B,
              long long ldb, int m, int trans, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && L && inv && B, -1);
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  return trans ? gpxb::trsm_right_lt(c, L, ldl, inv, B, ldb, m, n, (cudaStream_t)stream)
               : gpxb::trsm_right_l(c, L, ldl, inv, B, ldb, m, n, (cudaStream_t)stream);
}

int gpxb_logdet(gpxb_ctx* ctx, const double* L, long long ldl, int n, double* out,
                gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx && L && out, -1);
  return gpxb::logdet_lower(reinterpret_cast<Ctx*>(ctx), L, ldl, n, out, 1.0, (cudaStream_t)stream);
}

int gpxb_potri(gpxb_ctx* ctx_, const double* L, long long ldl, const double* inv, int n, double* out,
               long long ldo, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && L && inv && out && n > 0, -1);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf w, t;
  GPXB_TRY(w.alloc(sizeof(double) * (size_t)n * n, s));
  GPXB_TRY(t.alloc(sizeof(double) * gpxb::potri_scratch_doubles(n), s));
  return gpxb::potri_lower(ctx, L, ldl, inv, n, w.as<double>(), t.as<double>(), out, ldo, s);
}

}  // extern "C"
