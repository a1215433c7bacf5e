// XLA-FFI (jax.ffi) shim over the C ABI of libgpxb200 -- the custom-call layer BASELINE.json's
// north star names.  NOT built in this repository: JAX / jaxlib (and xla/ffi/api/ffi.h) are absent
// from the image.  Where jaxlib exists:
//
//   g++ -std=c++17 -shared -fPIC ffi_shim.cc -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
//       -I../../../include -L../.. -lgpxb200 -lcudart -o libgpxb200_ffi.so        (one command line)
//
// tests/test_abi_cpu.py type-checks this file against include/gpxb200.h with a MOCK of the few XLA-FFI names it
// uses (tests/c/mock_xla), so the handler bodies cannot drift from the C ABI unnoticed.
//
// and on the Python side:
//   jax.ffi.register_ffi_target("gpxb_gpr_lml_grad", jax.ffi.pycapsule(lib.GpxbGprLmlGrad), platform="CUDA")
//
// Contract kept by every handler: buffers are XLA-owned device pointers (row-major, outputs
// pre-allocated), the call only enqueues on the stream XLA passes, never throws, and reports
// failures as ffi::Error with the library's message.  Numerical failure (non-PD matrix) is not an
// error: NaNs propagate, as with jax.scipy.linalg.cholesky.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include <mutex>

#include "gpxb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {
constexpr int kMaxDevices = 64;
gpxb_ctx* g_ctx[kMaxDevices] = {nullptr};
std::mutex g_mu;

gpxb_ctx* ctx_for_current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_mu);
  if (dev < 0 || dev >= kMaxDevices) return nullptr;
  if (!g_ctx[dev]) gpxb_ctx_create(dev, &g_ctx[dev]);
  return g_ctx[dev];
}

ffi::Error status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  char msg[1024];
  gpxb_last_error(msg, sizeof msg);
  return ffi::Error::Internal(msg);
}
}  // namespace

// out3 = [lml/N, d/d ell, d/d sigma]  (gpx/models/_gpr.py:737-769 + value_and_grad)
static ffi::Error GprLmlGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y, double ell,
                                 double sigma, int32_t kind, int32_t mean_mode, int32_t want_grad,
                                 ffi::ResultBuffer<ffi::F64> out3) {
  const auto dx = x.dimensions();
  const auto dy = y.dimensions();
  const int T = dy.size() > 1 ? (int)dy[1] : 1;
  return status(gpxb_gpr_lml_grad(ctx_for_current_device(), kind, x.typed_data(), y.typed_data(), (int)dx[0],
                                  (int)dx[1], T, ell, sigma, mean_mode, want_grad, out3->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbGprLmlGrad, GprLmlGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("sigma")
                                  .Attr<int32_t>("kind")
                                  .Attr<int32_t>("mean_mode")
                                  .Attr<int32_t>("want_grad")
                                  .Ret<ffi::Buffer<ffi::F64>>());

// K = k(x1, x2) + diag_add I  (Kernel.k; gpx/kernels/kernels.py:760-769)
static ffi::Error GramImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x1, ffi::Buffer<ffi::F64> x2, double ell,
                           double diag_add, int32_t kind, ffi::ResultBuffer<ffi::F64> out) {
  const auto d1 = x1.dimensions();
  const auto d2 = x2.dimensions();
  return status(gpxb_gram(ctx_for_current_device(), kind, x1.typed_data(), (int)d1[0], x2.typed_data(), (int)d2[0],
                          (int)d1[1], ell, diag_add, out->typed_data(), d2[0], 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbGram, GramImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("diag_add")
                                  .Attr<int32_t>("kind")
                                  .Ret<ffi::Buffer<ffi::F64>>());

// W = (K(x, x) + diag_add I) V without storing K  (rowfun_to_matvec; gpx/operations.py:54-123)
static ffi::Error KmatvecImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> v, double ell,
                              double diag_add, int32_t kind, ffi::ResultBuffer<ffi::F64> w) {
  const auto dx = x.dimensions();
  const auto dv = v.dimensions();
  const int P = dv.size() > 1 ? (int)dv[1] : 1;
  return status(gpxb_kmatvec(ctx_for_current_device(), kind, x.typed_data(), (int)dx[0], 0, x.typed_data(),
                             (int)dx[0], (int)dx[1], ell, diag_add, v.typed_data(), P, P, w->typed_data(), P, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbKmatvec, KmatvecImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("diag_add")
                                  .Attr<int32_t>("kind")
                                  .Ret<ffi::Buffer<ffi::F64>>());

// pivots (int64[M]) of rpcholesky(key, x, M, kernel, params)  (gpx/kernels/approximations.py:30-129)
static ffi::Error RpcholeskyImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::U32> key, double ell,
                                 int32_t kind, int32_t partitionable, ffi::ResultBuffer<ffi::S64> pivots) {
  const auto dx = x.dimensions();
  uint32_t k[2];
  cudaMemcpyAsync(k, key.typed_data(), sizeof k, cudaMemcpyDeviceToHost, stream);
  cudaStreamSynchronize(stream);  // the two key words are a host-side scalar of the C ABI
  return status(gpxb_rpcholesky(ctx_for_current_device(), kind, x.typed_data(), (int)dx[0], (int)dx[1], ell, k[0],
                                k[1], (int)pivots->dimensions()[0], partitionable, nullptr,
                                reinterpret_cast<long long*>(pivots->typed_data()), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbRpcholesky, RpcholeskyImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Attr<double>("ell")
                                  .Attr<int32_t>("kind")
                                  .Attr<int32_t>("partitionable")
                                  .Ret<ffi::Buffer<ffi::S64>>());

// (c (N, T), mu (1,)) = _fit_dense  (gpx/models/_gpr.py:262-282)
static ffi::Error GprFitImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y, double ell,
                             double sigma, int32_t kind, int32_t mean_mode, ffi::ResultBuffer<ffi::F64> c,
                             ffi::ResultBuffer<ffi::F64> mu) {
  const auto dx = x.dimensions();
  const auto dy = y.dimensions();
  const int T = dy.size() > 1 ? (int)dy[1] : 1;
  return status(gpxb_gpr_fit(ctx_for_current_device(), kind, x.typed_data(), y.typed_data(), (int)dx[0], (int)dx[1], T,
                             ell, sigma, mean_mode, c->typed_data(), mu->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbGprFit, GprFitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("sigma")
                                  .Attr<int32_t>("kind")
                                  .Attr<int32_t>("mean_mode")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// mean (ns, T) and the full covariance (ns, ns) of _predict_dense  (gpx/models/_gpr.py:434-463); mu is a device scalar
static ffi::Error GprPredictImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x_train, ffi::Buffer<ffi::F64> x,
                                 ffi::Buffer<ffi::F64> c, ffi::Buffer<ffi::F64> mu, double ell, double sigma,
                                 int32_t kind, ffi::ResultBuffer<ffi::F64> mean, ffi::ResultBuffer<ffi::F64> cov) {
  const auto dt = x_train.dimensions();
  const auto dc = c.dimensions();
  const int T = dc.size() > 1 ? (int)dc[1] : 1;
  return status(gpxb_gpr_predict(ctx_for_current_device(), kind, x_train.typed_data(), (int)dt[0], (int)dt[1],
                                 x.typed_data(), (int)x.dimensions()[0], c.typed_data(), T, mu.typed_data(), ell, sigma,
                                 mean->typed_data(), cov->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbGprPredict, GprPredictImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("sigma")
                                  .Attr<int32_t>("kind")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// out3 = [lml/N, d/d ell, d/d sigma] of the SGPR LML with fixed landmarks  (gpx/models/_sgpr.py:659-697 + value_and_grad)
static ffi::Error SgprLmlGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y,
                                  ffi::Buffer<ffi::F64> x_locs, double ell, double sigma, int32_t kind,
                                  int32_t mean_mode, int32_t want_grad, ffi::ResultBuffer<ffi::F64> out3) {
  const auto dx = x.dimensions();
  return status(gpxb_sgpr_lml_grad(ctx_for_current_device(), kind, x.typed_data(), y.typed_data(), (int)dx[0],
                                   (int)dx[1], x_locs.typed_data(), (int)x_locs.dimensions()[0], ell, sigma, mean_mode,
                                   want_grad, out3->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbSgprLmlGrad, SgprLmlGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("sigma")
                                  .Attr<int32_t>("kind")
                                  .Attr<int32_t>("mean_mode")
                                  .Attr<int32_t>("want_grad")
                                  .Ret<ffi::Buffer<ffi::F64>>());

// (c (M, T), mu (1,)) = _sgpr._fit_dense  (gpx/models/_sgpr.py:319-339)
static ffi::Error SgprFitImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y,
                              ffi::Buffer<ffi::F64> x_locs, double ell, double sigma, int32_t kind, int32_t mean_mode,
                              ffi::ResultBuffer<ffi::F64> c, ffi::ResultBuffer<ffi::F64> mu) {
  const auto dx = x.dimensions();
  const auto dy = y.dimensions();
  const int T = dy.size() > 1 ? (int)dy[1] : 1;
  return status(gpxb_sgpr_fit(ctx_for_current_device(), kind, x.typed_data(), y.typed_data(), (int)dx[0], (int)dx[1], T,
                              x_locs.typed_data(), (int)x_locs.dimensions()[0], ell, sigma, mean_mode, c->typed_data(),
                              mu->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbSgprFit, SgprFitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("sigma")
                                  .Attr<int32_t>("kind")
                                  .Attr<int32_t>("mean_mode")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// mean (ns, T) and full covariance (ns, ns) of _sgpr._predict_dense  (gpx/models/_sgpr.py:434-467)
static ffi::Error SgprPredictImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x_locs, ffi::Buffer<ffi::F64> x,
                                  ffi::Buffer<ffi::F64> c, ffi::Buffer<ffi::F64> mu, double ell, double sigma,
                                  int32_t kind, ffi::ResultBuffer<ffi::F64> mean, ffi::ResultBuffer<ffi::F64> cov) {
  const auto dl = x_locs.dimensions();
  const auto dc = c.dimensions();
  const int T = dc.size() > 1 ? (int)dc[1] : 1;
  return status(gpxb_sgpr_predict(ctx_for_current_device(), kind, x_locs.typed_data(), (int)dl[0], (int)dl[1],
                                  x.typed_data(), (int)x.dimensions()[0], c.typed_data(), T, mu.typed_data(), ell, sigma,
                                  mean->typed_data(), cov->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbSgprPredict, SgprPredictImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("sigma")
                                  .Attr<int32_t>("kind")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// out3 = [lml, d/d ell, d/d sigma] of the derivative-trained GPR  (gpx/models/_gpr.py:824-870 + value_and_grad);
// x (N, nf), jacobian (N, nf, nv), y (N, nv)
static ffi::Error GprLmlGradDerivsImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> jac,
                                       ffi::Buffer<ffi::F64> y, double ell, double sigma, int32_t kind,
                                       int32_t want_grad, ffi::ResultBuffer<ffi::F64> out3) {
  const auto dj = jac.dimensions();
  return status(gpxb_gpr_lml_grad_derivs(ctx_for_current_device(), kind, x.typed_data(), jac.typed_data(),
                                         y.typed_data(), (int)dj[0], (int)dj[1], (int)dj[2], ell, sigma, want_grad,
                                         out3->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpxbGprLmlGradDerivs, GprLmlGradDerivsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("ell")
                                  .Attr<double>("sigma")
                                  .Attr<int32_t>("kind")
                                  .Attr<int32_t>("want_grad")
                                  .Ret<ffi::Buffer<ffi::F64>>());
#else
// jaxlib headers not present: nothing to build (see the header comment).
#endif
