// Small fixed-order reduction / element-wise kernels shared by the model-level pipelines.
#pragma once
#include "common.cuh"
#include "mma.cuh"

namespace gpxb {
namespace sk {

// mean over all n elements (data_mean, gpx/mean_functions.py:29-30), or 0 (zero_mean)
static __global__ void mean_kernel(const double* __restrict__ y, long long n, int data_mean,
                                   double* __restrict__ mu) {
  __shared__ double sh[32];
  double s = 0.0;
  if (data_mean)
    for (long long i = threadIdx.x; i < n; i += 1024) s += y[i];
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *mu = data_mean ? s / (double)n : 0.0;
}

// rT[t][i] = y[i][t] - mu
static __global__ void center_transpose_kernel(const double* __restrict__ y, int N, int T,
                                               const double* __restrict__ mu, double* __restrict__ rT) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * T) return;
  const int i = (int)(idx / T), t = (int)(idx % T);
  rT[(long long)t * N + i] = y[idx] - *mu;
}

// r = y - mu (same layout)
static __global__ void center_kernel(const double* __restrict__ y, long long n, const double* __restrict__ mu,
                                     double* __restrict__ r) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) r[idx] = y[idx] - *mu;
}

// out[i][t] = in[t][i]
static __global__ void transpose_out_kernel(const double* __restrict__ in, int N, int T,
                                            double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * T) return;
  const int i = (int)(idx / T), t = (int)(idx % T);
  out[idx] = in[(long long)t * N + i];
}

// *out (+)= sum v[i]^2   (single CTA, fixed order)
static __global__ void sumsq_kernel(const double* __restrict__ v, long long n, double* __restrict__ out,
                                    int accumulate) {
  __shared__ double sh[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) s += v[i] * v[i];
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = accumulate ? *out + s : s;
}

static __global__ void sum_kernel(const double* __restrict__ v, long long n, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) s += v[i];
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = s;
}

static __global__ void diag_sum_kernel(const double* __restrict__ A, long long ld, int n,
                                       double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += A[(long long)i * ld + i];
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = s;
}

static __global__ void add_diag_kernel(double* __restrict__ A, long long ld, int n, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(long long)i * ld + i] += v;
}

static __global__ void add_scalar_kernel(double* __restrict__ v, long long n, const double* __restrict__ mu) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) v[idx] += *mu;
}

// c += a
static __global__ void add_vec_kernel(double* __restrict__ c, const double* __restrict__ a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] += a[i];
}

// x_out[k][:] = x[idx[k]][:]
static __global__ void gather_rows_kernel(const double* __restrict__ x, int D, const long long* __restrict__ idx,
                                          int M, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)M * D) return;
  const int k = (int)(t / D), d = (int)(t % D);
  out[t] = x[idx[k] * D + d];
}

// mirror the lower triangle of a k x k matrix (ld k) into the upper one
static __global__ void symmetrize_lower_kernel(double* __restrict__ A, int k) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)k * k) return;
  const int i = (int)(idx / k), j = (int)(idx % k);
  if (j > i) A[idx] = A[(long long)j * k + i];
}

// out = a*X + b*Y (element-wise, n elements)
static __global__ void axpby_kernel(double* out, double a, const double* X, double b, const double* Y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * X[i] + b * Y[i];
}

// out = a * X o Y (element-wise product, n elements; out may alias X or Y)
static __global__ void hadamard_kernel(double* out, double a, const double* X, const double* Y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * X[i] * Y[i];
}

// *out = sum_ij X_ij Y_ij over the full symmetric k x k matrices given by their lower triangles
static __global__ void sym_frobenius_kernel(const double* __restrict__ X, const double* __restrict__ Y, int k,
                                            double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (long long idx = threadIdx.x; idx < (long long)k * k; idx += 1024) {
    const int i = (int)(idx / k), j = (int)(idx % k);
    if (j < i) s += 2.0 * X[idx] * Y[idx];
    else if (j == i) s += X[idx] * Y[idx];
  }
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = s;
}

// *out = sum_i x_i y_i (single CTA, fixed order)
static __global__ void dot_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n,
                                  double* __restrict__ out, int accumulate) {
  __shared__ double sh[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) s += x[i] * y[i];
  s = block_sum<1024>(s, sh);
  if (threadIdx.x == 0) *out = accumulate ? *out + s : s;
}

// A *= a
static __global__ void scale_kernel(double* __restrict__ A, long long n, double a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) A[idx] *= a;
}

}  // namespace sk

#define GPXB_LAUNCHED(ctx) \
  do {                     \
    GPXB_LAUNCH_CHECK();   \
    (ctx)->launches++;     \
  } while (0)

}  // namespace gpxb
