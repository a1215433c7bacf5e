// Threefry-2x32-20 (Random123) and the jax.random bit layouts built on it, usable on host and
// device.  Restates the third-party generator the reference samples with
// (jax.random.split / uniform / choice / rademacher; call sites
// gpx/kernels/approximations.py:100-101, gpx/lanczos.py:168-169) for both values of
// jax_threefry_partitionable.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#define GPXB_HD __host__ __device__ __forceinline__
#else
#define GPXB_HD inline
#endif

namespace gpxb {

GPXB_HD uint32_t rotl32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }

GPXB_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& y0, uint32_t& y1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = rotl32(x1, R[i & 1][r]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  y0 = x0;
  y1 = x1;
}

// key_out = jax.random.split(key, 2)[which]
GPXB_HD void jax_split2(uint32_t k0, uint32_t k1, int which, int partitionable, uint32_t& o0, uint32_t& o1) {
  if (partitionable) {
    threefry2x32(k0, k1, 0u, (uint32_t)which, o0, o1);
  } else {
    uint32_t a0, a1, b0, b1;
    threefry2x32(k0, k1, 0u, 2u, a0, a1);  // (y0[0], y1[0])
    threefry2x32(k0, k1, 1u, 3u, b0, b1);  // (y0[1], y1[1])
    if (which == 0) { o0 = a0; o1 = b0; } else { o0 = a1; o1 = b1; }
  }
}

// element e of jax.random.uniform(key, shape, float64) with prod(shape) == total
GPXB_HD double jax_uniform_f64(uint32_t k0, uint32_t k1, unsigned long long e, unsigned long long total,
                               int partitionable) {
  uint32_t hi, lo;
  if (partitionable) threefry2x32(k0, k1, (uint32_t)(e >> 32), (uint32_t)e, hi, lo);
  else threefry2x32(k0, k1, (uint32_t)e, (uint32_t)(total + e), hi, lo);
  const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
  const unsigned long long fb = (bits >> 12) | 0x3FF0000000000000ULL;
  double f;
#ifdef __CUDA_ARCH__
  f = __longlong_as_double((long long)fb);
#else
  union { unsigned long long u; double d; } cv;
  cv.u = fb;
  f = cv.d;
#endif
  return f - 1.0;
}

}  // namespace gpxb
