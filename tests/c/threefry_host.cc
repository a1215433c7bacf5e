// Host build of gpx_b200/csrc/threefry.cuh (the header is written for host and device): C wrappers so that
// tests/test_device_prng_cpu.py can run the very source the kernels use -- Threefry-2x32-20, jax.random.split and
// the float64 uniform draw in both jax_threefry_partitionable layouts -- on the CPU and compare it bit for bit with
// the pinned oracle.  Test infrastructure only.
#include "../../gpx_b200/csrc/threefry.cuh"

extern "C" {

void t_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out) {
  gpxb::threefry2x32(k0, k1, x0, x1, out[0], out[1]);
}

void t_split2(uint32_t k0, uint32_t k1, int which, int partitionable, uint32_t* out) {
  gpxb::jax_split2(k0, k1, which, partitionable, out[0], out[1]);
}

void t_uniform(uint32_t k0, uint32_t k1, unsigned long long total, int partitionable, double* out) {
  for (unsigned long long e = 0; e < total; ++e) out[e] = gpxb::jax_uniform_f64(k0, k1, e, total, partitionable);
}

}  // extern "C"
