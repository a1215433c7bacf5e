// Warp-level FP64 tensor-core (DMMA) and async-copy primitives for sm_100a.
//
// sm_100a has no tcgen05.mma kind for FP64; FP64 tensor math is the warp-level
// mma.sync m8n8k4 (SASS: DMMA.8x8x4) with register operands fed from shared
// memory.  Fragment layout (PTX ISA, mma.m8n8k4 .f64):
//   A (8x4, row): lane holds A[lane>>2][lane&3]
//   B (4x8, col): lane holds B[lane&3][lane>>2]
//   C (8x8)     : lane holds C[lane>>2][2*(lane&3) + {0,1}]
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace gpxb {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte LDGSTS with zero fill: copies src_bytes (0, 8 or 16) and zero-fills the rest.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
// 8-byte LDGSTS with zero fill (src_bytes 0 or 8); used when 16-byte alignment does not hold.
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) for contiguous tiles.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity));
  return ok != 0;
}
// Bounded wait: a lost transaction traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > (1u << 24)) __trap();
  }
}
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Bulk store shared -> global (SASS: UBLKCP.S.G); dst/src 16-byte aligned, bytes a multiple of 16.  The issuing
// thread commits and waits for the READ of shared memory before the CTA may reuse or release it.
__device__ __forceinline__ void tma_bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

// The warp index through a lane-0 shuffle: a value the compiler can prove warp-uniform.  Branching on it keeps the
// warp converged in the compiler's eyes, so __shfl_sync / warp reductions below such a branch compile to bare SHFL;
// with `threadIdx.x >> 5` each one is wrapped in WARPSYNC.COLLECTIVE ... ENDCOLLECTIVE and register moves (the
// Cholesky leaf shrank from 14 440 to 6 448 instructions).  Must be called by all lanes of the warp.
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum (fixed tree): result valid in thread 0.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh /* >= NT/32 doubles */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < NT / 32) ? sh[threadIdx.x] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace gpxb
