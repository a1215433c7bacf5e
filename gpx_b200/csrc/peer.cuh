// Device-side primitives of the peer-memory mailbox (comm.cuh): release / acquire flags at system scope and a
// bounded wait.  Used by the small all-reduce (comm.cu) and the fused RPCholesky pivot exchange (rpchol.cu).
#pragma once
#include <cstdint>

#include "comm.cuh"

namespace gpxb {

constexpr unsigned long long PEER_TIMEOUT_NS = 120ULL * 1000ULL * 1000ULL * 1000ULL;  // a lost peer traps, never hangs

__device__ __forceinline__ unsigned long long* peer_flags(double* base) {
  return reinterpret_cast<unsigned long long*>(base + peer_off_flags());
}
__device__ __forceinline__ void peer_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= seq (sequence numbers only grow)
__device__ __forceinline__ void peer_wait(const unsigned long long* p, unsigned long long seq) {
  if (peer_ld_acquire(p) >= seq) return;
  const unsigned long long t0 = peer_now_ns();
  while (peer_ld_acquire(p) < seq) {
    if (peer_now_ns() - t0 > PEER_TIMEOUT_NS) __trap();
  }
}

}  // namespace gpxb
