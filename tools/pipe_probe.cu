// Micro-benchmark: do DMMA.8x8x4 and the FP64 FMA pipe overlap on sm_100a, or share a datapath?
// Runs (a) DMMA only, (b) DFMA only, (c) both in every warp, (d) half the warps each, with
// identical per-kind instruction counts, and prints the time of each.  If (c) ~= (a)+(b) the two
// share the FP64 datapath; if (c) ~= max((a),(b)) they overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe.bin tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode bit0: DMMA stream, bit1: DFMA stream; split: 1 -> even warps DMMA, odd warps DFMA
template <int NM, int NF>
__global__ void probe(double* out, int iters, int mode, int split) {
  const int warp = threadIdx.x >> 5;
  bool do_m = mode & 1, do_f = mode & 2;
  if (split) { do_m = do_m && !(warp & 1); do_f = do_f && (warp & 1); }
  double c[NM][2], f[NF];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = i;
  for (int it = 0; it < iters; ++it) {
    if (do_m) {
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    if (do_f) {
#pragma unroll
      for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF>
float run(double* d, int iters, int mode, int split, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<NM, NF><<<148, threads>>>(d, 10, mode, split);
  cudaEventRecord(e0);
  probe<NM, NF><<<148, threads>>>(d, iters, mode, split);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
  const int iters = 20000;
  // per iteration per warp: 16 DMMA (16 x 512 flop) and 32 DFMA warp-instructions (32 x 64 flop)
  for (int threads : {128, 256, 512}) {
    const float m = run<16, 32>(d, iters, 1, 0, threads);
    const float f = run<16, 32>(d, iters, 2, 0, threads);
    const float both = run<16, 32>(d, iters, 3, 0, threads);
    const float split = run<16, 32>(d, iters, 3, 1, threads);
    const double warps = 148.0 * threads / 32;
    const double tf_m = warps * iters * 16 * 512.0 / (m * 1e-3) / 1e12;
    const double tf_f = warps * iters * 32 * 64.0 / (f * 1e-3) / 1e12;
    printf("{\"threads_per_sm\": %d, \"dmma_only_ms\": %.3f, \"dfma_only_ms\": %.3f, \"both_same_warp_ms\": %.3f, "
           "\"both_split_warps_ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}\n",
           threads, m, f, both, split, tf_m, tf_f);
  }
  return 0;
}
