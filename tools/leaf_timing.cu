// Developer tool (not part of the library): where the time of the panel factorisation's dependency chain goes.
//   1. the one-CTA Cholesky leaf timed to each phase boundary (the same kernel returning after phase k: no clock
//      reads inside the measured code), for the element-wise and the bulk-TMA instantiation;
//   2. the three chain kernels back to back on one stream;
//   3. the in-context timeline (%globaltimer stamps of block 0) of leaf / head solve / head update inside
//      potrf_lower at n = 1024 and 2000.
// chol.cu is compiled into this translation unit with -DGPXB_LEAF_STAMPS (which adds the phase exits and the stamps);
// everything else comes from the built library:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DGPXB_LEAF_STAMPS \
//        -I gpx_b200/csrc -I include tools/leaf_timing.cu -o tools/_leaf_timing \
//        -L gpx_b200 -l:libgpxb200.so -Xlinker -rpath='$ORIGIN/../gpx_b200' -ldl
// Output of the round-2 runs: profiles/r2_leaf_phase_timeline.txt.
#include "../gpx_b200/csrc/chol.cu"

#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <array>
#include "gpxb200.h"
#include <random>

int main() {
  using namespace gpxb;
  const int n = 128;
  std::vector<double> h(n * n), g(n * n);
  std::mt19937_64 rng(1);
  std::normal_distribution<double> nd;
  for (auto& v : g) v = nd(rng);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = (i == j) ? n : 0.0;
      for (int k = 0; k < n; ++k) s += g[i * n + k] * g[j * n + k];
      h[i * n + j] = s;
    }
  double *A, *inv;
  cudaMalloc(&A, n * n * 8);
  cudaMalloc(&inv, n * n * 8);
  cudaFuncSetAttribute((const void*)potrf_leaf_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF2_SMEM);
  cudaFuncSetAttribute((const void*)potrf_leaf_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF2_SMEM);
  for (int io : {0, 1}) {  // 1: the FAST instantiation (bulk-TMA I/O only)
  printf("bulk io = %d\n", io);
  const int stops[] = {0, 1, 4, 7, 10, 13, 16, 19, 22, 25, 27, 28, 29, 31};
  for (int stop : stops) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaMemcpy(A, h.data(), n * n * 8, cudaMemcpyHostToDevice);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (io) potrf_leaf_dmma_kernel<true><<<1, LEAF_NT, LEAF2_SMEM>>>(A, n, n, inv, 1 | ((stop + 1) << 8));
      else potrf_leaf_dmma_kernel<false><<<1, LEAF_NT, LEAF2_SMEM>>>(A, n, n, inv, 1 | ((stop + 1) << 8));
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("stop %2d: %.1f us  err=%s\n", stop, best * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  std::vector<double> L(n * n), X(n * n);
  cudaMemcpy(L.data(), A, n * n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(X.data(), inv, n * n * 8, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, up = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (j > i) { up = std::max(up, std::abs(X[i * n + j])); continue; }
      double s1 = 0, s2 = 0;
      for (int k = 0; k <= j; ++k) s1 += L[i * n + k] * L[j * n + k];
      for (int k = j; k <= i; ++k) s2 += X[i * n + k] * L[k * n + j];
      e1 = std::max(e1, std::abs(s1 - h[i * n + j]));
      e2 = std::max(e2, std::abs(s2 - (i == j ? 1.0 : 0.0)));
    }
  printf("max|LL^T-A| = %.3e  max|XL-I| = %.3e  max|upper X| = %.3e\n", e1, e2, up);
  }
  // the panel's critical chain in isolation: back-to-back launches on one stream (no bulk stream beside it)
  {
    const int N = 1024, lda2 = N;
    double *M, *invs;
    cudaMalloc(&M, (size_t)N * N * 8);
    cudaMalloc(&invs, 8 * 128 * 128 * 8);
    std::vector<double> hm((size_t)N * N, 0.0);
    for (int i = 0; i < N; ++i)
      for (int j = 0; j <= i; ++j) hm[(size_t)i * N + j] = (i == j) ? 4.0 : 0.001;
    cudaMemcpy(M, hm.data(), (size_t)N * N * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute((const void*)panel_trsm32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM32_SMEM);
    cudaFuncSetAttribute((const void*)panel_syrk32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK32_SMEM);
    for (int io = 0; io < 2; ++io)
      for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        const int reps = 50;
        for (int t = 0; t < 4; ++t) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          cudaEventRecord(e0);
          for (int r = 0; r < reps; ++r) {
            double* Ajj = M;                       // values drift; timing only
            double* B = M + (size_t)128 * lda2;
            if (mode == 0 || mode == 3) { if (io) potrf_leaf_dmma_kernel<true><<<1, LEAF_NT, LEAF2_SMEM>>>(Ajj, lda2, 128, invs, 1); else potrf_leaf_dmma_kernel<false><<<1, LEAF_NT, LEAF2_SMEM>>>(Ajj, lda2, 128, invs, 1); }
            if (mode == 1 || mode == 3) panel_trsm32_kernel<<<4, 256, TRSM32_SMEM>>>(B, lda2, 128, 128, invs, io);
            if (mode == 2 || mode == 3) panel_syrk32_kernel<<<10, 128, SYRK32_SMEM>>>(B, lda2, B + 128, lda2, 128, 128, io);
          }
          cudaEventRecord(e1);
          cudaDeviceSynchronize();
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (t >= 1) best = std::min(best, ms);
          cudaMemcpy(M, hm.data(), (size_t)N * N * 8, cudaMemcpyHostToDevice);
        }
        const char* names[] = {"leaf", "trsm32 (4 CTAs)", "syrk32 (10 CTAs)", "leaf + trsm32 + syrk32"};
        printf("bulk io = %d  %-24s %.2f us per launch group  err=%s\n", io, names[mode], best * 1e3 / reps,
               cudaGetErrorString(cudaGetLastError()));
      }
  }
  // timeline of the chain kernels inside a whole factorisation (the GEMMs of the bulk stream come from the library)
  {
    gpxb_ctx* c;
    if (gpxb_ctx_create(0, &c) != 0) { printf("ctx failed\n"); return 1; }
    Ctx* ctx = reinterpret_cast<Ctx*>(c);
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int N : {1024, 2000}) {
      double *M, *invs;
      cudaMalloc(&M, (size_t)N * N * 8);
      cudaMalloc(&invs, 16 * 128 * 128 * 8);
      std::vector<double> hm((size_t)N * N, 0.0);
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) hm[(size_t)i * N + j] = (i == j) ? 4.0 : 1.0 / (1.0 + std::abs(i - j));
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(M, hm.data(), (size_t)N * N * 8, cudaMemcpyHostToDevice);
        unsigned int zero = 0;
        cudaMemcpyToSymbol(g_chain_n, &zero, sizeof(zero));
        cudaDeviceSynchronize();
        // merge = 1: no phase stop
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        int rc = potrf_lower(ctx, M, N, N, invs, st);
        cudaEventRecord(e1, st);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("potrf_lower N=%d rc=%d: %.1f us  err=%s\n", N, rc, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
      }
      unsigned int cnt;
      cudaMemcpyFromSymbol(&cnt, g_chain_n, sizeof(cnt));
      std::vector<unsigned long long> lg(3 * 4096);
      cudaMemcpyFromSymbol(lg.data(), g_chain_log, lg.size() * 8);
      std::vector<std::array<unsigned long long, 3>> ev(cnt);
      for (unsigned i = 0; i < cnt; ++i) ev[i] = {lg[3 * i], lg[3 * i + 1], lg[3 * i + 2]};
      std::sort(ev.begin(), ev.end(), [](auto& a, auto& b) { return a[1] < b[1]; });
      const char* nm[] = {"leaf", "trsm32", "syrk32"};
      for (unsigned i = 0; i < cnt; ++i)
        printf("  %-7s begin %8.2f us  dur %6.2f us  gap-from-prev-end %6.2f us\n", nm[ev[i][0]],
               (ev[i][1] - ev[0][1]) * 1e-3, (ev[i][2] - ev[i][1]) * 1e-3,
               i ? ((double)ev[i][1] - (double)ev[i - 1][2]) * 1e-3 : 0.0);
    }
  }
  return 0;
}
