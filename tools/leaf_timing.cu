// Time to each phase boundary of the one-CTA Cholesky leaf (developer tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DGPXB_LEAF_STAMPS \
//        -I gpx_b200/csrc -I include tools/leaf_timing.cu $(ls gpx_b200/build/*.o | grep -v chol.o) -o tools/_leaf_timing -ldl
#include "../gpx_b200/csrc/chol.cu"

#include <cstdio>
#include <cstring>
#include <cmath>
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
  cudaFuncSetAttribute((const void*)potrf_leaf_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF2_SMEM);
  for (int io = 0; io < 2; ++io) {
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
      potrf_leaf_dmma_kernel<<<1, LEAF_NT, LEAF2_SMEM>>>(A, n, n, inv, 1 | (stop << 8), io);
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
  return 0;
}
