// Internal dense-factorisation interface (row-major, lower triangle).
#pragma once
#include "common.cuh"

namespace gpxb {

constexpr int CHOL_LEAF = 128;

inline size_t leaf_inv_doubles(int n) {
  return (size_t)((n + CHOL_LEAF - 1) / CHOL_LEAF) * CHOL_LEAF * CHOL_LEAF;
}

int potrf_lower(Ctx* ctx, double* A, long long lda, int n, double* inv, cudaStream_t s);
int trsm_right_lt(Ctx* ctx, const double* L, long long ldl, const double* inv, double* B,
                  long long ldb, int m, int n, cudaStream_t s);
int trsm_right_l(Ctx* ctx, const double* L, long long ldl, const double* inv, double* B,
                 long long ldb, int m, int n, cudaStream_t s);
int potri_lower(Ctx* ctx, const double* L, long long ldl, const double* inv, int n, double* W,
                double* T, double* out, long long ldo, cudaStream_t s);
int logdet_lower(Ctx* ctx, const double* L, long long ld, int n, double* out, double scale,
                 cudaStream_t s);
size_t potri_scratch_doubles(int n);

// Block-column packed lower storage ("BCP") for systems whose square matrix does not fit in HBM
// (the 180 000-row Hessian system is 259 GB square, 131 GB packed): block column p = columns
// [p*1024, (p+1)*1024) of rows p*1024..n is a dense row-major (n - p*1024) x 1024 panel and the
// panels are stored back to back.  Every GEMM of the right-looking factorisation touches one
// panel per operand, so the same kernels run on it with ld = 1024.
struct BcpLayout {
  static constexpr int NB = 1024;
  long long n;
  __host__ __device__ long long off(long long p) const { return p * n * NB - (long long)NB * NB * (p * (p - 1) / 2); }
  __host__ __device__ int panels() const { return (int)((n + NB - 1) / NB); }
  __host__ __device__ size_t doubles() const { return (size_t)off(panels()); }
  // v with v[r * NB + c] == element (r, c) for c in block column p and r >= p * NB
  template <class T>
  __host__ __device__ T* vbase(T* base, long long p) const { return base + off(p) - p * NB * NB - p * NB; }
  __host__ __device__ long long index(long long r, long long c) const {
    const long long p = c / NB;
    return off(p) + (r - p * NB) * NB + (c - p * NB);
  }
};
int potrf_lower_bcp(Ctx* ctx, double* Ap, int n, double* inv, cudaStream_t s);
// m <= 4 right-hand sides stored as rows of B: forward solves L x = b, otherwise L^T x = b
int trsv_bcp(Ctx* ctx, const double* Lp, int n, const double* inv, double* B, long long ldb, int m, bool forward,
             cudaStream_t s);
int logdet_bcp(Ctx* ctx, const double* Lp, int n, double* out, double scale, cudaStream_t s);

}  // namespace gpxb
