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

}  // namespace gpxb
