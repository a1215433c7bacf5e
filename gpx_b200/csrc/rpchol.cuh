// Internal interface of the randomly pivoted Cholesky loop (Family 3c).
#pragma once
#include "common.cuh"
#include "gram.cuh"

namespace gpxb {
// number of doubles of the row-blocked factor buffer for n rows and M pivots
size_t rp_factor_doubles(int n, int M);
int rpcholesky_z(Ctx* ctx, int kind, const double* Z, int n, long long row0, long long Ntot, ZLayout zl,
                 const uint32_t key[2], int M, int partitionable, double* FB, long long* pivots, cudaStream_t s);
// F (N x M row-major) <- blocked factor
int rp_unblock_rowmajor(Ctx* ctx, const double* FB, int n, int M, double* F, cudaStream_t s);
// FT (M x ldf, one contiguous row per pivot) <- blocked factor
int rp_unblock_transposed(Ctx* ctx, const double* FB, int n, int M, double* FT, long long ldf, cudaStream_t s);
}  // namespace gpxb
