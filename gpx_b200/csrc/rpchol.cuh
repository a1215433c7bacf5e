// Internal interface of the randomly pivoted Cholesky loop (Family 3c).
#pragma once
#include "common.cuh"
#include "gram.cuh"

namespace gpxb {
int rpcholesky_z(Ctx* ctx, int kind, const double* Z, int n, ZLayout zl, const uint32_t key[2], int M,
                 int partitionable, double* FT, long long ldf, long long* pivots, cudaStream_t s);
int transpose_f(Ctx* ctx, const double* FT, long long ldf, int n, int M, double* F, cudaStream_t s);
}  // namespace gpxb
