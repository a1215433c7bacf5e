// Internal interface of the randomly pivoted Cholesky loop (Family 3c).
#pragma once
#include <functional>

#include "common.cuh"
#include "gram.cuh"

namespace gpxb {
// Generator of the pivot's matrix column for the external-column variant: pbuf (device) holds
// [0] the pivot index, [2, 2 + S) the pivot's row of `rows`; it must launch work on the stream that fills the
// column buffer handed to rpcholesky_ext.
using RpColumnGen = std::function<int(const double* pbuf, cudaStream_t s)>;
// number of doubles of the row-blocked factor buffer for n rows and M pivots
size_t rp_factor_doubles(int n, int M);
int rpcholesky_z(Ctx* ctx, int kind, const double* Z, int n, long long row0, long long Ntot, ZLayout zl,
                 const uint32_t key[2], int M, int partitionable, double* FB, long long* pivots, cudaStream_t s);
int rpcholesky_ext(Ctx* ctx, const double* rows, int n, ZLayout zl, const double* diag0, double* ext_col,
                   const RpColumnGen& gen, const uint32_t key[2], int M, int partitionable, double* FB,
                   long long* pivots, cudaStream_t s);
// F (N x M row-major) <- blocked factor
int rp_unblock_rowmajor(Ctx* ctx, const double* FB, int n, int M, double* F, cudaStream_t s);
// FT (M x ldf, one contiguous row per pivot) <- blocked factor
int rp_unblock_transposed(Ctx* ctx, const double* FB, int n, int M, double* FT, long long ldf, cudaStream_t s);
}  // namespace gpxb
