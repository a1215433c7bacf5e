// Collectives for the row-sharded matrix-free paths (NCCL over NVLink 5 / NVSwitch).  One
// process per GPU; with world == 1 every call is a no-op.  NCCL is bound at run time from the
// libnccl.so.2 the process already holds (torch's bundled copy), so the library itself links
// against nothing but cudart.
#pragma once
#include "common.cuh"

namespace gpxb {
int comm_init(Ctx* ctx, const void* unique_id_128bytes, int rank, int world);
int comm_unique_id(void* out_128bytes);
int comm_destroy(Ctx* ctx);
int comm_allreduce_sum(Ctx* ctx, double* buf, size_t count, cudaStream_t s);
int comm_allreduce_sum2(Ctx* ctx, const double* send, double* recv, size_t count, cudaStream_t s);
// all[N][PS] <- concatenation over ranks of local[n_loc][PS] (row counts may differ by rank)
int comm_allgather_rows(Ctx* ctx, const double* local, double* all, int n_loc, int PS, cudaStream_t s);
// row partition used everywhere: rank r owns rows [row_begin(r), row_begin(r+1))
inline long long shard_begin(long long N, int world, int r, int align = 1024) {
  const long long chunks = (N + align - 1) / align;
  const long long per = chunks / world, rem = chunks % world;
  const long long c = (long long)r * per + (r < rem ? r : rem);
  const long long b = c * align;
  return b < N ? b : N;
}
}  // namespace gpxb
