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
// ---- peer-memory mailbox (NVLink / NVSwitch, CUDA IPC) ---------------------------------------------------------
// Latency-bound exchanges (a few KB, thousands per call) do not go through NCCL: every rank owns a small mailbox in
// device memory that all peers map (cudaIpc) and write with plain stores over NVLink; a sequence-numbered flag per
// (parity, rank) publishes the data (st.release.sys / ld.acquire.sys).  Two parities alternate, so a fast rank can
// post exchange k + 1 while a slow one still reads exchange k.  One process per GPU, so a rank that waits never
// shares a GPU with the rank it waits for; waits are bounded (trap after PEER_TIMEOUT_NS).
constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_SLOT = 8192;  // doubles per (parity, rank) slot of the all-reduce area, and capacity of the RPCholesky areas
struct PeerView {
  double* base[PEER_MAX_WORLD];  // mailbox of every rank, as mapped into this process (base[me] is local)
  int world, me;
};
// mailbox layout (doubles): [2][world][PEER_SLOT] all-reduce slots | [2][PEER_SLOT] RPCholesky block sums |
// [2][PEER_SLOT] RPCholesky pivot buffer | flags (uint64): [2][world] all-reduce, [2][world] block sums, [2] pivot
__host__ __device__ inline size_t peer_off_ar(int par, int r) { return ((size_t)par * PEER_MAX_WORLD + r) * PEER_SLOT; }
__host__ __device__ inline size_t peer_off_bs(int par) { return (size_t)2 * PEER_MAX_WORLD * PEER_SLOT + (size_t)par * PEER_SLOT; }
__host__ __device__ inline size_t peer_off_pb(int par) { return (size_t)(2 * PEER_MAX_WORLD + 2) * PEER_SLOT + (size_t)par * PEER_SLOT; }
__host__ __device__ inline size_t peer_off_flags() { return (size_t)(2 * PEER_MAX_WORLD + 4) * PEER_SLOT; }
constexpr size_t PEER_FLAG_WORDS = 2 * PEER_MAX_WORLD + 2 * PEER_MAX_WORLD + 2;
inline size_t peer_mailbox_bytes() { return (peer_off_flags() + PEER_FLAG_WORDS) * sizeof(double); }
// flag word indices
__host__ __device__ inline int peer_flag_ar(int par, int r) { return par * PEER_MAX_WORLD + r; }
__host__ __device__ inline int peer_flag_bs(int par, int r) { return 2 * PEER_MAX_WORLD + par * PEER_MAX_WORLD + r; }
__host__ __device__ inline int peer_flag_pb(int par) { return 4 * PEER_MAX_WORLD + par; }

bool comm_peer_ready(const Ctx* ctx);
// fills v; returns false when the mailbox is not available (single GPU, IPC refused, GPXB_PEER=0)
bool comm_peer_view(const Ctx* ctx, PeerView* v);
// next exchange sequence number (identical on every rank: all ranks make the same calls in the same order)
unsigned long long comm_peer_next_seq(Ctx* ctx);

// row partition used everywhere: rank r owns rows [row_begin(r), row_begin(r+1))
inline long long shard_begin(long long N, int world, int r, int align = 1024) {
  const long long chunks = (N + align - 1) / align;
  const long long per = chunks / world, rem = chunks % world;
  const long long c = (long long)r * per + (r < rem ? r : rem);
  const long long b = c * align;
  return b < N ? b : N;
}
}  // namespace gpxb
