// NCCL plumbing (see comm.cuh).  Only allreduce(sum) on small buffers and the all-gather of
// the N x P Lanczos / CG block cross NVLink; the dense Cholesky never does.
#include "comm.cuh"

#include <dlfcn.h>

namespace gpxb {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g;

int load_nccl() {
  if (g.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g.h) break;
  }
  if (!g.h) {
    set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return 900;
  }
#define SYM(field, name)                                        \
  *(void**)(&g.field) = dlsym(g.h, name);                       \
  if (!g.field) {                                               \
    set_error("libnccl lacks %s", name);                        \
    return 901;                                                 \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}

#define NCCL_CHECK(expr)                                                         \
  do {                                                                           \
    ncclResult_t _r = (expr);                                                    \
    if (_r != 0) {                                                               \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g.GetErrorString(_r)); \
      return 1000 + (int)_r;                                                     \
    }                                                                            \
  } while (0)
}  // namespace

int comm_unique_id(void* out) {
  GPXB_TRY(load_nccl());
  ncclUniqueId id;
  NCCL_CHECK(g.GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return 0;
}

int comm_init(Ctx* ctx, const void* uid, int rank, int world) {
  GPXB_ARG_CHECK(ctx && world >= 1 && rank >= 0 && rank < world, -1);
  ctx->rank = rank;
  ctx->world = world;
  if (world == 1) return 0;
  GPXB_TRY(load_nccl());
  ncclUniqueId id;
  memcpy(&id, uid, sizeof(id));
  ncclComm_t c;
  GPXB_CUDA_CHECK(cudaSetDevice(ctx->device));
  NCCL_CHECK(g.CommInitRank(&c, world, id, rank));
  ctx->nccl_comm = c;
  return 0;
}

int comm_destroy(Ctx* ctx) {
  if (ctx && ctx->nccl_comm) {
    g.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return 0;
}

int comm_allreduce_sum(Ctx* ctx, double* buf, size_t count, cudaStream_t s) {
  if (ctx->world <= 1) return 0;
  PhaseScope ph(ctx, PH_ALLREDUCE, s);
  NCCL_CHECK(g.AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, s));
  return 0;
}

int comm_allreduce_sum2(Ctx* ctx, const double* send, double* recv, size_t count, cudaStream_t s) {
  if (ctx->world <= 1) {
    if (send != recv)
      GPXB_CUDA_CHECK(cudaMemcpyAsync(recv, send, sizeof(double) * count, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  PhaseScope ph(ctx, PH_ALLREDUCE, s);
  NCCL_CHECK(g.AllReduce(send, recv, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, s));
  return 0;
}

// The row partition is 1024-row aligned, so shards may differ by one 1024-row chunk.  Equal shards: one in-place
// ncclAllGather.  Unequal shards: the "all-gather-v" pattern, every rank sends its rows to every peer and receives
// theirs, all transfers of one call fused into a single NCCL group (one kernel; over NVSwitch every pair has its own
// full-bandwidth path, so the transfers run concurrently).
int comm_allgather_rows(Ctx* ctx, const double* local, double* all, int n_loc, int PS, cudaStream_t s) {
  if (ctx->world <= 1) return 0;
  PhaseScope ph(ctx, PH_ALLGATHER, s);
  const long long N = ctx->shard_N;  // every rank derives the same partition from the global row count
  const int W = ctx->world, me = ctx->rank;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  const long long b_me = shard_begin(N, W, me), e_me = shard_begin(N, W, me + 1);
  GPXB_ARG_CHECK(e_me - b_me == (long long)n_loc, -50);
  bool equal = true;
  for (int r = 0; r < W; ++r) equal = equal && (shard_begin(N, W, r + 1) - shard_begin(N, W, r) == e_me - b_me);
  double* mine = all + b_me * PS;
  if (local != mine)
    GPXB_CUDA_CHECK(cudaMemcpyAsync(mine, local, sizeof(double) * (size_t)n_loc * PS, cudaMemcpyDeviceToDevice, s));
  if (equal) {
    NCCL_CHECK(g.AllGather(mine, all, (size_t)n_loc * PS, ncclFloat64, comm, s));
    return 0;
  }
  ncclResult_t rc = g.GroupStart();
  for (int k = 1; k < W && rc == 0; ++k) {
    const int to = (me + k) % W, from = (me - k + W) % W;
    const long long bf = shard_begin(N, W, from), ef = shard_begin(N, W, from + 1);
    if (n_loc > 0) rc = g.Send(mine, (size_t)n_loc * PS, ncclFloat64, to, comm, s);
    if (rc == 0 && ef > bf) rc = g.Recv(all + bf * PS, (size_t)(ef - bf) * PS, ncclFloat64, from, comm, s);
  }
  const ncclResult_t rc_end = g.GroupEnd();  // always closed, also on the error path
  NCCL_CHECK(rc);
  NCCL_CHECK(rc_end);
  return 0;
}

}  // namespace gpxb
