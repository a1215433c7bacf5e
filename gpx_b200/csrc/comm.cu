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
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
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
  SYM(Broadcast, "ncclBroadcast")
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
  NCCL_CHECK(g.AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, s));
  return 0;
}

int comm_allreduce_sum2(Ctx* ctx, const double* send, double* recv, size_t count, cudaStream_t s) {
  if (ctx->world <= 1) {
    if (send != recv)
      GPXB_CUDA_CHECK(cudaMemcpyAsync(recv, send, sizeof(double) * count, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  NCCL_CHECK(g.AllReduce(send, recv, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, s));
  return 0;
}

int comm_allgather_rows(Ctx* ctx, const double* local, double* all, int n_loc, int PS, cudaStream_t s) {
  if (ctx->world <= 1) return 0;
  // unequal row counts: one broadcast per owner, grouped into a single launch
  long long N = 0;
  // every rank derives the same partition from the global row count, which is the sum of
  // n_loc over ranks; callers pass consistent shards, so recompute from shard_begin
  // (N is recovered by the caller-side convention: ctx->shard_N)
  N = ctx->shard_N;
  NCCL_CHECK(g.GroupStart());
  for (int r = 0; r < ctx->world; ++r) {
    const long long b = shard_begin(N, ctx->world, r), e = shard_begin(N, ctx->world, r + 1);
    const size_t cnt = (size_t)(e - b) * PS;
    if (cnt == 0) continue;
    const void* src = (r == ctx->rank) ? (const void*)local : (const void*)(all + b * PS);
    NCCL_CHECK(g.Broadcast(src, all + b * PS, cnt, ncclFloat64, r, (ncclComm_t)ctx->nccl_comm, s));
  }
  NCCL_CHECK(g.GroupEnd());
  (void)n_loc;
  return 0;
}

}  // namespace gpxb
