// NCCL plumbing (see comm.cuh).  Only allreduce(sum) on small buffers and the all-gather of
// the N x P Lanczos / CG block cross NVLink; the dense Cholesky never does.
#include "comm.cuh"

#include <dlfcn.h>

#include <cstdlib>

#include "peer.cuh"

namespace gpxb {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclFloat64 = 8, ncclSum = 0 };
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

struct PeerState {
  PeerView view{};
  double* local = nullptr;
  unsigned long long seq = 0;
};

namespace {
// all-reduce (sum, fixed rank order) of count <= PEER_SLOT doubles through the mailboxes; one CTA
__global__ void __launch_bounds__(1024, 1) peer_allreduce_kernel(PeerView pv, double* __restrict__ buf, int count,
                                                                 unsigned long long seq) {
  const int par = (int)(seq & 1ULL), tid = threadIdx.x, W = pv.world, me = pv.me;
  // publish my contribution in slot `me` of every rank's mailbox (my own included)
  for (int r = 0; r < W; ++r) {
    double* dst = pv.base[r] + peer_off_ar(par, me);
    for (int i = tid; i < count; i += 1024) dst[i] = buf[i];
  }
  __syncthreads();
  if (tid < W) {
    __threadfence_system();
    peer_st_release(peer_flags(pv.base[tid]) + peer_flag_ar(par, me), seq);
  }
  // wait for every rank's contribution in MY mailbox
  if (tid < W) peer_wait(peer_flags(pv.base[me]) + peer_flag_ar(par, tid), seq);
  __syncthreads();
  const double* mine = pv.base[me] + peer_off_ar(par, 0);
  for (int i = tid; i < count; i += 1024) {
    double s = 0.0;
    for (int r = 0; r < W; ++r) s += __ldcg(mine + (size_t)r * PEER_SLOT + i);
    buf[i] = s;
  }
}

// Maps every rank's mailbox.  Any failure leaves ctx->peer null (the callers then use NCCL).
int peer_setup(Ctx* ctx) {
  const char* e = getenv("GPXB_PEER");
  if (e && e[0] == '0') return 0;
  if (ctx->world > PEER_MAX_WORLD) return 0;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  const int W = ctx->world, me = ctx->rank;
  PeerState* ps = new PeerState();
  unsigned char* stage = nullptr;
  cudaStream_t st = nullptr;
  bool ok = cudaMalloc(&ps->local, peer_mailbox_bytes()) == cudaSuccess &&
            cudaMemset(ps->local, 0, peer_mailbox_bytes()) == cudaSuccess &&
            cudaMalloc(&stage, (size_t)W * 64) == cudaSuccess && cudaStreamCreate(&st) == cudaSuccess &&
            cudaDeviceSynchronize() == cudaSuccess;
  cudaIpcMemHandle_t mine, all[PEER_MAX_WORLD];
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  int have = ok && cudaIpcGetMemHandle(&mine, ps->local) == cudaSuccess ? 1 : 0;
  if (!have) memset(&mine, 0, sizeof(mine));
  // the exchange is collective: every rank takes part even when its own export failed (zero handle = "no")
  if (stage && st) {
    cudaMemcpyAsync(stage + (size_t)me * 64, &mine, 64, cudaMemcpyHostToDevice, st);
    if (g.AllGather(stage + (size_t)me * 64, stage, 64, ncclUint8, comm, st) != 0) have = 0;
    if (cudaMemcpyAsync(all, stage, (size_t)W * 64, cudaMemcpyDeviceToHost, st) != cudaSuccess) have = 0;
    if (cudaStreamSynchronize(st) != cudaSuccess) have = 0;
  } else {
    have = 0;
  }
  ps->view.world = W;
  ps->view.me = me;
  for (int r = 0; r < PEER_MAX_WORLD; ++r) ps->view.base[r] = nullptr;
  cudaIpcMemHandle_t zero;
  memset(&zero, 0, sizeof(zero));
  for (int r = 0; r < W && have; ++r) {
    if (r == me) { ps->view.base[r] = ps->local; continue; }
    if (memcmp(&all[r], &zero, 64) == 0) { have = 0; break; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { have = 0; break; }
    ps->view.base[r] = (double*)p;
  }
  // agree: the mailbox is used only if EVERY rank mapped every peer (sum of `have` == world)
  double agree = have ? 1.0 : 0.0, *dagree = nullptr;
  if (stage && st && cudaMalloc(&dagree, sizeof(double)) == cudaSuccess) {
    cudaMemcpyAsync(dagree, &agree, sizeof(double), cudaMemcpyHostToDevice, st);
    g.AllReduce(dagree, dagree, 1, ncclFloat64, ncclSum, comm, st);
    cudaMemcpyAsync(&agree, dagree, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(dagree);
  } else {
    agree = 0.0;
  }
  if (stage) cudaFree(stage);
  if (st) cudaStreamDestroy(st);
  (void)cudaGetLastError();
  if ((int)(agree + 0.5) == W) {
    ctx->peer = ps;
  } else {
    for (int r = 0; r < W; ++r)
      if (r != me && ps->view.base[r]) cudaIpcCloseMemHandle(ps->view.base[r]);
    if (ps->local) cudaFree(ps->local);
    delete ps;
    (void)cudaGetLastError();
  }
  return 0;
}
}  // namespace

bool comm_peer_ready(const Ctx* ctx) { return ctx && ctx->peer != nullptr; }
bool comm_peer_view(const Ctx* ctx, PeerView* v) {
  if (!comm_peer_ready(ctx)) return false;
  *v = reinterpret_cast<const PeerState*>(ctx->peer)->view;
  return true;
}
unsigned long long comm_peer_next_seq(Ctx* ctx) { return ++reinterpret_cast<PeerState*>(ctx->peer)->seq; }

namespace {
// NCCL sets up the channels of a communication pattern at its first use (hundreds of ms for the all-to-all send /
// receive of the unequal-shard all-gather).  Run every pattern the library uses once, on a few bytes, at communicator
// creation, so that the first sharded mat-vec does not pay for it.  Collective: every rank calls it from comm_init.
int comm_warmup(Ctx* ctx) {
  const int W = ctx->world, me = ctx->rank;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  double* buf = nullptr;
  GPXB_CUDA_CHECK(cudaMalloc(&buf, sizeof(double) * 8 * (size_t)(W + 1)));
  GPXB_CUDA_CHECK(cudaMemset(buf, 0, sizeof(double) * 8 * (size_t)(W + 1)));
  cudaStream_t st = nullptr;  // the legacy default stream: ordered with the memset, synchronised below
  ncclResult_t rc = g.AllReduce(buf, buf, 8, ncclFloat64, ncclSum, comm, st);
  if (rc == 0) rc = g.AllGather(buf + 8 * (size_t)me, buf, 8, ncclFloat64, comm, st);
  if (rc == 0) {
    rc = g.GroupStart();
    for (int k = 1; k < W && rc == 0; ++k) {
      const int to = (me + k) % W, from = (me - k + W) % W;
      rc = g.Send(buf + 8 * (size_t)W, 8, ncclFloat64, to, comm, st);
      if (rc == 0) rc = g.Recv(buf + 8 * (size_t)from, 8, ncclFloat64, from, comm, st);
    }
    const ncclResult_t rc_end = g.GroupEnd();
    if (rc == 0) rc = rc_end;
  }
  const cudaError_t ce = cudaStreamSynchronize(st);
  cudaFree(buf);
  NCCL_CHECK(rc);
  GPXB_CUDA_CHECK(ce);
  return 0;
}
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
  GPXB_TRY(peer_setup(ctx));
  return comm_warmup(ctx);
}

int comm_destroy(Ctx* ctx) {
  if (ctx && ctx->peer) {
    PeerState* ps = reinterpret_cast<PeerState*>(ctx->peer);
    for (int r = 0; r < ps->view.world; ++r)
      if (r != ps->view.me && ps->view.base[r]) cudaIpcCloseMemHandle(ps->view.base[r]);
    cudaFree(ps->local);
    delete ps;
    ctx->peer = nullptr;
  }
  if (ctx && ctx->nccl_comm) {
    g.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return 0;
}

int comm_allreduce_sum(Ctx* ctx, double* buf, size_t count, cudaStream_t s) {
  if (ctx->world <= 1) return 0;
  PhaseScope ph(ctx, PH_ALLREDUCE, s);
  if (ctx->peer && count <= (size_t)PEER_SLOT) {  // latency-bound: one kernel over the peer mailboxes
    PeerState* ps = reinterpret_cast<PeerState*>(ctx->peer);
    peer_allreduce_kernel<<<1, 1024, 0, s>>>(ps->view, buf, (int)count, ++ps->seq);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
    return 0;
  }
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
