// Common device/host helpers for libgpxb200 (sm_100a only).
#pragma once
#include <unordered_map>
#include <mutex>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace gpxb {

// ---------------------------------------------------------------------------
// Error convention: 0 = OK, <0 = argument error, >0 = CUDA/NCCL.
// The message is kept per thread and read back through gpxb_last_error().
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define GPXB_CUDA_CHECK(expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::gpxb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                        cudaGetErrorString(_e));                                      \
      return (int)_e > 0 ? (int)_e : 1;                                               \
    }                                                                                 \
  } while (0)

#define GPXB_ARG_CHECK(cond, code)                                                   \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::gpxb::set_error("%s:%d: argument check failed: %s", __FILE__, __LINE__, #cond); \
      return (code);                                                                  \
    }                                                                                 \
  } while (0)

#define GPXB_TRY(expr)                 \
  do {                                 \
    int _rc = (expr);                  \
    if (_rc != 0) return _rc;          \
  } while (0)

#define GPXB_LAUNCH_CHECK() GPXB_CUDA_CHECK(cudaGetLastError())

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------
// Device context: device properties, a side stream + events for fork/join of
// independent sub-problems inside one call, a launch counter, and (optionally)
// a NCCL communicator for the row-sharded matrix-free paths (comm.cu).
// All scratch memory is stream-ordered (cudaMallocAsync on the caller's stream),
// so every entry point is enqueue-only and re-entrant per context.
// ---------------------------------------------------------------------------
// Raises a kernel's dynamic shared-memory limit the first time (or when a larger size is asked for): the driver call
// costs microseconds and the launch paths run thousands of times per LML evaluation.
inline cudaError_t ensure_dyn_smem(const void* fn, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> seen;
  std::lock_guard<std::mutex> lk(mu);
  auto it = seen.find(fn);
  if (it != seen.end() && it->second >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) seen[fn] = bytes;
  return e;
}

// Launch `kernel` as a programmatic dependent of the previous kernel on the stream (its CTAs may become resident
// before that kernel has finished; it must execute griddepcontrol.wait before touching anything the predecessor
// writes).  Hides the dependent-launch latency of chains of short kernels.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// inside such a kernel: wait for the predecessor grid (no-op for a plain launch), then let the successor in
#define GPXB_PDL_WAIT_THEN_RELEASE()                              \
  do {                                                            \
    asm volatile("griddepcontrol.wait;" ::: "memory");            \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

struct Ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t side = nullptr;  // high-priority stream (panel factorisation look-ahead)
  cudaStream_t crit = nullptr;  // high-priority stream of the leaf -> head-solve -> head-update chain inside a panel
  // ticket counter of panel_head_kernel (chol.cu): CTAs of launch k have all read their inputs once the device
  // counter reaches the sum of the CTA counts of launches 0..k (launches are ordered on `crit`)
  unsigned long long* head_counter = nullptr;
  unsigned long long head_ticket = 0;
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};  // concurrent branches of the triangular inverse
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_pool;  // reusable timing-free events, grown on demand
  size_t ev_next = 0;
  // round-robin event from the pool (safe to reuse: an event is re-recorded only after
  // far more than pool-size later records on in-order streams)
  int next_event(cudaEvent_t* out) {
    if (ev_pool.size() < 256) {
      cudaEvent_t e;
      GPXB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev_pool.push_back(e);
      *out = e;
      return 0;
    }
    *out = ev_pool[ev_next++ % ev_pool.size()];
    return 0;
  }
  unsigned long long launches = 0;  // kernels launched by this library via this ctx
  // optional per-launch timing of the DMMA GEMM (bench.py roofline): event pairs + algorithmic flops
  bool prof = false;
  bool serialize = false;  // profiling aid: look-ahead work stays on the caller's stream (clean per-launch times)
  struct ProfRec { cudaEvent_t e0, e1; double flops; };
  std::vector<ProfRec> prof_recs;
  void* nccl_comm = nullptr;
  // Lanczos breakdown branch (gpx/lanczos.py:92-108): key carried into lanczos_tridiagonal; off: breakdown only flags
  bool restart_on = false;
  uint32_t restart_key[2] = {0u, 0u};
  int restart_partitionable = 1;
  void* peer = nullptr;  // PeerState* (comm.cu): IPC-mapped mailboxes of all ranks, or null
  int rank = 0, world = 1;
  long long shard_N = 0;  // global row count of the sharded problem in flight
  // optional per-phase timing of the row-sharded paths (bench.py "phases"): event pairs per phase id
  bool phase_prof = false;
  double phase_weight = 1.0;  // a loop that brackets only every k-th iteration records with weight k
  struct PhaseRec { cudaEvent_t e0, e1; int id; double weight; };
  std::vector<PhaseRec> phase_recs;
};

// Phase ids of gpxb_phase_collect (include/gpxb200.h lists the same names).
enum Phase {
  PH_MATVEC = 0,      // matrix-free block mat-vec (+ its column-split reduction)
  PH_ALLGATHER = 1,   // all-gather of the N x P vector block (NCCL; includes waiting for the slowest rank)
  PH_DOTS = 2,        // column dot products / norms (local partial sums)
  PH_ALLREDUCE = 3,   // every all-reduce (NCCL; includes waiting for the slowest rank)
  PH_VECOPS = 4,      // axpy / normalise / copy passes over the local vector blocks
  PH_RP_PIVOT = 5,    // RPCholesky: draw + inverse-CDF search + pivot row gather
  PH_RP_COLUMN = 6,   // RPCholesky: fused kernel column / GEMV / diagonal update
  PH_SGPR_CHUNK = 7,  // SGPR: per-chunk K_nm block, TRSM against L_m, SYRK into B (streams overlap)
  PH_POTRF_REPL = 8,  // SGPR: replicated M x M Cholesky factorisations (K_mm and B) and the final solves
  PH_PRECOND = 9,     // PCG: preconditioner application
  PH_COUNT = 16
};

// Brackets the work enqueued on stream s during its lifetime with a CUDA event pair when phase profiling is on.
struct PhaseScope {
  Ctx* ctx;
  cudaStream_t s;
  int idx = -1;
  PhaseScope(Ctx* c, int id, cudaStream_t stream) : ctx(c), s(stream) {
    if (!c || !c->phase_prof) return;
    Ctx::PhaseRec r{nullptr, nullptr, id, c->phase_weight};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, s);
    idx = (int)c->phase_recs.size();
    c->phase_recs.push_back(r);
  }
  ~PhaseScope() {
    if (idx >= 0) cudaEventRecord(ctx->phase_recs[idx].e1, s);
  }
  PhaseScope(const PhaseScope&) = delete;
  PhaseScope& operator=(const PhaseScope&) = delete;
};

// Stream-ordered scratch buffer (RAII). alloc() returns 0 or a CUDA error code.
struct DevBuf {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  int alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    if (bytes == 0) bytes = 8;
    GPXB_CUDA_CHECK(cudaMallocAsync(&p, bytes, stream));
    return 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
  }
  ~DevBuf() { release(); }
};

}  // namespace gpxb
