// Family 3c: randomly pivoted Cholesky landmark loop on device (no host round trip per pivot).
//
// Replaces gpx/kernels/approximations.py:30-129 (rpcholesky).  Per pivot i:
//   prob = diag / sum(diag);  key, _ = split(key);  s = choice(key, N, p = prob)
//   g = k(x_s, X) - F[:, :i] F[s, :i]^T;  F[:, i] = g / (sqrt(g_s) + 1e-6);  diag -= F[:, i]^2
// The factor is kept in a row-blocked transposed layout: for every block of 1024 data rows the
// pivot columns are consecutive 8 KB chunks (element (c, r) at ((r/1024)*M + c)*1024 + r%1024),
// so the per-pivot GEMV -- the HBM-bound part, 8*N*i bytes at step i -- is, for each CTA, one
// sequential stream (no page hopping), fully coalesced with one thread per data row; the
// kernel column, the scaling, the diagonal update and the fixed-order block sums for the next
// draw are fused into the same pass.
// Sampling reproduces jax.random.choice (inverse CDF: cumsum, r = cum[-1]*(1-u),
// searchsorted-left) with a two-level scan over fixed 1024-row blocks, so the result does not
// depend on the launch geometry.
#include "../../include/gpxb200.h"
#include "gram.cuh"
#include "mma.cuh"
#include "rpchol.cuh"
#include "comm.cuh"
#include "peer.cuh"

#include <algorithm>
#include "threefry.cuh"

namespace gpxb {

namespace {

constexpr int BLK = 1024;

__device__ __forceinline__ double kfun(int kind, double d2) {
  if (kind == KIND_SE) return exp(-d2);
  const double r = sqrt(fmax(d2, 1e-36));
  if (kind == KIND_M12) return exp(-r);
  if (kind == KIND_M32) {
    const double d = 1.7320508075688772 * r;
    return (1.0 + d) * exp(-d);
  }
  const double d = 2.23606797749979 * r;
  return (1.0 + d + d * d / 3.0) * exp(-d);
}

// inclusive scan of one value per thread across a 1024-thread CTA (fixed tree order)
__device__ __forceinline__ double block_scan_1024(double v, double* sh /*32*/) {
  const int lane = threadIdx.x & 31, w = uniform_warp_id();  // warp-uniform: the `w == 0` branch keeps bare shuffles
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (lane == 31) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double s = sh[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    sh[lane] = s;
  }
  __syncthreads();
  if (w > 0) v += sh[w - 1];
  return v;
}

struct RpState {
  double* diag;       // [n]
  double* blocksum;   // [nblk] block sums of diag over ALL rows (after the all-reduce)
  double* FB;         // row-blocked factor, see fb_index
  int M;
  const double* Z;    // [n][S]
  int n, S, Dp, nblk, kind;
  long long* pivots;  // [M]
  uint32_t* key;      // [2] carried key
  double* pbuf;       // pivot buffer: [0] s, [1] g_s, [2, 2+S) z_s, [2+S, 2+S+M) F[s, :]  (all-reduced)
  double* l1;         // [0] block b*, [1] prefix, [2] r, [3] total   (identical on every rank)
  double* bs_local;   // [nblk] this rank's block sums, zero elsewhere (send buffer)
  long long row0;     // global index of local row 0
  int blk0;           // global index of local block 0
  int nblk_loc;
  int partitionable;
  // external-column mode (rpcholesky_derivs): the kernel column of the pivot and the initial diagonal come
  // from buffers instead of the stationary kernel of Z
  const double* ext_col;  // [n] column of the pivot, refreshed by the caller's generator each step
  const double* diag0;    // [n] initial diagonal
  double* part;           // [ceil(M / CW)][nblk_loc * BLK] per-chunk partial dot products (rp_gemv_kernel)
  int peer_mode;          // 1: the pivot buffer is the peer mailbox, filled by the owner (non-owners must not touch it)
};

__host__ __device__ __forceinline__ long long fb_index(long long r, int c, int M) {
  return ((r >> 10) * M + c) * BLK + (r & (BLK - 1));
}

__global__ void rp_init_kernel(RpState st, double d0) {
  __shared__ double sh[32];
  const long long r = (long long)blockIdx.x * BLK + threadIdx.x;
  double v = 0.0;
  if (r < st.n) {
    v = st.diag0 ? st.diag0[r] : d0;
    st.diag[r] = v;
  }
  v = block_sum<BLK>(v, sh);
  if (threadIdx.x == 0) st.bs_local[st.blk0 + blockIdx.x] = v;
}

// One CTA, identical on every rank: draw u, scan the block sums, locate the block b* that holds the
// pivot.  Writes l1 = {b*, prefix, r, total}.
__device__ __forceinline__ void rp_pivot_l1(const RpState& st) {
  __shared__ double sh[32];
  __shared__ double s_total, s_r, s_prefix;
  __shared__ int s_cnt, s_bstar;
  const int tid = threadIdx.x;

  // total = sum(diag) from the block sums (fixed order)
  double t = 0.0;
  for (int b = tid; b < st.nblk; b += BLK) t += __ldcg(st.blocksum + b);
  t = block_sum<BLK>(t, sh);
  if (tid == 0) {
    s_total = t;
    // key, subkey = split(key); the draw uses the carried key (approximations.py:100-101)
    uint32_t k0, k1;
    jax_split2(st.key[0], st.key[1], 0, st.partitionable, k0, k1);
    st.key[0] = k0;
    st.key[1] = k1;
    const double u = jax_uniform_f64(k0, k1, 0ULL, 1ULL, st.partitionable);
    s_r = 1.0 - u;  // scaled by cum[-1] below
    s_cnt = 0;
  }
  __syncthreads();
  const double total = s_total;

  // first sweep over the per-block probabilities to get cum[-1]
  double carry = 0.0;
  for (int base = 0; base < st.nblk; base += BLK) {
    const int b = base + tid;
    double v = (b < st.nblk) ? __ldcg(st.blocksum + b) / total : 0.0;
    v = block_scan_1024(v, sh);
    __syncthreads();
    if (tid == BLK - 1) sh[0] = v;
    __syncthreads();
    carry += sh[0];
    __syncthreads();
  }
  const double rr = carry * s_r;
  // second sweep: count blocks whose cumulative probability is < r
  carry = 0.0;
  for (int base = 0; base < st.nblk; base += BLK) {
    const int b = base + tid;
    double v = (b < st.nblk) ? __ldcg(st.blocksum + b) / total : 0.0;
    v = block_scan_1024(v, sh) + carry;
    const int below = (b < st.nblk && v < rr) ? 1 : 0;
    const int wcnt = __popc(__ballot_sync(0xffffffffu, below));
    if ((tid & 31) == 0 && wcnt) atomicAdd(&s_cnt, wcnt);
    __syncthreads();
    if (tid == BLK - 1) sh[0] = v;
    __syncthreads();
    carry = sh[0];
    __syncthreads();
  }
  if (tid == 0) s_bstar = min(s_cnt, st.nblk - 1);
  __syncthreads();
  const int bstar = s_bstar;
  // prefix = cumulative probability of the blocks before bstar (recomputed in the same order)
  carry = 0.0;
  for (int base = 0; base <= bstar; base += BLK) {
    const int b = base + tid;
    double v = (b < st.nblk) ? __ldcg(st.blocksum + b) / total : 0.0;
    v = block_scan_1024(v, sh) + carry;
    if (b == bstar - 1) s_prefix = v;
    __syncthreads();
    if (tid == BLK - 1) sh[0] = v;
    __syncthreads();
    carry = sh[0];
    __syncthreads();
  }
  if (tid == 0) {
    st.l1[0] = (double)bstar;
    st.l1[1] = bstar > 0 ? s_prefix : 0.0;
    st.l1[2] = rr;
    st.l1[3] = total;
  }
}

// One CTA: the rank that owns block b* finds the pivot row inside it and fills the pivot buffer
// (s, g_s, z_s, F[s, :i]); every other rank zeroes its buffer so that a sum all-reduce
// distributes the owner's values.
__device__ __forceinline__ void rp_pivot_l2(const RpState& st, int i) {
  __shared__ double sh[32];
  __shared__ int s_cnt;
  const int tid = threadIdx.x;
  const int bstar = (int)st.l1[0];
  const int nbuf = 2 + st.S + i;
  if (bstar < st.blk0 || bstar >= st.blk0 + st.nblk_loc) {
    if (!st.peer_mode)
      for (int k = tid; k < nbuf; k += BLK) st.pbuf[k] = 0.0;
    return;
  }
  const double prefix = st.l1[1], rr = st.l1[2], total = st.l1[3];
  const long long r = (long long)(bstar - st.blk0) * BLK + tid;  // local row
  double pv = (r < st.n) ? st.diag[r] / total : 0.0;
  pv = block_scan_1024(pv, sh) + prefix;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  {
    const int below = (r < st.n && pv < rr) ? 1 : 0;
    const int wcnt = __popc(__ballot_sync(0xffffffffu, below));
    if ((tid & 31) == 0 && wcnt) atomicAdd(&s_cnt, wcnt);
  }
  __syncthreads();
  const long long nin = min((long long)BLK, (long long)st.n - (long long)(bstar - st.blk0) * BLK);
  const long long sl = (long long)(bstar - st.blk0) * BLK + min((long long)s_cnt, nin - 1);  // local pivot row

  double* zs = st.pbuf + 2;
  double* frow = st.pbuf + 2 + st.S;
  for (int d = tid; d < st.S; d += BLK) zs[d] = st.Z[sl * st.S + d];
  double acc = 0.0;
  for (int c = tid; c < i; c += BLK) {
    const double f = st.FB[fb_index(sl, c, st.M)];
    frow[c] = f;
    acc += f * f;
  }
  acc = block_sum<BLK>(acc, sh);
  if (tid == 0) {
    st.pbuf[0] = (double)(st.row0 + sl);
    st.pbuf[1] = (st.diag0 ? st.diag0[sl] : kfun(st.kind, 1e-12)) - acc;
  }
}

// One launch per pivot: level 1 (all ranks, identical), then level 2 (the owner of block b*; st.l1 was written by
// this CTA's thread 0, made visible by the barrier).
__global__ void __launch_bounds__(BLK, 1) rp_pivot_kernel(RpState st, int i) {
  rp_pivot_l1(st);
  __threadfence_block();
  __syncthreads();
  rp_pivot_l2(st, i);
}

// Multi-GPU pivot step in ONE launch over the peer mailboxes (comm.cuh), replacing
// [NCCL all-reduce of the block sums] -> [pivot kernel] -> [NCCL all-reduce of the pivot buffer]:
//   A. every rank stores its block sums into every rank's mailbox (its own range of the global array) and raises
//      its flag there; everyone waits for all flags: the global block sums are now local;
//   B. level 1 (identical on every rank) locates the pivot's block, the owner runs level 2 and fills the pivot
//      buffer (s, g_s, z_s, F[s, :i]) in its own mailbox;
//   C. the owner copies the pivot buffer into every peer's mailbox and raises the pivot flag; the others wait.
// st.blocksum / st.pbuf point into this rank's mailbox at the parity of `seq`.
__global__ void __launch_bounds__(BLK, 1) rp_pivot_exchange_kernel(RpState st, int i, PeerView pv, unsigned long long seq) {
  const int par = (int)(seq & 1ULL), tid = threadIdx.x, W = pv.world, me = pv.me;
  for (int r = 0; r < W; ++r) {
    double* dst = pv.base[r] + peer_off_bs(par) + st.blk0;
    for (int b = tid; b < st.nblk_loc; b += BLK) dst[b] = st.bs_local[st.blk0 + b];
  }
  __syncthreads();
  if (tid < W) {
    __threadfence_system();
    peer_st_release(peer_flags(pv.base[tid]) + peer_flag_bs(par, me), seq);
  }
  if (tid < W) peer_wait(peer_flags(pv.base[me]) + peer_flag_bs(par, tid), seq);
  __syncthreads();
  rp_pivot_l1(st);
  __threadfence_block();
  __syncthreads();
  const int bstar = (int)st.l1[0];
  const bool owner = bstar >= st.blk0 && bstar < st.blk0 + st.nblk_loc;
  if (owner) {
    rp_pivot_l2(st, i);
    __threadfence_block();
    __syncthreads();
    const int nbuf = 2 + st.S + i;
    for (int r = 0; r < W; ++r) {
      if (r == me) continue;
      double* dst = pv.base[r] + peer_off_pb(par);
      for (int k = tid; k < nbuf; k += BLK) dst[k] = st.pbuf[k];
    }
    __syncthreads();
    if (tid < W && tid != me) {
      __threadfence_system();
      peer_st_release(peer_flags(pv.base[tid]) + peer_flag_pb(par), seq);
    }
  } else {
    if (tid == 0) peer_wait(peer_flags(pv.base[me]) + peer_flag_pb(par), seq);
    __syncthreads();
  }
}

// One thread per data row: kernel column, GEMV against the pivot row, scale, diagonal
// update, block sum of the new diagonal.
// MINB: CTAs per SM ptxas must allow for.  2 (default) caps the kernel at 32 registers per thread and spills part of
// the 8 + 8 FP64 values of the inner loop (profiles/r1_sass_summary.txt); 1 (GPXB_RP_MINBLOCKS=1) gives 64 registers
// and no spill at half the resident threads -- kept switchable for the round-2 measurement.
template <int MINB>
__global__ void __launch_bounds__(BLK, MINB) rp_column_kernel(RpState st, int i) {
  extern __shared__ double sm[];  // frow[i], zs[S]
  __shared__ double sh[32];
  double* sf = sm;
  double* sz = sm + i;
  for (int c = threadIdx.x; c < i; c += BLK) sf[c] = st.pbuf[2 + st.S + c];
  for (int d = threadIdx.x; d < st.S; d += BLK) sz[d] = st.pbuf[2 + d];
  __syncthreads();
  const long long r = (long long)blockIdx.x * BLK + threadIdx.x;
  const long long s = (long long)st.pbuf[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) st.pivots[i] = s;
  double dnew = 0.0;
  if (r < st.n) {
    double g;
    if (st.ext_col) {
      g = st.ext_col[r];
    } else {
      const double* zr = st.Z + r * st.S;
      double dot = 0.0;
      for (int d = 0; d < st.Dp; ++d) dot += sz[d] * zr[d];
      double d2 = ((sz[st.Dp] - 2.0 * dot) + zr[st.Dp]) + 1e-12;
      if (r + st.row0 == s) d2 = 1e-12;
      g = kfun(st.kind, d2);
    }
    const double* f = st.FB + fb_index(r, 0, st.M);
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = 0.0;
    int c = 0;
    for (; c + 8 <= i; c += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(f + (long long)(c + u) * BLK);  // streaming: read once
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += v[u] * sf[c + u];
    }
    for (; c < i; ++c) a[0] += f[(long long)c * BLK] * sf[c];
    g -= ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    const double fnew = g / (sqrt(st.pbuf[1]) + 1e-6);
    st.FB[fb_index(r, i, st.M)] = fnew;
    dnew = st.diag[r] - fnew * fnew;
    st.diag[r] = dnew;
  }
  dnew = block_sum<BLK>(dnew, sh);
  if (threadIdx.x == 0) st.bs_local[st.blk0 + blockIdx.x] = dnew;
}

// ---- second generation of the column step (default) ---------------------------------------------------------
// The GEMV  g_r -= sum_c F[r][c] F[s][c]  is split into units of (1024-row block) x (CW pivot columns), one CTA
// each, so that the grid is thousands of equal units whatever the row count (the first generation ran one CTA per
// row block over ALL columns: 977 CTAs on 296 slots = 3.3 waves, a quarter of the machine idle in the last one, and
// on a row shard fewer CTAs than SMs).  512 threads, two adjacent rows per thread through 16-byte streaming loads,
// eight loads in flight per thread, no spill at 64 registers.  Every unit leaves one partial sum per row; the
// finishing kernel adds them in chunk order, so a row's arithmetic does not depend on the launch geometry or on
// the number of GPUs.
constexpr int CW = 128;  // pivot columns per unit (fixed: it defines the summation order)

__global__ void __launch_bounds__(BLK / 2, 2) rp_gemv_kernel(RpState st, int i) {
  __shared__ double sf[CW];
  const int c0 = blockIdx.y * CW, cn = min(CW, i - c0);
  if (threadIdx.x < cn) sf[threadIdx.x] = st.pbuf[2 + st.S + c0 + threadIdx.x];
  __syncthreads();
  const long long rb = (long long)blockIdx.x * BLK;  // first row of the block
  const double2* f = reinterpret_cast<const double2*>(st.FB + (rb * st.M + (long long)c0 * BLK)) + threadIdx.x;
  double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
  int c = 0;
  for (; c + 8 <= cn; c += 8) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(f + (long long)(c + u) * (BLK / 2));  // streaming: read once
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0[u & 3] = fma(v[u].x, sf[c + u], a0[u & 3]);
      a1[u & 3] = fma(v[u].y, sf[c + u], a1[u & 3]);
    }
  }
  if (c < cn) {  // ragged tail, same accumulator assignment (c is a multiple of 8 here)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c + u < cn) {
        const double2 v = __ldcs(f + (long long)(c + u) * (BLK / 2));
        a0[u & 3] = fma(v.x, sf[c + u], a0[u & 3]);
        a1[u & 3] = fma(v.y, sf[c + u], a1[u & 3]);
      }
    }
  }
  double2 o;
  o.x = (a0[0] + a0[1]) + (a0[2] + a0[3]);
  o.y = (a1[0] + a1[1]) + (a1[2] + a1[3]);
  reinterpret_cast<double2*>(st.part + ((long long)blockIdx.y * st.nblk_loc * BLK + rb))[threadIdx.x] = o;
}

// One thread per data row: kernel column, minus the partial dot products in chunk order, scale, store the new
// factor column, diagonal update, block sum of the new diagonal.
__global__ void __launch_bounds__(BLK, 1) rp_finish_kernel(RpState st, int i) {
  extern __shared__ double sz[];  // zs[S]
  __shared__ double sh[32];
  for (int d = threadIdx.x; d < st.S; d += BLK) sz[d] = st.pbuf[2 + d];
  __syncthreads();
  const long long r = (long long)blockIdx.x * BLK + threadIdx.x;
  const long long s = (long long)st.pbuf[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) st.pivots[i] = s;
  double dnew = 0.0;
  if (r < st.n) {
    double g;
    if (st.ext_col) {
      g = st.ext_col[r];
    } else {
      const double* zr = st.Z + r * st.S;
      double dot = 0.0;
      for (int d = 0; d < st.Dp; ++d) dot += sz[d] * zr[d];
      double d2 = ((sz[st.Dp] - 2.0 * dot) + zr[st.Dp]) + 1e-12;
      if (r + st.row0 == s) d2 = 1e-12;
      g = kfun(st.kind, d2);
    }
    const int nch = (i + CW - 1) / CW;
    const long long stride = (long long)st.nblk_loc * BLK;
    double acc = 0.0;
    for (int k = 0; k < nch; ++k) acc += st.part[k * stride + r];
    g -= acc;
    const double fnew = g / (sqrt(st.pbuf[1]) + 1e-6);
    st.FB[fb_index(r, i, st.M)] = fnew;
    dnew = st.diag[r] - fnew * fnew;
    st.diag[r] = dnew;
  }
  dnew = block_sum<BLK>(dnew, sh);
  if (threadIdx.x == 0) st.bs_local[st.blk0 + blockIdx.x] = dnew;
}

// F[r][c] (row-major N x M) <- blocked factor
__global__ void unblock_rowmajor_kernel(const double* __restrict__ FB, int n, int M, double* __restrict__ F) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + threadIdx.x;
    tile[k][threadIdx.x] = (c < M && r < n) ? FB[fb_index(r, c, M)] : 0.0;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + threadIdx.x;
    if (r < n && c < M) F[(long long)r * M + c] = tile[threadIdx.x][k];
  }
}

// FT[c][r] (M x ldf, one contiguous row per pivot) <- blocked factor
__global__ void unblock_transposed_kernel(const double* __restrict__ FB, int n, int M, double* __restrict__ FT,
                                          long long ldf) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * M) return;
  const int c = (int)(idx / n);
  const long long r = idx % n;
  FT[(long long)c * ldf + r] = FB[fb_index(r, c, M)];
}

}  // namespace

size_t rp_factor_doubles(int n, int M) { return (size_t)ceil_div(n, BLK) * (size_t)M * BLK; }

// Z: this rank's pre-scaled rows (n x S), global rows [row0, row0 + n) of Ntot (row0 a multiple of
// 1024).  FB: rp_factor_doubles(n, M) doubles (local rows).  pivots: [M] int64 device (global indices,
// identical on every rank).  With world > 1 two small all-reduces per pivot distribute the block
// sums and the pivot's rows.
static int rp_core(Ctx* ctx, int kind, const double* Z, int n, long long row0, long long Ntot, ZLayout zl,
                   const uint32_t key[2], int M, int partitionable, double* FB, long long* pivots,
                   const double* diag0, const double* ext_col, const RpColumnGen* gen, cudaStream_t s);

int rpcholesky_z(Ctx* ctx, int kind, const double* Z, int n, long long row0, long long Ntot, ZLayout zl,
                 const uint32_t key[2], int M, int partitionable, double* FB, long long* pivots, cudaStream_t s) {
  return rp_core(ctx, kind, Z, n, row0, Ntot, zl, key, M, partitionable, FB, pivots, nullptr, nullptr, nullptr, s);
}

// External-column variant (single rank): rows [n x S] are only copied into the pivot buffer (pbuf[2 + d]) for
// the generator, which must fill ext_col[n] with the pivot's column of the matrix before every column step.
int rpcholesky_ext(Ctx* ctx, const double* rows, int n, ZLayout zl, const double* diag0, double* ext_col,
                   const RpColumnGen& gen, const uint32_t key[2], int M, int partitionable, double* FB,
                   long long* pivots, cudaStream_t s) {
  GPXB_ARG_CHECK(diag0 && ext_col && gen, -42);
  return rp_core(ctx, KIND_SE, rows, n, 0, n, zl, key, M, partitionable, FB, pivots, diag0, ext_col, &gen, s);
}

static int rp_core(Ctx* ctx, int kind, const double* Z, int n, long long row0, long long Ntot, ZLayout zl,
                   const uint32_t key[2], int M, int partitionable, double* FB, long long* pivots,
                   const double* diag0, const double* ext_col, const RpColumnGen* gen, cudaStream_t s) {
  GPXB_ARG_CHECK(n >= 0 && M > 0 && M <= 6000, -40);  // frow + zs must fit in 48 KB of smem
  GPXB_ARG_CHECK(row0 % BLK == 0, -41);
  const int nblk = ceil_div(Ntot, BLK);
  const int nblk_loc = ceil_div(n, BLK);
  DevBuf bDiag, bBs, bBsl, bKey, bP, bL1;
  GPXB_TRY(bDiag.alloc(sizeof(double) * std::max(n, 1), s));
  GPXB_TRY(bBs.alloc(sizeof(double) * nblk, s));
  GPXB_TRY(bBsl.alloc(sizeof(double) * nblk, s));
  GPXB_TRY(bKey.alloc(sizeof(uint32_t) * 2, s));
  GPXB_TRY(bP.alloc(sizeof(double) * (2 + zl.S + M), s));
  GPXB_TRY(bL1.alloc(sizeof(double) * 4, s));
  RpState st;
  st.diag = bDiag.as<double>(); st.blocksum = bBs.as<double>(); st.bs_local = bBsl.as<double>();
  st.FB = FB; st.M = M; st.Z = Z; st.n = n; st.S = zl.S; st.Dp = zl.Dp; st.nblk = nblk; st.kind = kind;
  st.pivots = pivots; st.key = bKey.as<uint32_t>(); st.pbuf = bP.as<double>(); st.l1 = bL1.as<double>();
  st.row0 = row0; st.blk0 = (int)(row0 / BLK); st.nblk_loc = nblk_loc; st.partitionable = partitionable;
  st.ext_col = ext_col; st.diag0 = diag0; st.part = nullptr;
  const bool multi = ctx->world > 1 && !gen;  // the external-column variant runs replicated
  if (!multi) st.blocksum = st.bs_local;  // one buffer, no reduction
  st.peer_mode = 0;
  PeerView pv;
  const bool fused = multi && comm_peer_view(ctx, &pv) && nblk <= PEER_SLOT && 2 + zl.S + M <= PEER_SLOT;
  double* pbuf_nccl = st.pbuf;
  GPXB_CUDA_CHECK(cudaMemcpyAsync(st.key, key, sizeof(uint32_t) * 2, cudaMemcpyHostToDevice, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(st.bs_local, 0, sizeof(double) * nblk, s));
  // diag_i = k(x_i, x_i) via the batched kernel: d2 = 1e-12 (approximations.py:88-93)
  double d0;
  {
    const double d2 = 1e-12;
    if (kind == KIND_SE) d0 = std::exp(-d2);
    else {
      const double r = std::sqrt(d2);
      if (kind == KIND_M12) d0 = std::exp(-r);
      else if (kind == KIND_M32) { const double d = 1.7320508075688772 * r; d0 = (1.0 + d) * std::exp(-d); }
      else { const double d = 2.23606797749979 * r; d0 = (1.0 + d + d * d / 3.0) * std::exp(-d); }
    }
  }
  if (nblk_loc > 0) {
    rp_init_kernel<<<nblk_loc, BLK, 0, s>>>(st, d0);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
  }
  static const bool one_cta = [] {
    const char* e = getenv("GPXB_RP_MINBLOCKS");
    return e && atoi(e) == 1;
  }();
  // GPXB_RP_COLUMN=v1: the first-generation single-kernel column step (kept for comparison)
  static const bool gen1 = [] {
    const char* e = getenv("GPXB_RP_COLUMN");
    return e && e[0] == 'v' && e[1] == '1';
  }();
  DevBuf bPart;
  if (!gen1) {
    GPXB_TRY(bPart.alloc(sizeof(double) * (size_t)ceil_div(M, CW) * std::max(nblk_loc, 1) * BLK, s));
    st.part = bPart.as<double>();
  }
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)(one_cta ? rp_column_kernel<1> : rp_column_kernel<2>), (int)sizeof(double) * (M + zl.S)));
  // phase profiling brackets every 16th pivot only (weight 16): thousands of event pairs would perturb the loop
  struct PhaseSampling {
    Ctx* c; bool on; double w;
    explicit PhaseSampling(Ctx* ctx) : c(ctx), on(ctx->phase_prof), w(ctx->phase_weight) {}
    ~PhaseSampling() { c->phase_prof = on; c->phase_weight = w; }
  } sampling(ctx);
  const int every = M >= 64 ? 16 : 1;
  ctx->phase_weight = every;
  for (int i = 0; i < M; ++i) {
    ctx->phase_prof = sampling.on && (i % every == 0);
    if (fused) {
      // one launch: block sums and pivot buffer travel through the peer mailboxes (NVLink stores)
      PhaseScope ph(ctx, PH_RP_PIVOT, s);
      const unsigned long long seq = comm_peer_next_seq(ctx);
      const int par = (int)(seq & 1ULL);
      st.peer_mode = 1;
      st.blocksum = pv.base[pv.me] + peer_off_bs(par);
      st.pbuf = pv.base[pv.me] + peer_off_pb(par);
      rp_pivot_exchange_kernel<<<1, BLK, 0, s>>>(st, i, pv, seq);
      GPXB_LAUNCH_CHECK();
    } else {
      st.pbuf = pbuf_nccl;
      if (multi) GPXB_TRY(comm_allreduce_sum2(ctx, st.bs_local, bBs.as<double>(), nblk, s));
      {
        PhaseScope ph(ctx, PH_RP_PIVOT, s);
        rp_pivot_kernel<<<1, BLK, 0, s>>>(st, i);
        GPXB_LAUNCH_CHECK();
      }
      if (multi) GPXB_TRY(comm_allreduce_sum(ctx, st.pbuf, 2 + zl.S + i, s));
    }
    ctx->launches += 1;
    if (gen) GPXB_TRY((*gen)(st.pbuf, s));
    if (nblk_loc > 0 && !gen1) {
      PhaseScope ph(ctx, PH_RP_COLUMN, s);
      if (i > 0) {
        rp_gemv_kernel<<<dim3(nblk_loc, ceil_div(i, CW)), BLK / 2, 0, s>>>(st, i);
        GPXB_LAUNCH_CHECK();
        ctx->launches++;
      }
      rp_finish_kernel<<<nblk_loc, BLK, sizeof(double) * zl.S, s>>>(st, i);
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
    } else if (nblk_loc > 0) {
      PhaseScope ph(ctx, PH_RP_COLUMN, s);
      if (one_cta) rp_column_kernel<1><<<nblk_loc, BLK, sizeof(double) * (i + zl.S), s>>>(st, i);
      else rp_column_kernel<2><<<nblk_loc, BLK, sizeof(double) * (i + zl.S), s>>>(st, i);
      GPXB_LAUNCH_CHECK();
      ctx->launches++;
    }
  }
  return 0;
}

int rp_unblock_rowmajor(Ctx* ctx, const double* FB, int n, int M, double* F, cudaStream_t s) {
  dim3 grid(ceil_div(n, 32), ceil_div(M, 32)), blk(32, 8);
  unblock_rowmajor_kernel<<<grid, blk, 0, s>>>(FB, n, M, F);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

int rp_unblock_transposed(Ctx* ctx, const double* FB, int n, int M, double* FT, long long ldf, cudaStream_t s) {
  unblock_transposed_kernel<<<ceil_div((long long)n * M, 256), 256, 0, s>>>(FB, n, M, FT, ldf);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

}  // namespace gpxb
