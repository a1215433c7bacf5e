// Fused Gram matrix of a Sum / Prod tree of kernels as ONE tile-level pass (gpx/kernels/kernels.py:1794-1848 Sum,
// :1851-1942 Prod, kernels/operations.py:103-426; leaves: the stationary kernels of gram.cu and Linear, kernels.py:226-236).
//
// Round 1 built every component matrix with the Gram kernel and combined them with element-wise passes: an n1 x n2
// matrix written and read again per leaf and per operator.  Here one CTA owns a 128 x 128 tile of the RESULT: the
// pre-scaled operand tiles of EVERY leaf arrive in shared memory by bulk TMA copies (each leaf has its own feature
// slice and lengthscale, hence its own operands), and for each 8 x 8 fragment the warp evaluates the postfix program
// of the tree -- leaf = DMMA dot over that leaf's features + nonlinearity, SUM / PROD = register arithmetic on a
// four-deep register stack -- and stores the result once.  Nothing but the result touches HBM.
#include <functional>

#include "../../include/gpxb200.h"
#include "gemm.cuh"
#include "gram.cuh"
#include "kfun.cuh"
#include "matvec.cuh"
#include "mma.cuh"
#include "rpchol.cuh"
#include "small_kernels.cuh"

namespace gpxb {

namespace {

constexpr int CTB = 128;      // tile edge
constexpr int CMAXL = 8;      // leaves
constexpr int CMAXP = 15;     // program length (8 leaves + 7 operators)
constexpr int CSTACK = 4;     // register stack depth (a balanced tree of 8 leaves needs 4)
constexpr int KIND_LINEAR = 4;
constexpr int OP_SUM = -1, OP_PROD = -2;

struct CompParams {
  int nleaf, nprog;
  int kind[CMAXL], S[CMAXL], Dp[CMAXL], off[CMAXL];  // off: doubles from the start of the operand area to the leaf's Z1 tile
  const double* Z1[CMAXL];
  const double* Z2[CMAXL];
  int prog[CMAXP];
  int n1, n2, tiles_n;
  double diag_add;
  int sym, lower_only;
  long long doff;  // the diagonal is where row + doff == col (the row operand is a row range of the column operand)
  double* out;
  long long ldo;
};

__global__ void __launch_bounds__(256, 1) gram_composite_kernel(const CompParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sZ = reinterpret_cast<double*>(smem_raw + 128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane >> 2, lq = lane & 3;

  int ti, tj;
  if (p.lower_only) {
    const long long t = blockIdx.x;
    long long i = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (i * (i + 1) / 2 > t) --i;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    ti = (int)i;
    tj = (int)(t - i * (i + 1) / 2);
  } else {
    ti = blockIdx.x / p.tiles_n;
    tj = blockIdx.x % p.tiles_n;
  }
  const int row0 = ti * CTB, col0 = tj * CTB;
  const int rows1 = min(CTB, p.n1 - row0), rows2 = min(CTB, p.n2 - col0);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t tot = 0;
    for (int l = 0; l < p.nleaf; ++l) tot += (uint32_t)(rows1 + rows2) * p.S[l] * 8u;
    mbar_arrive_expect_tx(bar, tot);
    for (int l = 0; l < p.nleaf; ++l) {
      double* d1 = sZ + p.off[l];
      tma_bulk_g2s(d1, p.Z1[l] + (long long)row0 * p.S[l], (uint32_t)rows1 * p.S[l] * 8u, bar);
      tma_bulk_g2s(d1 + CTB * p.S[l], p.Z2[l] + (long long)col0 * p.S[l], (uint32_t)rows2 * p.S[l] * 8u, bar);
    }
  }
  mbar_wait(bar, 0);

#pragma unroll 1
  for (int mf = 0; mf < 8; ++mf) {
    const int rloc = wm * 64 + mf * 8 + lr;
    const int gi = row0 + rloc;
#pragma unroll 1
    for (int nf = 0; nf < 4; ++nf) {
      const int cfrag = wn * 32 + nf * 8;          // first local column of the fragment
      const int cloc = cfrag + 2 * lq;             // this lane's two columns
      const int gj = col0 + cloc;
      double r0[CSTACK], r1[CSTACK];               // register stack (indexed through uniform switches only)
#pragma unroll
      for (int k = 0; k < CSTACK; ++k) r0[k] = r1[k] = 0.0;
      int sp = 0;
#pragma unroll 1
      for (int q = 0; q < p.nprog; ++q) {
        const int t = p.prog[q];
        if (t >= 0) {
          const int S = p.S[t], Dp = p.Dp[t], kind = p.kind[t];
          const double* z1 = sZ + p.off[t];
          const double* z2 = z1 + CTB * S;
          const double* ap = z1 + rloc * S + lq;
          const double* bp = z2 + (cfrag + lr) * S + lq;
          double c0 = 0.0, c1 = 0.0;
          for (int k = 0; k < Dp; k += 4) dmma884(c0, c1, ap[k], bp[k]);
          double v0 = c0, v1 = c1;                 // Linear: the raw dot product
          if (kind != KIND_LINEAR) {
            const double ni = z1[rloc * S + Dp], nj0 = z2[cloc * S + Dp], nj1 = z2[(cloc + 1) * S + Dp];
            double d20 = ((ni - 2.0 * c0) + nj0) + 1e-12, d21 = ((ni - 2.0 * c1) + nj1) + 1e-12;
            if (p.sym && gi + p.doff == gj) d20 = 1e-12;
            if (p.sym && gi + p.doff == gj + 1) d21 = 1e-12;
            v0 = kfun_fast(kind, d20);
            v1 = kfun_fast(kind, d21);
          }
          switch (sp) {  // push (uniform over the warp: no divergence, the stack stays in registers)
            case 0: r0[0] = v0; r1[0] = v1; break;
            case 1: r0[1] = v0; r1[1] = v1; break;
            case 2: r0[2] = v0; r1[2] = v1; break;
            default: r0[3] = v0; r1[3] = v1; break;
          }
          ++sp;
        } else {
          const bool sum = (t == OP_SUM);
          switch (sp) {  // the two topmost entries are combined into the lower one
            case 2:
              r0[0] = sum ? r0[0] + r0[1] : r0[0] * r0[1];
              r1[0] = sum ? r1[0] + r1[1] : r1[0] * r1[1];
              break;
            case 3:
              r0[1] = sum ? r0[1] + r0[2] : r0[1] * r0[2];
              r1[1] = sum ? r1[1] + r1[2] : r1[1] * r1[2];
              break;
            default:
              r0[2] = sum ? r0[2] + r0[3] : r0[2] * r0[3];
              r1[2] = sum ? r1[2] + r1[3] : r1[2] * r1[3];
              break;
          }
          --sp;
        }
      }
      double v0 = r0[0], v1 = r1[0];
      if (gi + p.doff == gj) v0 += p.diag_add;
      if (gi + p.doff == gj + 1) v1 += p.diag_add;
      const bool ok0 = (gi < p.n1) && (gj < p.n2) && (!p.lower_only || gj <= gi);
      const bool ok1 = (gi < p.n1) && (gj + 1 < p.n2) && (!p.lower_only || gj + 1 <= gi);
      double* o = p.out + (long long)gi * p.ldo + gj;
      if (ok0) o[0] = v0;
      if (ok1) o[1] = v1;
    }
  }
}

}  // namespace

// ---- host-side descriptor of a composite kernel over per-leaf feature slices ---------------------------------
struct CompDesc {
  int nleaf = 0, nprog = 0;
  int kind[CMAXL], D[CMAXL];
  double ell[CMAXL];
  const double* x[CMAXL];  // per leaf: n x D[l] row-major (already sliced to the leaf's active dimensions)
  int prog[CMAXP];
};

// Validates a postfix program over `nleaf` leaves: stack depth within CSTACK, one value left.
static int check_program(int nleaf, const int* prog, int nprog) {
  if (nleaf < 1 || nleaf > CMAXL || nprog < 1 || nprog > CMAXP) return -1;
  int sp = 0;
  for (int q = 0; q < nprog; ++q) {
    if (prog[q] >= 0) {
      if (prog[q] >= nleaf) return -1;
      if (++sp > CSTACK) return -2;
    } else if (prog[q] == OP_SUM || prog[q] == OP_PROD) {
      if (sp < 2) return -1;
      --sp;
    } else {
      return -1;
    }
  }
  return sp == 1 ? 0 : -1;
}

static int fill_desc(CompDesc& d, int nleaf, const int* kinds, const double* ells, const int* leaf_D,
                     const double* const* x_leaf, const int* prog, int nprog) {
  GPXB_ARG_CHECK(kinds && ells && leaf_D && x_leaf && prog, -1);
  GPXB_ARG_CHECK(check_program(nleaf, prog, nprog) == 0, -2);
  d.nleaf = nleaf; d.nprog = nprog;
  for (int q = 0; q < nprog; ++q) d.prog[q] = prog[q];
  for (int l = 0; l < nleaf; ++l) {
    GPXB_ARG_CHECK(kinds[l] >= 0 && kinds[l] <= KIND_LINEAR && leaf_D[l] >= 1 && x_leaf[l], -5);
    d.kind[l] = kinds[l]; d.D[l] = leaf_D[l]; d.x[l] = x_leaf[l];
    d.ell[l] = kinds[l] == KIND_LINEAR ? 1.0 : ells[l];
  }
  return 0;
}

static bool desc_fits(const CompDesc& d) {
  long long dbl = 0;
  for (int l = 0; l < d.nleaf; ++l) dbl += 2LL * CTB * z_layout(d.D[l]).S;
  return 128 + dbl * 8 <= 232448;
}

// Pre-scaled operands of every leaf for n rows of d.x (starting at row r0)
struct CompOperands {
  DevBuf z[CMAXL];
  int prepare(Ctx* ctx, const CompDesc& d, long long r0, int n, cudaStream_t s) {
    for (int l = 0; l < d.nleaf; ++l) {
      const ZLayout zl = z_layout(d.D[l]);
      GPXB_TRY(z[l].alloc(sizeof(double) * (size_t)std::max(n, 1) * zl.S, s));
      GPXB_TRY(prep_z(ctx, d.x[l] + r0 * d.D[l], d.D[l], n, zl, d.ell[l], z[l].as<double>(), s));
    }
    return 0;
  }
};

// out[n1 x n2] = composite Gram of the prepared operands (Z2 == nullptr: symmetric, Z1 against itself)
static int composite_gram_launch(Ctx* ctx, const CompDesc& d, const CompOperands& a, int n1, const CompOperands* b,
                                 int n2, double diag_add, int sym, long long diag_off, int lower_only, double* out,
                                 long long ldo, cudaStream_t s) {
  if (n1 <= 0 || n2 <= 0) return 0;
  GPXB_ARG_CHECK(desc_fits(d), -3);
  CompParams p;
  p.nleaf = d.nleaf; p.nprog = d.nprog;
  for (int q = 0; q < d.nprog; ++q) p.prog[q] = d.prog[q];
  int off = 0;
  for (int l = 0; l < d.nleaf; ++l) {
    const ZLayout zl = z_layout(d.D[l]);
    p.Z1[l] = a.z[l].as<double>();
    p.Z2[l] = b ? b->z[l].as<double>() : p.Z1[l];
    p.kind[l] = d.kind[l]; p.S[l] = zl.S; p.Dp[l] = zl.Dp; p.off[l] = off;
    off += 2 * CTB * zl.S;
  }
  p.n1 = n1; p.n2 = n2; p.diag_add = diag_add; p.sym = sym ? 1 : 0; p.lower_only = lower_only ? 1 : 0;
  p.doff = diag_off; p.out = out; p.ldo = ldo;
  const long long tm = ceil_div(n1, CTB), tn = ceil_div(n2, CTB);
  p.tiles_n = (int)tn;
  const long long grid = lower_only ? tm * (tm + 1) / 2 : tm * tn;
  GPXB_ARG_CHECK(grid < (1LL << 31), -6);
  const int smem = 128 + off * (int)sizeof(double);
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)gram_composite_kernel, smem));
  gram_composite_kernel<<<(int)grid, 256, smem, s>>>(p);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

// ---- composite kernels in the matrix-free paths ---------------------------------------------------------------
// The operator W = (K + diag_add I) V in row chunks: fused composite Gram of (chunk rows) x (all rows) into a buffer
// of at most ~1 GB, then one DMMA GEMM with the block V -- K is never held as a whole.  Several leaves make the
// per-entry kernel cost (sum of the leaves' 2 D + nonlinearity) comparable to a large-D stationary kernel, for
// which the same chunked form is used (matvec.cu: kmatvec_chunked).
struct CompOp {
  Ctx* ctx = nullptr;
  CompDesc d;
  int N = 0;
  double diag_add = 0.0;
  CompOperands all;
  DevBuf kc;
  long long ldk = 0, rc_max = 0;
  int prepare(Ctx* c, const CompDesc& desc, int n, double shift, cudaStream_t s) {
    ctx = c; d = desc; N = n; diag_add = shift;
    GPXB_TRY(all.prepare(ctx, d, 0, N, s));
    ldk = round_up((long long)N, 2);
    rc_max = std::max<long long>(128, (1LL << 27) / std::max<long long>(ldk, 1) / 128 * 128);
    rc_max = std::min<long long>(rc_max, round_up((long long)N, 128));
    return kc.alloc(sizeof(double) * (size_t)rc_max * ldk, s);
  }
  // W[N x ldw] (P columns) = (K + diag_add I) V, V: N x ldv
  int apply(const double* V, long long ldv, int P, double* W, long long ldw, cudaStream_t s) {
    for (long long r0 = 0; r0 < N; r0 += rc_max) {
      const int rc = (int)std::min<long long>(rc_max, N - r0);
      GPXB_TRY(launch_chunk(r0, rc, s));
      GemmArgs m;
      m.M = rc; m.N = P; m.K = N;
      m.A = kc.as<double>(); m.lda = ldk; m.a_kmajor = true;
      m.B = V; m.ldb = ldv; m.b_kmajor = false;
      m.C = W + r0 * ldw; m.ldc = ldw;
      GPXB_TRY(gemm_f64(ctx, m, s));
    }
    return 0;
  }
  int launch_chunk(long long r0, int rc, cudaStream_t s) {
    CompParams p;
    p.nleaf = d.nleaf; p.nprog = d.nprog;
    for (int q = 0; q < d.nprog; ++q) p.prog[q] = d.prog[q];
    int off = 0;
    for (int l = 0; l < d.nleaf; ++l) {
      const ZLayout zl = z_layout(d.D[l]);
      p.Z2[l] = all.z[l].as<double>();
      p.Z1[l] = p.Z2[l] + r0 * zl.S;
      p.kind[l] = d.kind[l]; p.S[l] = zl.S; p.Dp[l] = zl.Dp; p.off[l] = off;
      off += 2 * CTB * zl.S;
    }
    p.n1 = rc; p.n2 = N; p.diag_add = diag_add; p.sym = 1; p.lower_only = 0; p.doff = r0;
    p.out = kc.as<double>(); p.ldo = ldk;
    p.tiles_n = (int)ceil_div(N, CTB);
    const long long grid = (long long)ceil_div(rc, CTB) * p.tiles_n;
    GPXB_ARG_CHECK(grid < (1LL << 31), -6);
    const int smem = 128 + off * (int)sizeof(double);
    GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)gram_composite_kernel, smem));
    gram_composite_kernel<<<(int)grid, 256, smem, s>>>(p);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
    return 0;
  }
};

namespace {
// value of the tree for one pair of points: x_r against the pivot x_s (one thread per row r)
__device__ double composite_pair(const CompParams& p, const double* const* xl, const int* Dl, const double* ell,
                                 long long r, long long sidx) {
  double st[CSTACK + 1];
  int sp = 0;
  for (int q = 0; q < p.nprog; ++q) {
    const int t = p.prog[q];
    if (t >= 0) {
      const int D = Dl[t];
      const double* a = xl[t] + r * D;
      const double* b = xl[t] + sidx * D;
      double v;
      if (p.kind[t] == KIND_LINEAR) {
        v = 0.0;
        for (int f = 0; f < D; ++f) v += a[f] * b[f];
      } else {
        // the matrix kernel's distance: |z_r|^2 - 2 z_r . z_s + |z_s|^2 + 1e-12 with z = x / ell (gpx/utils.py:55-72)
        double nr = 0.0, ns = 0.0, dot = 0.0;
        for (int f = 0; f < D; ++f) {
          const double zr = a[f] / ell[t], zs = b[f] / ell[t];
          nr += zr * zr; ns += zs * zs; dot += zr * zs;
        }
        double d2 = ((nr - 2.0 * dot) + ns) + 1e-12;
        if (r == sidx) d2 = 1e-12;
        v = kfun_fast(p.kind[t], d2);
      }
      st[sp++] = v;
    } else {
      const double y = st[--sp], x = st[--sp];
      st[sp++] = (t == OP_SUM) ? x + y : x * y;
    }
  }
  return st[0];
}

struct CompColArgs {
  const double* x[CMAXL];
  int D[CMAXL];
  double ell[CMAXL];
};

// col[r] = k(x_r, x_pivot) (pivot index in pbuf[0]); diag mode (pbuf == nullptr): col[r] = k(x_r, x_r)
__global__ void composite_column_kernel(const CompParams p, const CompColArgs a, long long n,
                                        const double* __restrict__ pbuf, double* __restrict__ col) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const long long sidx = pbuf ? (long long)pbuf[0] : r;
  col[r] = composite_pair(p, a.x, a.D, a.ell, r, sidx);
}
}  // namespace

// randomly pivoted Cholesky of a composite kernel (gpx/kernels/approximations.py:30-129 with a Sum / Prod kernel):
// external-column mode of the pivot loop, the column of the pivot evaluated pair by pair.
static int composite_rpcholesky(Ctx* ctx, const CompDesc& d, int n, const uint32_t key[2], int M, int partitionable,
                                double* FB, long long* pivots, cudaStream_t s) {
  CompParams p;
  p.nleaf = d.nleaf; p.nprog = d.nprog;
  for (int q = 0; q < d.nprog; ++q) p.prog[q] = d.prog[q];
  CompColArgs a;
  for (int l = 0; l < d.nleaf; ++l) {
    p.kind[l] = d.kind[l];
    a.x[l] = d.x[l]; a.D[l] = d.D[l]; a.ell[l] = d.ell[l];
  }
  const ZLayout zl = z_layout(1);
  DevBuf bRows, bD0, bCol;
  GPXB_TRY(bRows.alloc(sizeof(double) * (size_t)n * zl.S, s));
  GPXB_TRY(bD0.alloc(sizeof(double) * (size_t)n, s));
  GPXB_TRY(bCol.alloc(sizeof(double) * (size_t)n, s));
  GPXB_CUDA_CHECK(cudaMemsetAsync(bRows.as<double>(), 0, sizeof(double) * (size_t)n * zl.S, s));
  double* col = bCol.as<double>();
  composite_column_kernel<<<ceil_div(n, 256), 256, 0, s>>>(p, a, n, nullptr, bD0.as<double>());
  GPXB_LAUNCHED(ctx);
  RpColumnGen gen = [&](const double* pbuf, cudaStream_t st) -> int {
    composite_column_kernel<<<ceil_div(n, 256), 256, 0, st>>>(p, a, n, pbuf, col);
    GPXB_LAUNCHED(ctx);
    return 0;
  };
  return rpcholesky_ext(ctx, bRows.as<double>(), n, zl, bD0.as<double>(), col, gen, key, M, partitionable, FB, pivots, s);
}

int precond_build(Ctx* ctx, const double* F, long long ldf, int k, int n_loc, double sigma, double* Pkk,
                  cudaStream_t s);
int pcg_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
             const double* b, const double* F, long long ldf, int k, const double* Pkk, double sigma, double tol,
             double atol, int maxiter, double* x, int* iters_out, cudaStream_t s);
int lanczos_core(Ctx* ctx, int n_loc, long long N, const std::function<int(const double*, double*)>& apply,
                 const double* U, int P, int PS, int m, double* alpha_out, double* beta_out, int* flag,
                 cudaStream_t s);

}  // namespace gpxb

using gpxb::Ctx;
using gpxb::DevBuf;

extern "C" {

int gpxb_gram_composite_fits(int nleaf, const int* leaf_D) {
  if (nleaf < 1 || nleaf > gpxb::CMAXL || !leaf_D) return 0;
  long long dbl = 0;
  for (int l = 0; l < nleaf; ++l) {
    if (leaf_D[l] < 1) return 0;
    dbl += 2LL * gpxb::CTB * gpxb::z_layout(leaf_D[l]).S;
  }
  return 128 + dbl * 8 <= 232448 ? 1 : 0;
}

int gpxb_gram_composite(gpxb_ctx* ctx_, int nleaf, const int* kinds, const double* ells, const int* leaf_D,
                        const double* const* x1_leaf, const double* const* x2_leaf, int n1, int n2, const int* prog,
                        int nprog, double diag_add, int sym, int lower_only, double* out, long long ldo,
                        gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && x2_leaf && out, -1);
  gpxb::CompDesc d1, d2;
  GPXB_TRY(gpxb::fill_desc(d1, nleaf, kinds, ells, leaf_D, x1_leaf, prog, nprog));
  GPXB_TRY(gpxb::fill_desc(d2, nleaf, kinds, ells, leaf_D, x2_leaf, prog, nprog));
  GPXB_ARG_CHECK(gpxb::desc_fits(d1), -3);
  GPXB_ARG_CHECK(!lower_only || (sym && n1 == n2), -4);
  if (n1 <= 0 || n2 <= 0) return 0;
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  gpxb::CompOperands a, b;
  GPXB_TRY(a.prepare(ctx, d1, 0, n1, s));
  if (!sym) GPXB_TRY(b.prepare(ctx, d2, 0, n2, s));
  return gpxb::composite_gram_launch(ctx, d1, a, n1, sym ? nullptr : &b, n2, diag_add, sym, 0, lower_only, out, ldo, s);
}

int gpxb_rpcholesky_composite(gpxb_ctx* ctx_, int nleaf, const int* kinds, const double* ells, const int* leaf_D,
                              const double* const* x_leaf, int N, const int* prog, int nprog, uint32_t key0,
                              uint32_t key1, int M, int partitionable, double* F_out, long long* pivots_out,
                              gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && pivots_out && N > 0 && M > 0, -1);
  gpxb::CompDesc d;
  GPXB_TRY(gpxb::fill_desc(d, nleaf, kinds, ells, leaf_D, x_leaf, prog, nprog));
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  GPXB_ARG_CHECK(ctx->world == 1, -61);  // replicas only
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf fb;
  GPXB_TRY(fb.alloc(sizeof(double) * std::max<size_t>(gpxb::rp_factor_doubles(N, M), 8), s));
  const uint32_t key[2] = {key0, key1};
  GPXB_TRY(gpxb::composite_rpcholesky(ctx, d, N, key, M, partitionable, fb.as<double>(), pivots_out, s));
  if (F_out) GPXB_TRY(gpxb::rp_unblock_rowmajor(ctx, fb.as<double>(), N, M, F_out, s));
  return 0;
}

int gpxb_gpr_iter_composite(gpxb_ctx* ctx_, int nleaf, const int* kinds, const double* ells, const int* leaf_D,
                            const double* const* x_leaf, int N, const int* prog, int nprog, const double* y,
                            double sigma, int mean_mode, int n_pivots, uint32_t key0, uint32_t key1, int partitionable,
                            double tol, double atol, int maxiter, double* c_out, double* mu_out, int* iters_out,
                            const double* U, long long ldu, int P, int m, double* alpha_out, double* beta_out,
                            int* breakdown_flag, gpxb_stream stream) {
  GPXB_ARG_CHECK(ctx_ && N > 0 && n_pivots >= 0, -1);
  GPXB_ARG_CHECK((c_out && y && mu_out) || (U && alpha_out && beta_out && breakdown_flag), -1);
  gpxb::CompDesc d;
  GPXB_TRY(gpxb::fill_desc(d, nleaf, kinds, ells, leaf_D, x_leaf, prog, nprog));
  GPXB_ARG_CHECK(gpxb::desc_fits(d), -3);
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  GPXB_ARG_CHECK(ctx->world == 1, -61);  // replicas only
  cudaStream_t s = (cudaStream_t)stream;
  const double s2 = sigma * sigma + 1e-10;
  gpxb::CompOp op;
  GPXB_TRY(op.prepare(ctx, d, N, s2, s));
  if (c_out) {
    DevBuf r, fb, ft, piv, pkk;
    GPXB_TRY(r.alloc(sizeof(double) * N, s));
    gpxb::sk::mean_kernel<<<1, 1024, 0, s>>>(y, (long long)N, mean_mode == GPXB_MEAN_DATA, mu_out);
    GPXB_LAUNCHED(ctx);
    gpxb::sk::center_kernel<<<gpxb::ceil_div(N, 256), 256, 0, s>>>(y, N, mu_out, r.as<double>());
    GPXB_LAUNCHED(ctx);
    const long long ldf = gpxb::round_up(N, 2);
    if (n_pivots > 0) {
      GPXB_TRY(fb.alloc(sizeof(double) * std::max<size_t>(gpxb::rp_factor_doubles(N, n_pivots), 8), s));
      GPXB_TRY(ft.alloc(sizeof(double) * (size_t)n_pivots * ldf, s));
      GPXB_TRY(piv.alloc(sizeof(long long) * n_pivots, s));
      GPXB_TRY(pkk.alloc(sizeof(double) * (size_t)n_pivots * n_pivots, s));
      const uint32_t key[2] = {key0, key1};
      GPXB_TRY(gpxb::composite_rpcholesky(ctx, d, N, key, n_pivots, partitionable, fb.as<double>(),
                                          piv.as<long long>(), s));
      GPXB_TRY(gpxb::rp_unblock_transposed(ctx, fb.as<double>(), N, n_pivots, ft.as<double>(), ldf, s));
      fb.release();
      GPXB_TRY(gpxb::precond_build(ctx, ft.as<double>(), ldf, n_pivots, N, sigma, pkk.as<double>(), s));
    }
    auto apply = [&](const double* p, double* Ap) -> int { return op.apply(p, 2, 1, Ap, 1, s); };
    GPXB_TRY(gpxb::pcg_core(ctx, N, N, apply, r.as<double>(), n_pivots > 0 ? ft.as<double>() : nullptr, ldf, n_pivots,
                            pkk.as<double>(), sigma, tol, atol, maxiter > 0 ? maxiter : 10 * N, c_out, iters_out, s));
  }
  if (U && alpha_out) {
    GPXB_CUDA_CHECK(cudaMemsetAsync(breakdown_flag, 0, sizeof(int), s));
    DevBuf up;
    for (int c0 = 0; c0 < P; c0 += 32) {
      const int pb = std::min(32, P - c0);
      const int PS = gpxb::v_padded_stride(pb);
      GPXB_TRY(up.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
      GPXB_TRY(gpxb::pack_v(ctx, U, ldu, N, c0, pb, PS, up.as<double>(), s));
      auto apply = [&](const double* v, double* w) -> int { return op.apply(v, PS, pb, w, PS, s); };
      GPXB_TRY(gpxb::lanczos_core(ctx, N, N, apply, up.as<double>(), pb, PS, m, alpha_out + (size_t)c0 * m,
                                  beta_out + (size_t)c0 * m, breakdown_flag, s));
      up.release();
    }
  }
  return 0;
}

}  // extern "C"
