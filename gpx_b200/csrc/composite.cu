// Fused Gram matrix of a Sum / Prod tree of kernels as ONE tile-level pass (gpx/kernels/kernels.py:1794-1848 Sum,
// :1851-1942 Prod, kernels/operations.py:103-426; leaves: the stationary kernels of gram.cu and Linear, kernels.py:226-236).
//
// Round 1 built every component matrix with the Gram kernel and combined them with element-wise passes: an n1 x n2
// matrix written and read again per leaf and per operator.  Here one CTA owns a 128 x 128 tile of the RESULT: the
// pre-scaled operand tiles of EVERY leaf arrive in shared memory by bulk TMA copies (each leaf has its own feature
// slice and lengthscale, hence its own operands), and for each 8 x 8 fragment the warp evaluates the postfix program
// of the tree -- leaf = DMMA dot over that leaf's features + nonlinearity, SUM / PROD = register arithmetic on a
// four-deep register stack -- and stores the result once.  Nothing but the result touches HBM.
#include "../../include/gpxb200.h"
#include "gram.cuh"
#include "kfun.cuh"
#include "mma.cuh"

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
            if (p.sym && gi == gj) d20 = 1e-12;
            if (p.sym && gi == gj + 1) d21 = 1e-12;
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
      if (gi == gj) v0 += p.diag_add;
      if (gi == gj + 1) v1 += p.diag_add;
      const bool ok0 = (gi < p.n1) && (gj < p.n2) && (!p.lower_only || gj <= gi);
      const bool ok1 = (gi < p.n1) && (gj + 1 < p.n2) && (!p.lower_only || gj + 1 <= gi);
      double* o = p.out + (long long)gi * p.ldo + gj;
      if (ok0) o[0] = v0;
      if (ok1) o[1] = v1;
    }
  }
}

}  // namespace

// Validates a postfix program over `nleaf` leaves: every leaf used, stack depth within CSTACK, one value left.
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
  GPXB_ARG_CHECK(ctx_ && kinds && ells && leaf_D && x1_leaf && x2_leaf && prog && out, -1);
  GPXB_ARG_CHECK(gpxb::check_program(nleaf, prog, nprog) == 0, -2);
  GPXB_ARG_CHECK(gpxb_gram_composite_fits(nleaf, leaf_D), -3);
  GPXB_ARG_CHECK(!lower_only || (sym && n1 == n2), -4);
  if (n1 <= 0 || n2 <= 0) return 0;
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t s = (cudaStream_t)stream;
  gpxb::CompParams p;
  p.nleaf = nleaf; p.nprog = nprog;
  for (int q = 0; q < nprog; ++q) p.prog[q] = prog[q];
  DevBuf z1[gpxb::CMAXL], z2[gpxb::CMAXL];
  int off = 0;
  for (int l = 0; l < nleaf; ++l) {
    GPXB_ARG_CHECK(kinds[l] >= 0 && kinds[l] <= gpxb::KIND_LINEAR && x1_leaf[l] && x2_leaf[l], -5);
    const gpxb::ZLayout zl = gpxb::z_layout(leaf_D[l]);
    const double ell = kinds[l] == gpxb::KIND_LINEAR ? 1.0 : ells[l];  // Linear: the raw features
    GPXB_TRY(z1[l].alloc(sizeof(double) * (size_t)n1 * zl.S, s));
    GPXB_TRY(gpxb::prep_z(ctx, x1_leaf[l], leaf_D[l], n1, zl, ell, z1[l].as<double>(), s));
    p.Z1[l] = z1[l].as<double>();
    if (sym) {
      p.Z2[l] = p.Z1[l];
    } else {
      GPXB_TRY(z2[l].alloc(sizeof(double) * (size_t)n2 * zl.S, s));
      GPXB_TRY(gpxb::prep_z(ctx, x2_leaf[l], leaf_D[l], n2, zl, ell, z2[l].as<double>(), s));
      p.Z2[l] = z2[l].as<double>();
    }
    p.kind[l] = kinds[l]; p.S[l] = zl.S; p.Dp[l] = zl.Dp; p.off[l] = off;
    off += 2 * gpxb::CTB * zl.S;
  }
  p.n1 = n1; p.n2 = n2; p.diag_add = diag_add; p.sym = sym ? 1 : 0; p.lower_only = lower_only ? 1 : 0;
  p.out = out; p.ldo = ldo;
  const long long tm = gpxb::ceil_div(n1, gpxb::CTB), tn = gpxb::ceil_div(n2, gpxb::CTB);
  p.tiles_n = (int)tn;
  const long long grid = lower_only ? tm * (tm + 1) / 2 : tm * tn;
  GPXB_ARG_CHECK(grid < (1LL << 31), -6);
  const int smem = 128 + off * (int)sizeof(double);
  GPXB_CUDA_CHECK(gpxb::ensure_dyn_smem((const void*)gpxb::gram_composite_kernel, smem));
  gpxb::gram_composite_kernel<<<(int)grid, 256, smem, s>>>(p);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

}  // extern "C"
