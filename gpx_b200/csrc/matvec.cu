// Family 3a: matrix-free block mat-vec  W = (K(x_rows, x_all) + diag_add*I) V  for sm_100a.
//
// Replaces gpx/operations.py:54-123 (rowfun_to_matvec: one kernel row per lax.scan step)
// driven by _Ax_lhs_fun (gpx/models/_gpr.py:94-119) and update_row_diagonal
// (gpx/operations.py:32-51).  K is never materialised: every 256 x 128 tile of it is
// recomputed on the fly and consumed from registers.
//
// One CTA owns 256 rows (8 warps x 32 rows) and streams over the columns j in tiles of 128:
//   - the pre-scaled column tile of Z and the matching rows of V arrive by 1-D bulk TMA
//     copies (cp.async.bulk + mbarrier), double buffered;
//   - GEMM 1 (DMMA, K = D):   dot = Z_rows . Z_cols^T   (8x8 fragments)
//   - epilogue in registers:  k = phi(|z_i|^2 - 2 dot + |z_j|^2 + 1e-12)
//   - GEMM 2 (DMMA, K = 128): W_rows += k . V_cols.  The 8x8 accumulator fragment of GEMM 1
//     is fed straight back as two A operands of GEMM 2 (columns {0,2,4,6} and {1,3,5,7}: the
//     contraction index may be permuted as long as A and B agree), so no shuffle is needed.
// V is stored with a padded row stride PS (PS % 8 == 2) so that a tile is contiguous (one
// bulk copy) and the GEMM-2 B-fragment reads are bank-conflict free.
#include "kfun.cuh"
#include "gemm.cuh"
#include "gram.cuh"
#include "matvec.cuh"
#include "mma.cuh"

#include <cmath>
#include <cstdlib>

namespace gpxb {

namespace {

constexpr int BJ = 128;

struct KmvParams {
  int kind;
  const double* Zr;
  int n_rows;
  long long row_offset;  // global index (in x_all) of row 0, or -1: no diagonal treatment
  const double* Za;
  int N;
  int S, Dp;
  const double* V;
  int PS, P;
  double* W;  // jsplit == 1: [n_rows][ldw]; else partial [jsplit][n_rows][8*NPF]
  long long ldw;
  double diag_add;
  int jsplit, jps;  // columns per split (multiple of BJ)
  // KMODE 3 (weighted derivative profile): k_ij <- g(d2_ij) * (wrow[i] * wcol[j] - wt(i, j)),
  // wt(i, j) = wt[i * ldwt + j] or, transposed storage, wt[j * ldwt + i]
  const double* wrow;
  const double* wcol;
  const double* wt;
  long long ldwt;
  int wt_trans;
  double rl;  // KMODE 4 (d K / d lengthscale): 1 / lengthscale
};

// MF: 8-row fragments per warp (rows per CTA = 64*MF); KS: Dp/4 when the row operand lives in
// registers, 0 = read it from shared memory; NPF: 8-column fragments of V/W (P <= 8*NPF).
// KMODE: 0 = kind chosen at run time (table exp), 1 = SquaredExponential with libm exp (GPXB_KMV_EXP=libm),
//        2 = SquaredExponential fast path: the distance FMA produces u = -d2 * 256 log2(e) directly and
//            exp(-d2) = 2^(u/256) by exp2_tab256 (10.25 FP64 ops per entry in all),
//        3 = radial derivative profile g of the kind chosen at run time, d k/d x1 = -c(l) g (x1 - x2):
//            SE exp(-d2) [c = 2/l^2], M12 exp(-r)/r [1/l^2], M32 exp(-d) [3/l^2], M52 (1+d) exp(-d) [5/(3 l^2)]
//        4 = d k/d lengthscale (kfun_dl) of the kind chosen at run time, evaluated in the loop body
template <int KMODE>
__device__ __forceinline__ double kfun_mode(int kind, double d2, const double* __restrict__ tab) {
  if (KMODE == 1) return exp(-d2);
  // exp(-t) = 2^(-t * 256 log2(e) / 256) through the shared 256-entry table (kfun.cuh: exp2_tab256)
  auto expn = [&](double t) { return exp2_tab256(-EXP2_256_SCALE * t, tab); };
  if (KMODE == 3) {
    if (kind == KIND_SE) return expn(d2);
    const double r = sqrt(fmax(d2, 1e-36));
    if (kind == KIND_M12) return expn(r) / r;
    if (kind == KIND_M32) return expn(1.7320508075688772 * r);
    const double d = 2.23606797749979 * r;
    return (1.0 + d) * expn(d);
  }
  if (kind == KIND_SE) return expn(d2);
  const double r = sqrt(fmax(d2, 1e-36));
  if (kind == KIND_M12) return expn(r);
  if (kind == KIND_M32) {
    const double d = 1.7320508075688772 * r;
    return (1.0 + d) * expn(d);
  }
  const double d = 2.23606797749979 * r;
  return (1.0 + d + d * d * 0.3333333333333333) * expn(d);
}

template <int MF, int KS, int NPF, int NWARP, int KMODE>
__global__ void __launch_bounds__(NWARP * 32, 1) kmatvec_kernel(const KmvParams p) {
  constexpr int BR = NWARP * 8 * MF;
  constexpr int NT = NWARP * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // bar[0], bar[1]
  double* sZr = reinterpret_cast<double*>(smem_raw + 128);
  const int za_dbl = BJ * p.S, v_dbl = BJ * p.PS + 8;
  double* stage0 = sZr + BR * p.S;
  const int stage_dbl = za_dbl + v_dbl;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lq = lane & 3;
  const int row0 = blockIdx.x * BR;
  const int rows1 = min(BR, p.n_rows - row0);
  const int jbeg = blockIdx.y * p.jps, jend = min(p.N, jbeg + p.jps);
  const int ntiles = (jend - jbeg + BJ - 1) / BJ;

  __shared__ double etab[256];  // 2^(j/256) for exp2_tab256
  if (KMODE != 1 && tid < 256) etab[tid] = exp2((double)tid * (1.0 / 256.0));
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int j0 = jbeg + t * BJ;
    const int rows2 = min(BJ, jend - j0);
    double* sa = stage0 + (t & 1) * stage_dbl;
    const uint32_t b1 = (uint32_t)rows2 * p.S * 8u, b2 = (uint32_t)rows2 * p.PS * 8u;
    uint32_t tot = b1 + b2;
    if (t == 0) tot += (uint32_t)rows1 * p.S * 8u;
    mbar_arrive_expect_tx(&bar[t & 1], tot);
    if (t == 0) tma_bulk_g2s(sZr, p.Zr + (long long)row0 * p.S, (uint32_t)rows1 * p.S * 8u, &bar[0]);
    tma_bulk_g2s(sa, p.Za + (long long)j0 * p.S, b1, &bar[t & 1]);
    tma_bulk_g2s(sa + za_dbl, p.V + (long long)j0 * p.PS, b2, &bar[t & 1]);
  };
  if (tid == 0 && ntiles > 0) issue(0);

  // NPF == 0: single-vector mode (PCG).  The 8-column DMMA GEMM 2 would spend 8 tensor instructions per 8 x 8 tile
  // of K on one useful column; instead every lane multiplies its two kernel entries by the two vector entries with
  // plain FMAs and the four lanes of a row are combined once, after the column loop.
  constexpr int NPA = NPF > 0 ? NPF : 1;
  double w[MF][NPA][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NPA; ++j) w[i][j][0] = w[i][j][1] = 0.0;

  double a1[MF][KS > 0 ? KS : 1];
  double ni[MF];
  const double* zr_base = sZr + (warp * 8 * MF + lr) * p.S;

  for (int t = 0; t < ntiles; ++t) {
    if (tid == 0 && t + 1 < ntiles) issue(t + 1);
    mbar_wait(&bar[t & 1], (t >> 1) & 1);
    const int j0 = jbeg + t * BJ;
    const int rows2 = min(BJ, jend - j0);
    double* sZa = stage0 + (t & 1) * stage_dbl;
    double* sV = sZa + za_dbl;
    if (t == 0) {
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) {
        ni[mf] = zr_base[mf * 8 * p.S + p.Dp];
        if (KMODE == 2) ni[mf] *= -EXP2_256_SCALE;  // the fast path works on u = -d2 * 256 log2(e)
        if (KS > 0) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) a1[mf][ks] = zr_base[mf * 8 * p.S + ks * 4 + lq];
        }
      }
    }
    if (rows2 < BJ) {  // last tile: rows beyond the data must contribute exactly zero
      for (int idx = rows2 * p.PS + tid; idx < v_dbl; idx += NT) sV[idx] = 0.0;
      __syncthreads();
    }
#pragma unroll 2
    for (int jf = 0; jf < BJ / 8; ++jf) {
      double c[MF][2];
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) c[mf][0] = c[mf][1] = 0.0;
      const double* zb = sZa + (jf * 8 + lr) * p.S + lq;
      if (KS > 0) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double b = zb[ks * 4];
#pragma unroll
          for (int mf = 0; mf < MF; ++mf) dmma884(c[mf][0], c[mf][1], a1[mf][ks], b);
        }
      } else {
        for (int k = 0; k < p.Dp; k += 4) {
          const double b = zb[k];
#pragma unroll
          for (int mf = 0; mf < MF; ++mf) dmma884(c[mf][0], c[mf][1], zr_base[mf * 8 * p.S + k + lq], b);
        }
      }
      const int jl = jf * 8 + lq * 2;
      // |z_j|^2 + 1e-12 once per column (amortised over the MF row fragments); the shifted diagonal
      // (diag_add * V) is added after the column loop, so the inner loop carries no extra FP64 add
      const double nj0 = sZa[jl * p.S + p.Dp] + 1e-12, nj1 = sZa[(jl + 1) * p.S + p.Dp] + 1e-12;
      const int gj = j0 + jl;
      if (KMODE == 2) {
        // SquaredExponential fast path: u = -d2 * 256 log2(e) comes straight out of one FMA and one add per
        // element (the scale is folded into the row / column norms), then exp(-d2) = 2^(u/256)
        const double njc0 = EXP2_256_SCALE * nj0, njc1 = EXP2_256_SCALE * nj1;
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) {
          const long long gi = p.row_offset + row0 + warp * 8 * MF + mf * 8 + lr;
          double u0 = fma(c[mf][0], 2.0 * EXP2_256_SCALE, ni[mf]) - njc0;
          double u1 = fma(c[mf][1], 2.0 * EXP2_256_SCALE, ni[mf]) - njc1;
          if ((p.row_offset >= 0) && (gi == gj)) u0 = -EXP2_256_SCALE * 1e-12;
          if ((p.row_offset >= 0) && (gi == gj + 1)) u1 = -EXP2_256_SCALE * 1e-12;
          const double k0 = exp2_tab256(u0, etab), k1 = exp2_tab256(u1, etab);
          c[mf][0] = (gj < jend) ? k0 : 0.0;
          c[mf][1] = (gj + 1 < jend) ? k1 : 0.0;
        }
      } else
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) {
        const long long gi = p.row_offset + row0 + warp * 8 * MF + mf * 8 + lr;
        double d20 = fma(-2.0, c[mf][0], ni[mf]) + nj0;
        double d21 = fma(-2.0, c[mf][1], ni[mf]) + nj1;
        const bool dg0 = (p.row_offset >= 0) && (gi == gj), dg1 = (p.row_offset >= 0) && (gi == gj + 1);
        if (dg0) d20 = 1e-12;
        if (dg1) d21 = 1e-12;
        double k0, k1;
        if (KMODE == 4) {
          k0 = kfun_dl(p.kind, d20, dg0 ? 0.0 : d20 - 1e-12, p.rl);
          k1 = kfun_dl(p.kind, d21, dg1 ? 0.0 : d21 - 1e-12, p.rl);
        } else {
          k0 = kfun_mode<KMODE>(p.kind, d20, etab);
          k1 = kfun_mode<KMODE>(p.kind, d21, etab);
        }
        if (KMODE == 3) {
          const int il = row0 + warp * 8 * MF + mf * 8 + lr;  // local row (wrow / wt are indexed locally)
          const bool rok = il < p.n_rows;
          const double wr = rok ? p.wrow[il] : 0.0;
          double t0 = 0.0, t1 = 0.0;
          if (rok && gj < jend) t0 = p.wt_trans ? p.wt[(long long)gj * p.ldwt + il] : p.wt[(long long)il * p.ldwt + gj];
          if (rok && gj + 1 < jend)
            t1 = p.wt_trans ? p.wt[(long long)(gj + 1) * p.ldwt + il] : p.wt[(long long)il * p.ldwt + gj + 1];
          k0 *= (gj < jend) ? wr * p.wcol[gj] - t0 : 0.0;
          k1 *= (gj + 1 < jend) ? wr * p.wcol[gj + 1] - t1 : 0.0;
        }
        c[mf][0] = (gj < jend) ? k0 : 0.0;
        c[mf][1] = (gj + 1 < jend) ? k1 : 0.0;
      }
      if (NPF == 0) {
        const double v0 = sV[(jf * 8 + 2 * lq) * p.PS], v1 = sV[(jf * 8 + 2 * lq + 1) * p.PS];
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) {
          w[mf][0][0] = fma(c[mf][0], v0, w[mf][0][0]);
          w[mf][0][1] = fma(c[mf][1], v1, w[mf][0][1]);
        }
      } else {
        const double* vb = sV + (jf * 8 + 2 * lq) * p.PS + lr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int pf = 0; pf < NPA; ++pf) {
            const double v = vb[h * p.PS + pf * 8];
#pragma unroll
            for (int mf = 0; mf < MF; ++mf) dmma884(w[mf][pf][0], w[mf][pf][1], c[mf][h], v);
          }
        }
      }
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }

  // ---- write the row block (+ diag_add * V[i] from the CTA whose column range holds the diagonal) ----
  if (NPF == 0) {
#pragma unroll
    for (int mf = 0; mf < MF; ++mf) {
      double t = w[mf][0][0] + w[mf][0][1];  // fixed order: lane pair sums, then the four lanes of the row
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      const int i = row0 + warp * 8 * MF + mf * 8 + lr;
      if (lq != 0 || i >= p.n_rows) continue;
      const long long gi = p.row_offset + i;
      if ((p.row_offset >= 0) && (gi >= jbeg) && (gi < jend) && (p.diag_add != 0.0)) t += p.diag_add * p.V[gi * p.PS];
      if (p.jsplit == 1) p.W[(long long)i * p.ldw] = t;
      else p.W[(long long)blockIdx.y * p.n_rows + i] = t;
    }
    return;
  }
  const int pw = 8 * NPA;
#pragma unroll
  for (int mf = 0; mf < MF; ++mf) {
    const int i = row0 + warp * 8 * MF + mf * 8 + lr;
    if (i >= p.n_rows) continue;
    const long long gi = p.row_offset + i;
    const bool own_diag = (p.row_offset >= 0) && (gi >= jbeg) && (gi < jend) && (p.diag_add != 0.0);
#pragma unroll
    for (int pf = 0; pf < NPA; ++pf) {
      const int pc = pf * 8 + 2 * lq;
      double w0 = w[mf][pf][0], w1 = w[mf][pf][1];
      if (own_diag) {
        if (pc < p.P) w0 += p.diag_add * p.V[gi * p.PS + pc];
        if (pc + 1 < p.P) w1 += p.diag_add * p.V[gi * p.PS + pc + 1];
      }
      if (p.jsplit == 1) {
        double* o = p.W + (long long)i * p.ldw + pc;
        if (pc < p.P) o[0] = w0;
        if (pc + 1 < p.P) o[1] = w1;
      } else {
        double* o = p.W + ((long long)blockIdx.y * p.n_rows + i) * pw + pc;
        o[0] = w0;
        o[1] = w1;
      }
    }
  }
}

// fixed-order sum of the column-split partials
__global__ void kmv_reduce_kernel(const double* __restrict__ part, int jsplit, int n_rows, int pw,
                                  int P, double* __restrict__ W, long long ldw) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_rows * pw) return;
  const int i = (int)(idx / pw), pc = (int)(idx % pw);
  if (pc >= P) return;
  double s = 0.0;
  for (int js = 0; js < jsplit; ++js) s += part[((long long)js * n_rows + i) * pw + pc];
  W[(long long)i * ldw + pc] = s;
}

// Vp[i][0..PS) = V[i][c0..c0+P) zero padded
__global__ void pack_v_kernel(const double* __restrict__ V, long long ldv, int n, int c0, int P, int PS,
                              double* __restrict__ Vp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * PS) return;
  const int i = (int)(idx / PS), c = (int)(idx % PS);
  Vp[idx] = c < P ? V[(long long)i * ldv + c0 + c] : 0.0;
}

template <int MF, int KS, int NPF, int NWARP, int KMODE>
int launch_k(Ctx* ctx, const KmvParams& p, dim3 grid, int smem, cudaStream_t s) {
  GPXB_CUDA_CHECK(ensure_dyn_smem((const void*)(kmatvec_kernel<MF, KS, NPF, NWARP, KMODE>), smem));
  kmatvec_kernel<MF, KS, NPF, NWARP, KMODE><<<grid, NWARP * 32, smem, s>>>(p);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

template <int MF, int KS, int NPF, int NWARP>
int launch_t(Ctx* ctx, const KmvParams& p, dim3 grid, int smem, cudaStream_t s) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("GPXB_KMV_EXP");  // "fast" (default: table-driven exp_tab, +10%) or "libm"
    mode = (e && e[0] == 'l') ? 1 : 2;
  }
  if (p.wt) return launch_k<MF, KS, NPF, NWARP, 3>(ctx, p, grid, smem, s);
  if (p.rl != 0.0) return launch_k<MF, KS, NPF, NWARP, 4>(ctx, p, grid, smem, s);
  if (p.kind == KIND_SE) {
    if (mode == 1) return launch_k<MF, KS, NPF, NWARP, 1>(ctx, p, grid, smem, s);
    return launch_k<MF, KS, NPF, NWARP, 2>(ctx, p, grid, smem, s);
  }
  return launch_k<MF, KS, NPF, NWARP, 0>(ctx, p, grid, smem, s);
}

template <int MF, int KS, int NPF, int NWARP>
int launch_unused(Ctx* ctx, const KmvParams& p, dim3 grid, int smem, cudaStream_t s) {
  kmatvec_kernel<MF, KS, NPF, NWARP, 0><<<grid, NWARP * 32, smem, s>>>(p);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

inline int smem_bytes(int BR, int S, int PS) {
  return 128 + (BR * S + 2 * (BJ * S + BJ * PS + 8)) * (int)sizeof(double);
}

}  // namespace

// largest feature count the matrix-free kernel's shared-memory tiles hold (S <= 68: D <= 64)
bool kmatvec_supports(int D) { return smem_bytes(128, z_layout(D).S, v_padded_stride(8)) <= 232448; }

int v_padded_stride(int P) {
  int ps = P < 2 ? 2 : P;
  while (ps % 8 != 2) ++ps;
  return ps;
}

namespace {
// More features than the fused kernel's shared-memory tiles hold (D > 64): the pair distances are a real GEMM then
// (2 D >= 130 flops per entry against 2 P for the product with V), so the operator is applied in row chunks --
// K-tiled Gram kernel (gram.cu) into a chunk buffer of at most ~1 GB, then one DMMA GEMM with the block V.  K is
// still never held as a whole.
int kmatvec_chunked(Ctx* ctx, const KmvArgs& a, cudaStream_t s) {
  const long long ldk = round_up((long long)a.N, 2);
  long long rc_max = std::max<long long>(128, (1LL << 27) / std::max<long long>(ldk, 1) / 128 * 128);
  rc_max = std::min<long long>(rc_max, round_up((long long)a.n_rows, 128));
  DevBuf kc;
  GPXB_TRY(kc.alloc(sizeof(double) * (size_t)rc_max * ldk, s));
  for (long long r0 = 0; r0 < a.n_rows; r0 += rc_max) {
    const int rc = (int)std::min<long long>(rc_max, a.n_rows - r0);
    GramArgs g;
    g.kind = a.kind;
    g.Z1 = a.Zr + r0 * a.zl.S; g.n1 = rc; g.Z2 = a.Za; g.n2 = a.N; g.zl = a.zl;
    g.out = kc.as<double>(); g.ldo = ldk;
    if (a.row_offset >= 0) { g.sym = 1; g.diag_off = a.row_offset + r0; g.diag_add = a.diag_add; }
    GPXB_TRY(gram_launch(ctx, g, s));
    GemmArgs m;
    m.M = rc; m.N = a.P; m.K = a.N;
    m.A = kc.as<double>(); m.lda = ldk; m.a_kmajor = true;
    m.B = a.V; m.ldb = a.PS; m.b_kmajor = false;
    m.C = a.W + r0 * a.ldw; m.ldc = a.ldw;
    GPXB_TRY(gemm_f64(ctx, m, s));
  }
  return 0;
}
}  // namespace

int kmatvec_launch(Ctx* ctx, const KmvArgs& a, cudaStream_t s) {
  if (a.n_rows <= 0) return 0;
  PhaseScope ph(ctx, PH_MATVEC, s);
  GPXB_ARG_CHECK(a.P >= 1 && a.P <= 32, -20);
  GPXB_ARG_CHECK(a.PS % 8 == 2 && a.PS >= 2 && a.PS <= 34, -21);
  GPXB_ARG_CHECK(a.PS >= a.P || a.P > 0, -22);
  KmvParams p;
  p.kind = a.kind;
  p.Zr = a.Zr; p.n_rows = a.n_rows; p.row_offset = a.row_offset;
  p.Za = a.Za; p.N = a.N; p.S = a.zl.S; p.Dp = a.zl.Dp;
  p.V = a.V; p.PS = a.PS; p.P = a.P;
  p.diag_add = a.diag_add;
  p.wrow = a.wrow; p.wcol = a.wcol; p.wt = a.wt; p.ldwt = a.ldwt; p.wt_trans = a.wt_trans;
  GPXB_ARG_CHECK(!a.wt || (a.wrow && a.wcol), -24);
  p.rl = a.dl_ell != 0.0 ? 1.0 / a.dl_ell : 0.0;
  GPXB_ARG_CHECK(!(a.wt && a.dl_ell != 0.0), -25);
  static int gemv_pref = -1;
  if (gemv_pref < 0) {
    const char* e = getenv("GPXB_KMV_GEMV");  // "0": single vectors go through the 8-column DMMA path as before
    gemv_pref = (e && e[0] == '0') ? 0 : 1;
  }
  const int NPF = (a.P == 1 && gemv_pref) ? 0 : (a.P <= 8 ? 1 : 4);
  const int limit = 232448;
  // variants: (MF=4, 8 warps, 256 rows) default; (MF=2, 16 warps, 256 rows) via GPXB_KMV_WARPS=16;
  // (MF=2, 8 warps, 128 rows) when the 256-row tile does not fit in shared memory
  static int warps_pref = -1;
  if (warps_pref < 0) {
    const char* e = getenv("GPXB_KMV_WARPS");
    warps_pref = (e && atoi(e) == 16) ? 16 : 8;  // 8 warps x 32 rows measured fastest with exp_tab
  }
  int BR = 256, MF = warps_pref == 16 ? 2 : 4, NW = warps_pref;
  if (smem_bytes(256, p.S, p.PS) > limit) { BR = 128; MF = 2; NW = 8; }
  if (smem_bytes(BR, p.S, p.PS) > limit) {
    GPXB_ARG_CHECK(!a.wt && a.dl_ell == 0.0, -23);  // the derivative profiles keep the shared-memory feature limit
    return kmatvec_chunked(ctx, a, s);
  }
  const int row_ctas = ceil_div(a.n_rows, BR);
  // column split so that small row counts still fill the GPU (deterministic two-pass reduce)
  int jsplit = 1;
  const int jt = ceil_div(a.N, BJ);
  const int target = 2 * ctx->num_sms;
  if (row_ctas < target) {
    jsplit = std::min(jt, std::max(1, target / row_ctas));
  } else {
    // few, fractional waves (e.g. 6.6 on a row shard): pick the column split with the smallest tail
    double best = 0.0;
    for (int k = 1; k <= 8 && k <= jt; ++k) {
      const double w = (double)row_ctas * k / ctx->num_sms;
      const double eff = w / std::ceil(w);
      if (eff > best + 0.015) { best = eff; jsplit = k; }
    }
  }
  int jps = ceil_div(jt, jsplit) * BJ;
  jsplit = ceil_div(a.N, jps);
  p.jsplit = jsplit; p.jps = jps;
  DevBuf part;
  const int pw = NPF == 0 ? 1 : 8 * NPF;
  if (jsplit > 1) {
    GPXB_TRY(part.alloc(sizeof(double) * (size_t)jsplit * a.n_rows * pw, s));
    p.W = part.as<double>();
    p.ldw = pw;
  } else {
    p.W = a.W;
    p.ldw = a.ldw;
  }
  dim3 grid(row_ctas, jsplit);
  const int smem = smem_bytes(BR, p.S, p.PS);
  const int KS = (BR == 256 && (p.Dp == 8 || p.Dp == 16 || p.Dp == 32)) ? p.Dp / 4 : 0;
  int rc;
#define GO(mf, ks, npf, nw) rc = launch_t<mf, ks, npf, nw>(ctx, p, grid, smem, s)
#define GO2(mf, ks, nw) \
  do { if (NPF == 0) GO(mf, ks, 0, nw); else if (NPF == 1) GO(mf, ks, 1, nw); else GO(mf, ks, 4, nw); } while (0)
  if (BR == 128) GO2(2, 0, 8);
  else if (NW == 16) {
    if (KS == 2) GO2(2, 2, 16); else if (KS == 4) GO2(2, 4, 16); else if (KS == 8) GO2(2, 8, 16); else GO2(2, 0, 16);
  } else {
    if (KS == 2) GO2(4, 2, 8); else if (KS == 4) GO2(4, 4, 8); else if (KS == 8) GO2(4, 8, 8); else GO2(4, 0, 8);
  }
#undef GO2
#undef GO
  GPXB_TRY(rc);
  if (jsplit > 1) {
    const long long tot = (long long)a.n_rows * pw;
    kmv_reduce_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(part.as<double>(), jsplit, a.n_rows, pw, a.P,
                                                          a.W, a.ldw);
    GPXB_LAUNCH_CHECK();
    ctx->launches++;
  }
  return 0;
}

int pack_v(Ctx* ctx, const double* V, long long ldv, int n, int c0, int P, int PS, double* Vp,
           cudaStream_t s) {
  const long long tot = (long long)n * PS;
  pack_v_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(V, ldv, n, c0, P, PS, Vp);
  GPXB_LAUNCH_CHECK();
  ctx->launches++;
  return 0;
}

// General entry: any ldv / P; columns are processed in blocks of 32.
int kmatvec(Ctx* ctx, int kind, const double* x_rows, int n_rows, long long row_offset,
            const double* x_all, int N, int D, double ell, double diag_add, const double* V,
            long long ldv, int P, double* W, long long ldw, cudaStream_t s) {
  GPXB_ARG_CHECK(n_rows >= 0 && N > 0 && D > 0 && P > 0, -1);
  if (n_rows == 0) return 0;
  const ZLayout zl = z_layout(D);
  DevBuf za, zr, vp;
  GPXB_TRY(za.alloc(sizeof(double) * ((size_t)N * zl.S + 16), s));
  GPXB_TRY(prep_z(ctx, x_all, D, N, zl, ell, za.as<double>(), s));
  const double* Zr = za.as<double>();
  if (!(x_rows == x_all && n_rows == N)) {
    GPXB_TRY(zr.alloc(sizeof(double) * ((size_t)n_rows * zl.S + 16), s));
    GPXB_TRY(prep_z(ctx, x_rows, D, n_rows, zl, ell, zr.as<double>(), s));
    Zr = zr.as<double>();
  }
  for (int c0 = 0; c0 < P; c0 += 32) {
    const int pb = std::min(32, P - c0);
    const int PS = v_padded_stride(pb);
    GPXB_TRY(vp.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
    GPXB_TRY(pack_v(ctx, V, ldv, N, c0, pb, PS, vp.as<double>(), s));
    KmvArgs a;
    a.kind = kind; a.Zr = Zr; a.n_rows = n_rows; a.row_offset = row_offset;
    a.Za = za.as<double>(); a.N = N; a.zl = zl;
    a.V = vp.as<double>(); a.PS = PS; a.P = pb;
    a.W = W + c0; a.ldw = ldw; a.diag_add = diag_add;
    GPXB_TRY(kmatvec_launch(ctx, a, s));
    vp.release();
  }
  return 0;
}

// W[n_rows x P] (+)= (g(Zr, Za) o (wrow wcol^T - wt)) V: the weighted derivative-profile product behind the
// gradient w.r.t. landmark coordinates.  Zr / Za are pre-scaled operands (z_layout), V is N x ldv.
int kprofile_weighted(Ctx* ctx, int kind, const double* Zr, int n_rows, long long row_offset, const double* Za,
                      int N, ZLayout zl, const double* V, long long ldv, int P, const double* wrow,
                      const double* wcol, const double* wt, long long ldwt, int wt_trans, double* W,
                      long long ldw, cudaStream_t s) {
  GPXB_ARG_CHECK(n_rows > 0 && N > 0 && P > 0 && wrow && wcol && wt, -1);
  DevBuf vp;
  for (int c0 = 0; c0 < P; c0 += 32) {
    const int pb = std::min(32, P - c0);
    const int PS = v_padded_stride(pb);
    GPXB_TRY(vp.alloc(sizeof(double) * ((size_t)N * PS + 16), s));
    GPXB_TRY(pack_v(ctx, V, ldv, N, c0, pb, PS, vp.as<double>(), s));
    KmvArgs a;
    a.kind = kind; a.Zr = Zr; a.n_rows = n_rows; a.row_offset = row_offset;
    a.Za = Za; a.N = N; a.zl = zl;
    a.V = vp.as<double>(); a.PS = PS; a.P = pb;
    a.W = W + c0; a.ldw = ldw;
    a.wrow = wrow; a.wcol = wcol; a.wt = wt; a.ldwt = ldwt; a.wt_trans = wt_trans;
    GPXB_TRY(kmatvec_launch(ctx, a, s));
    vp.release();
  }
  return 0;
}

}  // namespace gpxb
