// Stationary kernel nonlinearities k = phi(d2) on the device, with one exponential call site and a
// branch-free FP64 exp for non-positive arguments (the hot loop of the matrix-free mat-vec).
//   SE   exp(-d2)                                   gpx/kernels/kernels.py:751-757
//   M12  exp(-r),            r = sqrt(max(d2,1e-36))                     :922-927
//   M32  (1 + d) exp(-d),    d = sqrt(3) r                               :966-971
//   M52  (1 + d + d^2/3) exp(-d), d = sqrt(5) r                          :1050-1055
#pragma once
#include "gram.cuh"

namespace gpxb {

// exp(x) for x <= ~0 (|relative error| < 2 ulp for x >= -708; exactly 0 below, where the true
// value is < 3e-308).  n = rint(x log2 e), r = x - n ln2 (two-term Cody-Waite), degree-13 Taylor
// polynomial on |r| <= 0.347 (truncation 4e-18), scaling by 2^n through the exponent field.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double kf = fma(x, 1.4426950408889634, SHIFT);
  const int ki = __double2loint(kf);
  const double kd = kf - SHIFT;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;            // 1/13!
  p = fma(p, r, 2.08767569878681e-09);          // 1/12!
  p = fma(p, r, 2.505210838544172e-08);         // 1/11!
  p = fma(p, r, 2.755731922398589e-07);         // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
  p = fma(p, r, 2.48015873015873e-05);          // 1/8!
  p = fma(p, r, 1.984126984126984e-04);         // 1/7!
  p = fma(p, r, 1.388888888888889e-03);         // 1/6!
  p = fma(p, r, 8.333333333333333e-03);         // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double s = __hiloint2double(__double2hiint(p) + (ki << 20), __double2loint(p));
  return x < -708.0 ? 0.0 : s;
}

// Table-driven exp(x) for x <= ~0: x = (64 m + j) ln2/64 + r, |r| <= ln2/128, so a degree-5 Taylor
// polynomial suffices (truncation 3.5e-17) and exp(x) = 2^m * T[j] * p(r) with T[j] = 2^(j/64) read
// from a 64-entry table in shared memory.  10 FP64 operations instead of ~22: on sm_100a DMMA and
// the FP64 FMA pipe share one datapath (tools/pipe_probe.cu), so every FP64 op removed from the
// mat-vec epilogue is tensor throughput gained.
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ T) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double kf = fma(x, 92.332482616893657, SHIFT);  // 64 / ln2
  const int ki = __double2loint(kf);
  const double kd = kf - SHIFT;
  double r = fma(kd, -0.010830424667801708, x);      // ln2/64, high part (24 trailing zero bits)
  r = fma(kd, -2.8447437476627285e-11, r);           // low part
  double p = 8.333333333333333e-03;  // 1/120
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= T[ki & 63];
  const double s = __hiloint2double(__double2hiint(p) + ((ki >> 6) << 20), __double2loint(p));
  return ki < -65372 ? 0.0 : s;  // x < -708 (integer compare: no FP64 op); the true value is < 3e-308
}

// 2^(u/256) for u <= ~0 with u already on the table's grid (u = -d2 * 256 log2 e, produced by the caller's FMA):
// k = rint(u) by the shift trick (DADD), r = u - k exactly (|r| <= 1/2), 2^(r/256) = exp(r ln2/256) by a
// degree-4 polynomial (truncation 3.8e-17), times T[k mod 256] = 2^((k mod 256)/256) and 2^(k div 256) through
// the exponent field: 3 DADD + 4 DFMA + 1 DMUL -- no range-reduction multiplies at all.
constexpr double EXP2_256_SCALE = 369.3299304675746;  // 256 log2(e)
__device__ __forceinline__ double exp2_tab256(double u, const double* __restrict__ T) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double kf = u + SHIFT;
  const int ki = __double2loint(kf);
  const double r = u - (kf - SHIFT);
  double p = 2.239395190875157e-12;            // a^4/24, a = ln2/256
  p = fma(p, r, 3.308302680541371e-09);        // a^3/6
  p = fma(p, r, 3.665565596910106e-06);        // a^2/2
  p = fma(p, r, 2.7076061740622863e-03);       // a
  p = fma(p, r, 1.0);
  p *= T[ki & 255];
  const double s = __hiloint2double(__double2hiint(p) + ((ki >> 8) << 20), __double2loint(p));
  return ki < -261632 ? 0.0 : s;  // u < -1022 * 256: below the normal range (true value < 2.3e-308)
}

// dK/d ell given d2 = s + 1e-12 (the jitter is a constant; the clamp passes no gradient).
// rl = 1/ell is hoisted by the caller: no FP64 division in the inner loop except Matern-1/2's 1/r.
__device__ __forceinline__ double kfun_dl(int kind, double d2, double s, double rl) {
  if (kind == KIND_SE) return exp(-d2) * (2.0 * rl) * s;
  if (!(d2 > 1e-36)) return 0.0;
  const double r = sqrt(d2);
  if (kind == KIND_M12) return exp(-r) * s * rl / r;
  if (kind == KIND_M32) {
    const double d = 1.7320508075688772 * r;
    return (3.0 * rl) * s * exp(-d);
  }
  const double d = 2.23606797749979 * r;
  return ((5.0 / 3.0) * rl) * s * (1.0 + d) * exp(-d);
}

__device__ __forceinline__ double kfun_fast(int kind, double d2) {
  double t = d2;
  if (kind != KIND_SE) {
    const double c = kind == KIND_M12 ? 1.0 : (kind == KIND_M32 ? 1.7320508075688772 : 2.23606797749979);
    t = c * sqrt(fmax(d2, 1e-36));
  }
  const double e = exp_nonpos(-t);
  double poly = 1.0;
  if (kind == KIND_M32) poly = 1.0 + t;
  if (kind == KIND_M52) poly = 1.0 + t + t * t / 3.0;
  return poly * e;
}

}  // namespace gpxb
