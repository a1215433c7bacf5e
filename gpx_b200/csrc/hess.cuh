// Internal interface of the Hessian-Jacobian kernel family (Family 1b, hess.cu).
#pragma once
#include "common.cuh"

namespace gpxb {

// out[(N1 nv1) x (N2 nv2)] = d01kj(x1, x2; J1, J2) (+ diag_add I for identical operands); ldo < 0 with
// lower_only: block-column packed lower storage (chol.cuh: BcpLayout)
int hess_gram(Ctx* ctx, int kind, const double* x1, const double* J1, int Ns1, int nv1, const double* x2,
              const double* J2, int Ns2, int nv2, int nf, double ell, double diag_add, double* out, long long ldo,
              int lower_only, cudaStream_t s);
// *contract_out = sum_ij (alpha_i alpha_j - Ainv_ij) d(d01kj(x, x; J, J))_ij / d lengthscale (Ainv lower-stored)
int hess_contract_dl(Ctx* ctx, int kind, const double* x, const double* J, int N, int nv, int nf, double ell,
                     const double* Ainv, long long ldainv, const double* alpha, double* contract_out, cudaStream_t s);
// d0kj(x1, x2, J1) ((N1 nv) x N2) stored at out[i * ld_i + t * ld_t]; out == nullptr: the block (dl = 0) or its
// d/d lengthscale (dl = 1) contracted with (a_i b_t - Y[i * ldy + t]) into *contract_out
int d0kj_block(Ctx* ctx, int kind, const double* x1, const double* J1, int N1, int nv, const double* x2, int N2,
               int nf, double ell, int dl, double* out, long long ld_i, long long ld_t, const double* a,
               const double* b, const double* Y, long long ldy, double* contract_out, cudaStream_t s);

}  // namespace gpxb
