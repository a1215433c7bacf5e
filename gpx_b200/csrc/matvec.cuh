// Internal interface of the matrix-free block mat-vec (Family 3a).
#pragma once
#include <algorithm>

#include "common.cuh"
#include "gram.cuh"

namespace gpxb {

// Padded row stride of a V block with P (<= 32) columns: smallest PS >= max(P, 2), PS % 8 == 2.
int v_padded_stride(int P);
bool kmatvec_supports(int D);

struct KmvArgs {
  int kind = KIND_SE;
  const double* Zr = nullptr;  // pre-scaled rows (z_layout), n_rows x S
  int n_rows = 0;
  long long row_offset = -1;  // global index of row 0 inside x_all (diagonal handling), -1: none
  const double* Za = nullptr;  // pre-scaled columns, N x S (+ 16 doubles of slack)
  int N = 0;
  ZLayout zl{};
  const double* V = nullptr;  // N x PS padded block
  int PS = 2, P = 1;
  double* W = nullptr;  // n_rows x ldw, P columns written
  long long ldw = 0;
  double diag_add = 0.0;
  // weighted derivative-profile mode (wt != nullptr): see matvec.cu KMODE 3
  const double* wrow = nullptr;
  const double* wcol = nullptr;
  const double* wt = nullptr;
  long long ldwt = 0;
  int wt_trans = 0;
  double dl_ell = 0.0;  // != 0: the operator is d K / d lengthscale at this lengthscale (diag_add must be 0)
};
int kmatvec_launch(Ctx* ctx, const KmvArgs& a, cudaStream_t s);
int pack_v(Ctx* ctx, const double* V, long long ldv, int n, int c0, int P, int PS, double* Vp,
           cudaStream_t s);
int kmatvec(Ctx* ctx, int kind, const double* x_rows, int n_rows, long long row_offset,
            const double* x_all, int N, int D, double ell, double diag_add, const double* V,
            long long ldv, int P, double* W, long long ldw, cudaStream_t s);

int kprofile_weighted(Ctx* ctx, int kind, const double* Zr, int n_rows, long long row_offset, const double* Za,
                      int N, ZLayout zl, const double* V, long long ldv, int P, const double* wrow,
                      const double* wcol, const double* wt, long long ldwt, int wt_trans, double* W,
                      long long ldw, cudaStream_t s);

}  // namespace gpxb
