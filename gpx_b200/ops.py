"""Torch-tensor front end of the C ABI (device buffers in, device buffers out).

Every function enqueues on torch's current CUDA stream and returns CUDA float64
tensors; nothing here computes on the CPU.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, context, ptr, stream_ptr

KINDS = {"se": 0, "m12": 1, "m32": 2, "m52": 3}
MEAN_ZERO, MEAN_DATA = 0, 1


def _kind(kind) -> int:
    return KINDS[kind] if isinstance(kind, str) else int(kind)


def _f64(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.float64, "expected a CUDA float64 tensor"
    return t.contiguous()


def gram(kind, x1, x2, ell, diag_add=0.0, lower_only=False, out=None):
    """k(x1, x2) + diag_add*I.  Passing the same tensor twice selects the symmetric path."""
    ctx = context()
    x1 = _f64(x1)
    sym = x2 is x1 or x2 is None
    x2 = x1 if sym else _f64(x2)
    n1, D = x1.shape
    n2 = x2.shape[0]
    assert x2.shape[1] == D
    if out is None:
        out = torch.empty((n1, n2), dtype=torch.float64, device=x1.device)
        if lower_only:
            out.zero_()
    check(
        ctx.lib.gpxb_gram(ctx.handle, _kind(kind), ptr(x1), n1, ptr(x2), n2, D, float(ell), float(diag_add),
                          ptr(out), out.stride(0), int(bool(lower_only)), stream_ptr()),
        "gpxb_gram",
    )
    return out


def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None, lower_only=False):
    ctx = context()
    A, B = _f64(A), _f64(B)
    M, K = (A.shape[1], A.shape[0]) if transa else A.shape
    N = B.shape[0] if transb else B.shape[1]
    assert (B.shape[1] if transb else B.shape[0]) == K
    if C is None:
        C = torch.zeros((M, N), dtype=torch.float64, device=A.device)
    check(
        ctx.lib.gpxb_gemm(ctx.handle, int(transa), int(transb), M, N, K, float(alpha), ptr(A), A.stride(0),
                          ptr(B), B.stride(0), float(beta), ptr(C), C.stride(0), int(bool(lower_only)),
                          stream_ptr()),
        "gpxb_gemm",
    )
    return C


def potrf(A):
    """In-place lower Cholesky; returns (A, inv) with inv the diagonal-block inverses."""
    ctx = context()
    A = _f64(A)
    n = A.shape[0]
    inv = torch.empty(int(ctx.lib.gpxb_potrf_inv_size(n)), dtype=torch.float64, device=A.device)
    check(ctx.lib.gpxb_potrf(ctx.handle, ptr(A), A.stride(0), n, ptr(inv), stream_ptr()), "gpxb_potrf")
    return A, inv


def trsm(L, inv, B, trans=True):
    """B <- B L^-T (trans) or B L^-1, in place.  B is (m, n)."""
    ctx = context()
    B = _f64(B)
    m, n = B.shape
    check(
        ctx.lib.gpxb_trsm(ctx.handle, ptr(L), L.stride(0), ptr(inv), n, ptr(B), B.stride(0), m, int(bool(trans)),
                          stream_ptr()),
        "gpxb_trsm",
    )
    return B


def logdet(L):
    ctx = context()
    out = torch.empty(1, dtype=torch.float64, device=L.device)
    check(ctx.lib.gpxb_logdet(ctx.handle, ptr(L), L.stride(0), L.shape[0], ptr(out), stream_ptr()), "gpxb_logdet")
    return out


def potri(L, inv, out=None):
    ctx = context()
    n = L.shape[0]
    if out is None:
        out = torch.zeros((n, n), dtype=torch.float64, device=L.device)
    check(
        ctx.lib.gpxb_potri(ctx.handle, ptr(L), L.stride(0), ptr(inv), n, ptr(out), out.stride(0), stream_ptr()),
        "gpxb_potri",
    )
    return out


def gpr_lml_grad(kind, x, y, ell, sigma, mean_mode=MEAN_DATA, want_grad=True):
    """Device tensor [lml/N, d/d ell, d/d sigma]."""
    ctx = context()
    x, y = _f64(x), _f64(y)
    N, D = x.shape
    T = y.shape[1] if y.dim() > 1 else 1
    out = torch.empty(3, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_lml_grad(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, T, float(ell), float(sigma),
                                  int(mean_mode), int(bool(want_grad)), ptr(out), stream_ptr()),
        "gpxb_gpr_lml_grad",
    )
    return out


def gpr_fit(kind, x, y, ell, sigma, mean_mode=MEAN_DATA):
    ctx = context()
    x, y = _f64(x), _f64(y)
    N, D = x.shape
    T = y.shape[1] if y.dim() > 1 else 1
    c = torch.empty((N, T), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_fit(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, T, float(ell), float(sigma),
                             int(mean_mode), ptr(c), ptr(mu), stream_ptr()),
        "gpxb_gpr_fit",
    )
    return c, mu


def gpr_predict(kind, x_train, x, c, mu, ell, sigma, full_covariance=False):
    ctx = context()
    x_train, x, c = _f64(x_train), _f64(x), _f64(c)
    N, D = x_train.shape
    ns = x.shape[0]
    T = c.shape[1] if c.dim() > 1 else 1
    mean = torch.empty((ns, T), dtype=torch.float64, device=x.device)
    cov = torch.empty((ns, ns), dtype=torch.float64, device=x.device) if full_covariance else None
    check(
        ctx.lib.gpxb_gpr_predict(ctx.handle, _kind(kind), ptr(x_train), N, D, ptr(x), ns, ptr(c), T, ptr(mu),
                                 float(ell), float(sigma), ptr(mean), ptr(cov), stream_ptr()),
        "gpxb_gpr_predict",
    )
    return (mean, cov) if full_covariance else mean


def launches() -> int:
    return _lib.context().launches()
