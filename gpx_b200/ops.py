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


def _require(cond, msg) -> None:
    """Argument validation that survives `python -O` (the C ABI cannot see tensor shapes)."""
    if not cond:
        raise ValueError(msg)


def _rows(t: torch.Tensor) -> int:
    return t.shape[0] if t.dim() > 0 else 1


def _check_xy(x, y, what="x and y"):
    """x (N, D) and y (N,) or (N, T): the C ABI sees only pointers, so a mismatch would read past a buffer."""
    _require(x.dim() == 2, f"{what}: x must be (N, D), got {tuple(x.shape)}")
    _require(y.dim() in (1, 2) and y.shape[0] == x.shape[0],
             f"{what}: y has {tuple(y.shape)} but x has {x.shape[0]} rows")


def _check_xJ(x, J, what="x and jacobian"):
    """x (N, nf), J (N, nf, nv)."""
    _require(x.dim() == 2 and J.dim() == 3, f"{what}: expected x (N, nf) and jacobian (N, nf, nv)")
    _require(J.shape[0] == x.shape[0] and J.shape[1] == x.shape[1],
             f"{what}: jacobian {tuple(J.shape)} does not match x {tuple(x.shape)}")


def _f64(t: torch.Tensor) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.float64):
        raise _lib.GpxbError(f"expected a CUDA float64 tensor, got {t.device} {t.dtype} (there is no CPU path)")
    return t.contiguous()


def gram(kind, x1, x2, ell, diag_add=0.0, lower_only=False, out=None):
    """k(x1, x2) + diag_add*I.  Passing the same tensor twice selects the symmetric path."""
    ctx = context()
    x1 = _f64(x1)
    sym = x2 is x1 or x2 is None
    x2 = x1 if sym else _f64(x2)
    n1, D = x1.shape
    n2 = x2.shape[0]
    _require(x2.shape[1] == D, "shape mismatch between the operands")
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


KIND_LINEAR = 4
OP_SUM, OP_PROD = -1, -2


def gram_composite_fits(leaf_dims) -> bool:
    """True when the fused composite Gram kernel can hold the operand tiles of all leaves in shared memory."""
    import ctypes

    ctx = context()
    arr = (ctypes.c_int * len(leaf_dims))(*[int(d) for d in leaf_dims])
    return bool(ctx.lib.gpxb_gram_composite_fits(len(leaf_dims), arr))


def gram_composite(kinds, ells, x1_leaves, x2_leaves, prog, diag_add=0.0, lower_only=False, out=None):
    """Gram matrix of a Sum / Prod tree in one tile-level pass.  kinds: per leaf "se" | "m12" | "m32" | "m52" |
    "linear"; x1_leaves / x2_leaves: per leaf the (already sliced) device inputs (x2_leaves None: symmetric);
    prog: postfix program (leaf indices, OP_SUM, OP_PROD)."""
    import ctypes

    ctx = context()
    L = len(kinds)
    sym = x2_leaves is None
    x1_leaves = [_f64(t) for t in x1_leaves]
    x2_leaves = x1_leaves if sym else [_f64(t) for t in x2_leaves]
    n1, n2 = x1_leaves[0].shape[0], x2_leaves[0].shape[0]
    for a, b in zip(x1_leaves, x2_leaves):
        _require(a.dim() == 2 and b.dim() == 2 and a.shape[0] == n1 and b.shape[0] == n2 and a.shape[1] == b.shape[1],
                 "gram_composite: leaf operands disagree in shape")
    _require(len(ells) == L and len(x1_leaves) == L, "gram_composite: one lengthscale and one operand per leaf")
    kid = (ctypes.c_int * L)(*[KIND_LINEAR if k == "linear" else _kind(k) for k in kinds])
    el = (ctypes.c_double * L)(*[float(e) for e in ells])
    dd = (ctypes.c_int * L)(*[int(t.shape[1]) for t in x1_leaves])
    p1 = (ctypes.c_void_p * L)(*[t.data_ptr() for t in x1_leaves])
    p2 = (ctypes.c_void_p * L)(*[t.data_ptr() for t in x2_leaves])
    pr = (ctypes.c_int * len(prog))(*[int(v) for v in prog])
    if out is None:
        out = torch.empty((n1, n2), dtype=torch.float64, device=x1_leaves[0].device)
        if lower_only:
            out.zero_()
    check(
        ctx.lib.gpxb_gram_composite(ctx.handle, L, kid, el, dd, p1, p2, n1, n2, pr, len(prog), float(diag_add),
                                    int(sym), int(bool(lower_only)), ptr(out), out.stride(0), stream_ptr()),
        "gpxb_gram_composite",
    )
    return out


def _comp_desc(kinds, ells, x_leaves, prog):
    """ctypes arrays describing a composite kernel over per-leaf inputs (kept alive by the caller)."""
    import ctypes

    L = len(kinds)
    _require(len(ells) == L and len(x_leaves) == L, "composite kernel: one lengthscale and one operand per leaf")
    kid = (ctypes.c_int * L)(*[KIND_LINEAR if k == "linear" else _kind(k) for k in kinds])
    el = (ctypes.c_double * L)(*[float(e) for e in ells])
    dd = (ctypes.c_int * L)(*[int(t.shape[1]) for t in x_leaves])
    px = (ctypes.c_void_p * L)(*[t.data_ptr() for t in x_leaves])
    pr = (ctypes.c_int * len(prog))(*[int(v) for v in prog])
    return L, kid, el, dd, px, pr


def rpcholesky_composite(kinds, ells, x_leaves, prog, n_pivots, key, partitionable=True, want_factor=True):
    """(F (N, M) or None, pivots int64 (M,)) -- RPCholesky (gpx/kernels/approximations.py:30-129) of a Sum / Prod kernel."""
    ctx = context()
    x_leaves = [_f64(t) for t in x_leaves]
    N = x_leaves[0].shape[0]
    L, kid, el, dd, px, pr = _comp_desc(kinds, ells, x_leaves, prog)
    piv = torch.empty(n_pivots, dtype=torch.int64, device=x_leaves[0].device)
    F = torch.empty((N, n_pivots), dtype=torch.float64, device=piv.device) if want_factor else None
    check(
        ctx.lib.gpxb_rpcholesky_composite(ctx.handle, L, kid, el, dd, px, N, pr, len(prog), int(key[0]), int(key[1]),
                                          int(n_pivots), int(bool(partitionable)), ptr(F) if want_factor else None,
                                          ptr(piv), stream_ptr()),
        "gpxb_rpcholesky_composite",
    )
    return F, piv


def gpr_iter_composite(kinds, ells, x_leaves, prog, sigma, y=None, n_pivots=0, key=(0, 0), mean_mode=MEAN_DATA,
                       partitionable=True, tol=1e-5, atol=1e-7, maxiter=0, U=None, m=0, restart_key=None,
                       restart_partitionable=True):
    """The iterative GPR paths with a Sum / Prod kernel.  y given: PCG fit -> c (N, 1), mu (1,), iterations;
    U given: m-step block Lanczos on K + (sigma^2+1e-10) I -> alpha, beta (P, m), breakdown flag.
    Returns a dict with the parts that were asked for."""
    import ctypes

    ctx = context()
    x_leaves = [_f64(t) for t in x_leaves]
    N = x_leaves[0].shape[0]
    dev = x_leaves[0].device
    L, kid, el, dd, px, pr = _comp_desc(kinds, ells, x_leaves, prog)
    out = {}
    c = mu = alpha = beta = flag = None
    iters = ctypes.c_int(0)
    if y is not None:
        y = _f64(y)
        _require(y.numel() == N, "the iterative path supports a single output column")
        c = torch.empty((N, 1), dtype=torch.float64, device=dev)
        mu = torch.empty(1, dtype=torch.float64, device=dev)
    P = 0
    if U is not None:
        U = _f64(U)
        _require(U.dim() == 2 and U.shape[0] == N, f"U has {tuple(U.shape)} but x has {N} rows")
        P = U.shape[1]
        alpha = torch.empty((P, m), dtype=torch.float64, device=dev)
        beta = torch.empty((P, m), dtype=torch.float64, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
    with _lanczos_restart(restart_key, restart_partitionable):
        check(
            ctx.lib.gpxb_gpr_iter_composite(ctx.handle, L, kid, el, dd, px, N, pr, len(prog),
                                            ptr(y) if y is not None else None, float(sigma), int(mean_mode), int(n_pivots),
                                            int(key[0]), int(key[1]), int(bool(partitionable)), float(tol), float(atol),
                                            int(maxiter), ptr(c) if c is not None else None,
                                            ptr(mu) if mu is not None else None, ctypes.byref(iters),
                                            ptr(U) if U is not None else None, U.stride(0) if U is not None else 0, P,
                                            int(m), ptr(alpha) if alpha is not None else None,
                                            ptr(beta) if beta is not None else None,
                                            ptr(flag) if flag is not None else None, stream_ptr()),
            "gpxb_gpr_iter_composite",
        )
    if y is not None:
        out.update(c=c, mu=mu, iters=iters.value)
    if U is not None:
        out.update(alpha=alpha, beta=beta, flag=flag)
    return out


def gram_dl_contract(kind, x1, x2, ell, Y, sym_lower=False):
    """sum_ij Y_ij d k(x1_i, x2_j)/d lengthscale (device scalar tensor), dK never formed."""
    ctx = context()
    x1, x2, Y = _f64(x1), _f64(x2), _f64(Y)
    n1, D = x1.shape
    _require(x2.dim() == 2 and x2.shape[1] == D, "shape mismatch between the operands")
    _require(Y.dim() == 2 and Y.shape == (n1, x2.shape[0]), f"Y must be ({n1}, {x2.shape[0]}), got {tuple(Y.shape)}")
    out = torch.empty(1, dtype=torch.float64, device=x1.device)
    check(
        ctx.lib.gpxb_gram_dl_contract(ctx.handle, _kind(kind), ptr(x1), n1, ptr(x2), x2.shape[0], D, float(ell), ptr(Y),
                                      Y.stride(0), int(bool(sym_lower)), ptr(out), stream_ptr()),
        "gpxb_gram_dl_contract",
    )
    return out


def elementwise(op, X, Y, a=1.0, b=1.0, out=None):
    """out = a X + b Y (op "axpby") or a X o Y (op "mul") on contiguous device tensors; out may alias X or Y."""
    ctx = context()
    _require(X.is_contiguous() and Y.is_contiguous() and X.shape == Y.shape, "shape mismatch between the operands")
    if out is None:
        out = torch.empty_like(X)
    check(
        ctx.lib.gpxb_elementwise(ctx.handle, 0 if op == "axpby" else 1, X.numel(), float(a), ptr(X), float(b), ptr(Y),
                                 ptr(out), stream_ptr()),
        "gpxb_elementwise",
    )
    return out


def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None, lower_only=False):
    ctx = context()
    A, B = _f64(A), _f64(B)
    M, K = (A.shape[1], A.shape[0]) if transa else A.shape
    N = B.shape[0] if transb else B.shape[1]
    _require((B.shape[1] if transb else B.shape[0]) == K, "shape mismatch between the operands")
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
    _require(L.dim() == 2 and L.shape[0] >= n and L.shape[1] >= n, "trsm: B has more columns than L has rows")
    _require(inv.numel() >= int(ctx.lib.gpxb_potrf_inv_size(n)), "trsm: the leaf-inverse buffer is too small")
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
    _check_xy(x, y, "gpr_lml_grad")
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
    _check_xy(x, y, "gpr_fit")
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
    _require(x.dim() == 2 and x.shape[1] == D, f"gpr_predict: x has {tuple(x.shape)} but the model has {D} features")
    _require(c.shape[0] == N, f"gpr_predict: c has {c.shape[0]} rows but x_train has {N}")
    _require(mu.numel() >= 1, "gpr_predict: mu must hold the mean")
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


def hess_gram(kind, x1, J1, x2, J2, ell, diag_add=0.0, lower_only=False):
    """Kernel.d01kj(x1, x2, params, J1, J2): ((n1*nv1), (n2*nv2)).  Same tensors twice -> symmetric path."""
    ctx = context()
    x1, J1 = _f64(x1), _f64(J1)
    sym = (x2 is x1 or x2 is None) and (J2 is J1 or J2 is None)
    x2, J2 = (x1, J1) if sym else (_f64(x2), _f64(J2))
    _check_xJ(x1, J1, "hess_gram (first operand)")
    _check_xJ(x2, J2, "hess_gram (second operand)")
    _require(x2.shape[1] == x1.shape[1], "hess_gram: the operands have different feature counts")
    n1, nf, nv1 = J1.shape
    n2, _, nv2 = J2.shape
    out = torch.empty((n1 * nv1, n2 * nv2), dtype=torch.float64, device=x1.device)
    if lower_only:
        out.zero_()
    check(
        ctx.lib.gpxb_hess_gram(ctx.handle, _kind(kind), ptr(x1), ptr(J1), n1, nv1, ptr(x2), ptr(J2), n2, nv2, nf,
                               float(ell), float(diag_add), ptr(out), out.stride(0), int(bool(lower_only)),
                               stream_ptr()),
        "gpxb_hess_gram",
    )
    return out


def gpr_lml_derivs(kind, x, J, y, ell, sigma):
    ctx = context()
    x, J, y = _f64(x), _f64(J), _f64(y)
    _check_xJ(x, J, "gpr_lml_derivs")
    N, nf, nv = J.shape
    _require(y.numel() == N * nv, f"gpr_lml_derivs: y has {y.numel()} values, expected N*nv = {N * nv}")
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_lml_derivs(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), N, nf, nv, float(ell),
                                    float(sigma), ptr(out), stream_ptr()),
        "gpxb_gpr_lml_derivs",
    )
    return out


def gpr_lml_grad_derivs(kind, x, J, y, ell, sigma, want_grad=True):
    """-> device tensor [lml, d lml/d lengthscale, d lml/d sigma] of the derivative-trained GPR."""
    ctx = context()
    x, J, y = _f64(x), _f64(J), _f64(y)
    _check_xJ(x, J, "gpr_lml_grad_derivs")
    N, nf, nv = J.shape
    _require(y.numel() == N * nv, f"gpr_lml_grad_derivs: y has {y.numel()} values, expected N*nv = {N * nv}")
    out = torch.empty(3, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_lml_grad_derivs(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), N, nf, nv, float(ell),
                                         float(sigma), 1 if want_grad else 0, ptr(out), stream_ptr()),
        "gpxb_gpr_lml_grad_derivs",
    )
    return out


def gpr_predict_y_derivs(kind, x_train, jaccoef, x, ell, mu=0.0):
    """-> (ns,) predicted target values of a derivative-trained model."""
    ctx = context()
    x_train, jaccoef, x = _f64(x_train), _f64(jaccoef), _f64(x)
    N, nf = x_train.shape
    _require(tuple(jaccoef.shape) == (N, nf), f"gpr_predict_y_derivs: jaccoef must be ({N}, {nf}), got {tuple(jaccoef.shape)}")
    _require(x.dim() == 2 and x.shape[1] == nf, f"gpr_predict_y_derivs: x has {tuple(x.shape)}, expected {nf} features")
    ns = x.shape[0]
    out = torch.empty(ns, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_predict_y_derivs(ctx.handle, _kind(kind), ptr(x_train), ptr(jaccoef), N, nf, ptr(x), ns,
                                          float(ell), float(mu), ptr(out), stream_ptr()),
        "gpxb_gpr_predict_y_derivs",
    )
    return out


def gpr_predict_derivs_full(kind, x_train, J_train, x, J, c, ell, sigma, mu=0.0, full_covariance=False):
    """Derivative values from the uncontracted coefficients: mean (ns*nvs,) [, covariance (ns*nvs, ns*nvs)]."""
    ctx = context()
    x_train, J_train, x, J, c = _f64(x_train), _f64(J_train), _f64(x), _f64(J), _f64(c)
    _check_xJ(x_train, J_train, "gpr_predict_derivs_full (training set)")
    _check_xJ(x, J, "gpr_predict_derivs_full (prediction set)")
    N, nf, nv = J_train.shape
    _require(x.shape[1] == nf, "gpr_predict_derivs_full: feature counts differ")
    _require(c.numel() == N * nv, f"gpr_predict_derivs_full: c has {c.numel()} values, expected N*nv = {N * nv}")
    ns, _, nvs = J.shape
    mean = torch.empty(ns * nvs, dtype=torch.float64, device=x.device)
    cov = torch.empty((ns * nvs, ns * nvs), dtype=torch.float64, device=x.device) if full_covariance else None
    check(
        ctx.lib.gpxb_gpr_predict_derivs_full(ctx.handle, _kind(kind), ptr(x_train), ptr(J_train), N, nf, nv, ptr(x),
                                             ptr(J), ns, nvs, ptr(c), float(mu), float(ell), float(sigma), ptr(mean),
                                             ptr(cov) if cov is not None else None, stream_ptr()),
        "gpxb_gpr_predict_derivs_full",
    )
    return (mean, cov) if full_covariance else mean


def gpr_predict_y_derivs_full(kind, x_train, J_train, x, c, ell, sigma, mu=0.0, full_covariance=False):
    """Target values from the uncontracted coefficients: mean (ns,) [, covariance (ns, ns)]."""
    ctx = context()
    x_train, J_train, x, c = _f64(x_train), _f64(J_train), _f64(x), _f64(c)
    _check_xJ(x_train, J_train, "gpr_predict_y_derivs_full")
    N, nf, nv = J_train.shape
    _require(x.dim() == 2 and x.shape[1] == nf, "gpr_predict_y_derivs_full: feature counts differ")
    _require(c.numel() == N * nv, f"gpr_predict_y_derivs_full: c has {c.numel()} values, expected N*nv = {N * nv}")
    ns = x.shape[0]
    mean = torch.empty(ns, dtype=torch.float64, device=x.device)
    cov = torch.empty((ns, ns), dtype=torch.float64, device=x.device) if full_covariance else None
    check(
        ctx.lib.gpxb_gpr_predict_y_derivs_full(ctx.handle, _kind(kind), ptr(x_train), ptr(J_train), N, nf, nv, ptr(x),
                                               ns, ptr(c), float(mu), float(ell), float(sigma), ptr(mean),
                                               ptr(cov) if cov is not None else None, stream_ptr()),
        "gpxb_gpr_predict_y_derivs_full",
    )
    return (mean, cov) if full_covariance else mean


def gpr_fit_derivs_iter(kind, x, J, y, ell, sigma, n_pivots, key, partitionable=True, tol=1e-5, atol=1e-7,
                        maxiter=0):
    """(c (N*nv,), jaccoef (N, nf), iterations): PCG on the matrix-free Hessian operator with the
    rpcholesky_derivs preconditioner (n_pivots = 0: none)."""
    import ctypes

    ctx = context()
    x, J, y = _f64(x), _f64(J), _f64(y)
    _check_xJ(x, J, "gpr_fit_derivs_iter")
    N, nf, nv = J.shape
    _require(y.numel() == N * nv, f"gpr_fit_derivs_iter: y has {y.numel()} values, expected N*nv = {N * nv}")
    c = torch.empty(N * nv, dtype=torch.float64, device=x.device)
    jc = torch.empty((N, nf), dtype=torch.float64, device=x.device)
    iters = ctypes.c_int(0)
    check(
        ctx.lib.gpxb_gpr_fit_derivs_iter(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), N, nf, nv, float(ell),
                                         float(sigma), int(n_pivots), int(key[0]), int(key[1]),
                                         int(bool(partitionable)), float(tol), float(atol), int(maxiter), ptr(c),
                                         ptr(jc), ctypes.byref(iters), stream_ptr()),
        "gpxb_gpr_fit_derivs_iter",
    )
    return c, jc, iters.value


def hess_matvec_dl(kind, x, J, V, ell):
    """(d d01kj(x, x; J, J)/d lengthscale) V, matrix-free."""
    ctx = context()
    x, J, V = _f64(x), _f64(J), _f64(V)
    one_d = V.dim() == 1
    if one_d:
        V = V.reshape(-1, 1)
    _check_xJ(x, J, "hess_matvec_dl")
    N, nf, nv = J.shape
    _require(V.shape[0] == N * nv, f"hess_matvec_dl: V has {V.shape[0]} rows, expected N*nv = {N * nv}")
    P = V.shape[1]
    W = torch.empty((N * nv, P), dtype=torch.float64, device=V.device)
    check(
        ctx.lib.gpxb_hess_matvec_dl(ctx.handle, _kind(kind), ptr(x), ptr(J), N, nv, nf, float(ell), ptr(V), V.stride(0),
                                    P, ptr(W), W.stride(0), stream_ptr()),
        "gpxb_hess_matvec_dl",
    )
    return W.reshape(-1) if one_d else W


def lanczos_grad_derivs(kind, x, J, ell, sigma, U, m):
    """lanczos_derivs plus the tangents (2, P, m) of the tridiagonals w.r.t. (lengthscale, sigma)."""
    ctx = context()
    x, J, U = _f64(x), _f64(J), _f64(U)
    _check_xJ(x, J, "lanczos_grad_derivs")
    N, nf, nv = J.shape
    _require(U.dim() == 2 and U.shape[0] == N * nv, f"lanczos_grad_derivs: U has {tuple(U.shape)}, expected {N * nv} rows")
    P = U.shape[1]
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    dalpha = torch.empty((2, P, m), dtype=torch.float64, device=x.device)
    dbeta = torch.empty((2, P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    check(
        ctx.lib.gpxb_lanczos_grad_derivs(ctx.handle, _kind(kind), ptr(x), ptr(J), N, nf, nv, float(ell), float(sigma),
                                         ptr(U), U.stride(0), P, int(m), ptr(alpha), ptr(beta), ptr(dalpha),
                                         ptr(dbeta), ptr(flag), stream_ptr()),
        "gpxb_lanczos_grad_derivs",
    )
    return alpha, beta, dalpha, dbeta, flag


def rpcholesky_derivs(kind, x, J, n_pivots, ell, key, partitionable=True, want_factor=True):
    """(F (N*nv, M) or None, pivots int64 (M,)) -- gpx/kernels/approximations.py:132-242."""
    ctx = context()
    x, J = _f64(x), _f64(J)
    _check_xJ(x, J, "rpcholesky_derivs")
    N, nf, nv = J.shape
    piv = torch.empty(n_pivots, dtype=torch.int64, device=x.device)
    F = torch.empty((N * nv, n_pivots), dtype=torch.float64, device=x.device) if want_factor else None
    check(
        ctx.lib.gpxb_rpcholesky_derivs(ctx.handle, _kind(kind), ptr(x), ptr(J), N, nf, nv, float(ell), int(key[0]),
                                       int(key[1]), int(n_pivots), int(bool(partitionable)), ptr(F), ptr(piv),
                                       stream_ptr()),
        "gpxb_rpcholesky_derivs",
    )
    return F, piv


def lanczos_derivs(kind, x, J, ell, sigma, U, m, restart_key=None, restart_partitionable=True):
    """(alpha (P, m), beta (P, m), breakdown) of the block Lanczos on d01kj + (sigma^2+1e-10) I, matrix-free."""
    ctx = context()
    x, J, U = _f64(x), _f64(J), _f64(U)
    _check_xJ(x, J, "lanczos_derivs")
    N, nf, nv = J.shape
    _require(U.dim() == 2 and U.shape[0] == N * nv, f"lanczos_derivs: U has {tuple(U.shape)}, expected {N * nv} rows")
    P = U.shape[1]
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with _lanczos_restart(restart_key, restart_partitionable):
        check(
            ctx.lib.gpxb_lanczos_derivs(ctx.handle, _kind(kind), ptr(x), ptr(J), N, nf, nv, float(ell), float(sigma), ptr(U),
                                        U.stride(0), P, int(m), ptr(alpha), ptr(beta), ptr(flag), stream_ptr()),
            "gpxb_lanczos_derivs",
        )
    return alpha, beta, flag


def hess_matvec(kind, x1, J1, x2, J2, V, ell, diag_add=0.0):
    """(d01kj(x1, x2; J1, J2) [+ diag_add I]) V without forming the matrix; V: (n2*nv2, P) or (n2*nv2,)."""
    ctx = context()
    x1, J1, x2, J2, V = _f64(x1), _f64(J1), _f64(x2), _f64(J2), _f64(V)
    one_d = V.dim() == 1
    if one_d:
        V = V.reshape(-1, 1)
    _check_xJ(x1, J1, "hess_matvec (rows)")
    _check_xJ(x2, J2, "hess_matvec (columns)")
    n1, nf, nv1 = J1.shape
    n2, _, nv2 = J2.shape
    _require(x2.shape[1] == nf, "hess_matvec: the operands have different feature counts")
    _require(V.shape[0] == n2 * nv2, f"hess_matvec: V has {V.shape[0]} rows, expected n2*nv2 = {n2 * nv2}")
    P = V.shape[1]
    W = torch.empty((n1 * nv1, P), dtype=torch.float64, device=V.device)
    check(
        ctx.lib.gpxb_hess_matvec(ctx.handle, _kind(kind), ptr(x1), ptr(J1), n1, nv1, ptr(x2), ptr(J2), n2, nv2, nf,
                                 float(ell), float(diag_add), ptr(V), V.stride(0), P, ptr(W), W.stride(0), stream_ptr()),
        "gpxb_hess_matvec",
    )
    return W.reshape(-1) if one_d else W


def td_gram(kind, x1, J1, x2, J2, ell, add_targets=0.0, add_derivs=0.0, lower_only=False):
    """Block kernel matrix of GPR_TD: [[k, d1kj], [d0kj, d01kj]] -> (n1 + n1*nv1, n2 + n2*nv2)."""
    ctx = context()
    x1, J1, x2, J2 = _f64(x1), _f64(J1), _f64(x2), _f64(J2)
    _check_xJ(x1, J1, "td_gram (first operand)")
    _check_xJ(x2, J2, "td_gram (second operand)")
    _require(x2.shape[1] == x1.shape[1], "td_gram: the operands have different feature counts")
    n1, nf, nv1 = J1.shape
    n2, _, nv2 = J2.shape
    out = torch.zeros((n1 + n1 * nv1, n2 + n2 * nv2), dtype=torch.float64, device=x1.device)
    check(
        ctx.lib.gpxb_td_gram(ctx.handle, _kind(kind), ptr(x1), ptr(J1), n1, nv1, ptr(x2), ptr(J2), n2, nv2, nf,
                             float(ell), float(add_targets), float(add_derivs), ptr(out), out.stride(0),
                             int(bool(lower_only)), stream_ptr()),
        "gpxb_td_gram",
    )
    return out


def gpr_td_lml_grad(kind, x, J, y, y_derivs, ell, sigma_targets, sigma_derivs, mean_mode=MEAN_DATA, want_grad=True):
    """-> device tensor [lml, d/d lengthscale, d/d sigma_targets, d/d sigma_derivs]."""
    ctx = context()
    x, J, y, yd = _f64(x), _f64(J), _f64(y), _f64(y_derivs)
    _check_xJ(x, J, "gpr_td")
    N, nf, nv = J.shape
    _require(y.numel() == N and yd.numel() == N * nv, "shape mismatch between the operands")
    out = torch.empty(4, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_td_lml_grad(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), ptr(yd), N, nf, nv, float(ell),
                                     float(sigma_targets), float(sigma_derivs), int(mean_mode),
                                     1 if want_grad else 0, ptr(out), stream_ptr()),
        "gpxb_gpr_td_lml_grad",
    )
    return out


def gpr_td_fit(kind, x, J, y, y_derivs, ell, sigma_targets, sigma_derivs, mean_mode=MEAN_DATA):
    """-> (c (N + N*nv, 1), mu (1,))."""
    ctx = context()
    x, J, y, yd = _f64(x), _f64(J), _f64(y), _f64(y_derivs)
    _check_xJ(x, J, "gpr_td")
    N, nf, nv = J.shape
    _require(y.numel() == N and yd.numel() == N * nv, "shape mismatch between the operands")
    c = torch.empty((N + N * nv, 1), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_td_fit(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), ptr(yd), N, nf, nv, float(ell),
                                float(sigma_targets), float(sigma_derivs), int(mean_mode), ptr(c), ptr(mu),
                                stream_ptr()),
        "gpxb_gpr_td_fit",
    )
    return c, mu


def sgpr_fit_derivs_iter(kind, x, J, y, x_locs, J_locs, ell, sigma, tol=1e-5, atol=0.0, maxiter=0):
    """(c (M nv_locs, 1), iterations): CG solve of (sigma^2 H_mm + H_mn H_nm + 1e-10 I) c = H_mn y with the
    matrix-free Hessian operators (gpx/models/_sgpr.py:399-428)."""
    import ctypes

    ctx = context()
    x, J, y, x_locs, J_locs = _f64(x), _f64(J), _f64(y), _f64(x_locs), _f64(J_locs)
    _check_xJ(x, J, "sgpr_fit_derivs_iter (data)")
    _check_xJ(x_locs, J_locs, "sgpr_fit_derivs_iter (landmarks)")
    N, nf, nv = J.shape
    M, _, nvl = J_locs.shape
    _require(x_locs.shape[1] == nf, "sgpr_fit_derivs_iter: feature counts differ")
    _require(y.numel() == N * nv, f"sgpr_fit_derivs_iter: y has {y.numel()} values, expected N*nv = {N * nv}")
    c = torch.empty((M * nvl, 1), dtype=torch.float64, device=x.device)
    iters = ctypes.c_int(0)
    check(
        ctx.lib.gpxb_sgpr_fit_derivs_iter(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), N, nf, nv, ptr(x_locs),
                                          ptr(J_locs), M, nvl, float(ell), float(sigma), float(tol), float(atol),
                                          int(maxiter), ptr(c), ctypes.byref(iters), stream_ptr()),
        "gpxb_sgpr_fit_derivs_iter",
    )
    return c, iters.value


def gpr_td_fit_iter(kind, x, J, y, y_derivs, ell, sigma_targets, sigma_derivs, n_pivots, key, mean_mode=MEAN_DATA,
                    partitionable=True, tol=1e-5, atol=1e-7, maxiter=0, nv_first=None, sigma_derivs2=0.0):
    """(c (N + N nv, 1) rows [targets; (sample, variable)], mu (1,), iterations): PCG on the matrix-free block
    operator with the rpcholesky_TD preconditioner (gpx/models/gpr_td.py:685-747).  nv_first < nv: J holds two
    derivative blocks concatenated along the variable axis, the second with noise sigma_derivs2."""
    import ctypes

    ctx = context()
    x, J, y, yd = _f64(x), _f64(J), _f64(y), _f64(y_derivs)
    _check_xJ(x, J, "gpr_td_fit_iter")
    N, nf, nv = J.shape
    _require(y.numel() == N and yd.numel() == N * nv, "shape mismatch between the operands")
    c = torch.empty((N + N * nv, 1), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    iters = ctypes.c_int(0)
    check(
        ctx.lib.gpxb_gpr_td_fit_iter(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), ptr(yd), N, nf, nv,
                                     int(nv if nv_first is None else nv_first), float(ell),
                                     float(sigma_targets), float(sigma_derivs), float(sigma_derivs2), int(mean_mode),
                                     int(n_pivots),
                                     int(key[0]), int(key[1]), int(bool(partitionable)), float(tol), float(atol),
                                     int(maxiter), ptr(c), ptr(mu), ctypes.byref(iters), stream_ptr()),
        "gpxb_gpr_td_fit_iter",
    )
    return c, mu, iters.value


def lanczos_td(kind, x, J, ell, sigma_targets, sigma_derivs, U, m, nv_first=None, sigma_derivs2=0.0, restart_key=None,
               restart_partitionable=True):
    """(alpha (P, m), beta (P, m), breakdown) of the block Lanczos on the GPR_TD operator, matrix-free."""
    ctx = context()
    x, J, U = _f64(x), _f64(J), _f64(U)
    _check_xJ(x, J, "lanczos_td")
    N, nf, nv = J.shape
    _require(U.dim() == 2 and U.shape[0] == N + N * nv, f"lanczos_td: U has {tuple(U.shape)}, expected {N + N * nv} rows")
    P = U.shape[1]
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with _lanczos_restart(restart_key, restart_partitionable):
        check(
            ctx.lib.gpxb_lanczos_td(ctx.handle, _kind(kind), ptr(x), ptr(J), N, nf, nv,
                                    int(nv if nv_first is None else nv_first), float(ell), float(sigma_targets),
                                    float(sigma_derivs), float(sigma_derivs2), ptr(U), U.stride(0), P, int(m), ptr(alpha),
                                    ptr(beta), ptr(flag), stream_ptr()),
            "gpxb_lanczos_td",
        )
    return alpha, beta, flag


def rpcholesky_td(kind, x, J, n_pivots, ell, key, partitionable=True, want_factor=True, nv_first=None):
    """(F (N + N nv, M) or None, pivots int64 (M,)) -- rpcholesky_TD, gpx/models/gpr_td.py:313-460; rows and pivot
    indices in the reference's order [targets; block 1; block 2] (nv_first < nv: two blocks concatenated in J)."""
    ctx = context()
    x, J = _f64(x), _f64(J)
    _check_xJ(x, J, "rpcholesky_td")
    N, nf, nv = J.shape
    piv = torch.empty(n_pivots, dtype=torch.int64, device=x.device)
    F = torch.empty((N + N * nv, n_pivots), dtype=torch.float64, device=x.device) if want_factor else None
    check(
        ctx.lib.gpxb_rpcholesky_td(ctx.handle, _kind(kind), ptr(x), ptr(J), N, nf, nv,
                                   int(nv if nv_first is None else nv_first), float(ell), int(key[0]),
                                   int(key[1]), int(n_pivots), int(bool(partitionable)), ptr(F) if want_factor else None,
                                   ptr(piv), stream_ptr()),
        "gpxb_rpcholesky_td",
    )
    return F, piv


def gpr_td2_lml_grad(kind, x, J, y, y_derivs, nv_first, ell, sigma_targets, sigma_derivs, sigma_derivs2,
                     mean_mode=MEAN_DATA, want_grad=True):
    """GPR_TD with two derivative blocks concatenated along the variable axis (block 1 = the first nv_first variables)
    -> device tensor [lml, d/d lengthscale, d/d sigma_targets, d/d sigma_derivs, d/d sigma_derivs2]."""
    ctx = context()
    x, J, y, yd = _f64(x), _f64(J), _f64(y), _f64(y_derivs)
    _check_xJ(x, J, "gpr_td")
    N, nf, nv = J.shape
    _require(y.numel() == N and yd.numel() == N * nv and 0 < nv_first <= nv, "shape mismatch between the operands")
    out = torch.empty(5, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_td2_lml_grad(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), ptr(yd), N, nf, nv, int(nv_first),
                                      float(ell), float(sigma_targets), float(sigma_derivs), float(sigma_derivs2),
                                      int(mean_mode), 1 if want_grad else 0, ptr(out), stream_ptr()),
        "gpxb_gpr_td2_lml_grad",
    )
    return out


def gpr_td2_fit(kind, x, J, y, y_derivs, nv_first, ell, sigma_targets, sigma_derivs, sigma_derivs2,
                mean_mode=MEAN_DATA):
    """-> (c (N + N*nv, 1) in the library's row order [targets; (sample, variable)], mu (1,))."""
    ctx = context()
    x, J, y, yd = _f64(x), _f64(J), _f64(y), _f64(y_derivs)
    _check_xJ(x, J, "gpr_td")
    N, nf, nv = J.shape
    _require(y.numel() == N and yd.numel() == N * nv and 0 < nv_first <= nv, "shape mismatch between the operands")
    c = torch.empty((N + N * nv, 1), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_td2_fit(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), ptr(yd), N, nf, nv, int(nv_first),
                                 float(ell), float(sigma_targets), float(sigma_derivs), float(sigma_derivs2),
                                 int(mean_mode), ptr(c), ptr(mu), stream_ptr()),
        "gpxb_gpr_td2_fit",
    )
    return c, mu


def gpr_fit_derivs(kind, x, J, y, ell, sigma):
    ctx = context()
    x, J, y = _f64(x), _f64(J), _f64(y)
    _check_xJ(x, J, "gpr_fit_derivs")
    N, nf, nv = J.shape
    _require(y.numel() == N * nv, f"gpr_fit_derivs: y has {y.numel()} values, expected N*nv = {N * nv}")
    c = torch.empty((N * nv, 1), dtype=torch.float64, device=x.device)
    jc = torch.empty((N, nf), dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_gpr_fit_derivs(ctx.handle, _kind(kind), ptr(x), ptr(J), ptr(y), N, nf, nv, float(ell),
                                    float(sigma), ptr(c), ptr(jc), stream_ptr()),
        "gpxb_gpr_fit_derivs",
    )
    return c, jc


def kmatvec(kind, x_rows, x_all, V, ell, diag_add=0.0, row_offset=None):
    """(K(x_rows, x_all) + diag_add*I) V without storing K.  Passing the same tensor for
    x_rows and x_all puts the shifted diagonal in place (row_offset = 0)."""
    ctx = context()
    x_all = _f64(x_all)
    same = x_rows is x_all
    x_rows = x_all if same else _f64(x_rows)
    V = _f64(V)
    if V.dim() == 1:
        V = V.reshape(-1, 1)
    n_rows, D = x_rows.shape
    N, P = V.shape
    _require(x_all.shape[0] == N and x_all.shape[1] == D, "shape mismatch between the operands")
    if row_offset is None:
        row_offset = 0 if same else -1
    W = torch.empty((n_rows, P), dtype=torch.float64, device=V.device)
    check(
        ctx.lib.gpxb_kmatvec(ctx.handle, _kind(kind), ptr(x_rows), n_rows, int(row_offset), ptr(x_all), N, D,
                             float(ell), float(diag_add), ptr(V), V.stride(0), P, ptr(W), W.stride(0), stream_ptr()),
        "gpxb_kmatvec",
    )
    return W


def rpcholesky(kind, x, n_pivots, ell, key, partitionable=True, want_factor=True):
    """(F (N, M) or None, pivots int64 (M,)) -- gpx/kernels/approximations.py:30-129."""
    ctx = context()
    x = _f64(x)
    N, D = x.shape
    piv = torch.empty(n_pivots, dtype=torch.int64, device=x.device)
    F = torch.empty((N, n_pivots), dtype=torch.float64, device=x.device) if want_factor else None
    check(
        ctx.lib.gpxb_rpcholesky(ctx.handle, _kind(kind), ptr(x), N, D, float(ell), int(key[0]), int(key[1]),
                                int(n_pivots), int(bool(partitionable)), ptr(F), ptr(piv), stream_ptr()),
        "gpxb_rpcholesky",
    )
    return F, piv


def rademacher_probes(key, N, P, partitionable=True):
    ctx = context()
    U = torch.empty((N, P), dtype=torch.float64, device="cuda")
    check(
        ctx.lib.gpxb_rademacher_probes(ctx.handle, int(key[0]), int(key[1]), N, P, int(bool(partitionable)), ptr(U),
                                       stream_ptr()),
        "gpxb_rademacher_probes",
    )
    return U


class _lanczos_restart:
    """Scope in which the library follows the reference's breakdown branch (gpx/lanczos.py:92-108) with `key` the key
    lanczos_tridiagonal carries; key None: breakdown only flags."""

    def __init__(self, key, partitionable=True):
        self.key, self.part = key, partitionable

    def __enter__(self):
        if self.key is not None:
            ctx = context()
            check(ctx.lib.gpxb_lanczos_set_restart(ctx.handle, 1, int(self.key[0]), int(self.key[1]),
                                                   int(bool(self.part))), "gpxb_lanczos_set_restart")

    def __exit__(self, *exc):
        if self.key is not None:
            ctx = context()
            check(ctx.lib.gpxb_lanczos_set_restart(ctx.handle, 0, 0, 0, 1), "gpxb_lanczos_set_restart")
        return False


def lanczos(kind, x, ell, diag_add, U, m, restart_key=None, partitionable=True):
    """(alpha (P, m), beta (P, m), breakdown) for the unit start vectors in U's columns.  restart_key: the key
    lanczos_tridiagonal receives; given, a breakdown restarts with a random orthonormal vector as in the reference."""
    ctx = context()
    x, U = _f64(x), _f64(U)
    N, D = x.shape
    _require(U.dim() == 2 and U.shape[0] == N, f"lanczos: U has {tuple(U.shape)} but x has {N} rows")
    P = U.shape[1]
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with _lanczos_restart(restart_key, partitionable):
        check(
            ctx.lib.gpxb_lanczos(ctx.handle, _kind(kind), ptr(x), N, D, float(ell), float(diag_add), ptr(U),
                                 U.stride(0), P, int(m), ptr(alpha), ptr(beta), ptr(flag), stream_ptr()),
            "gpxb_lanczos",
        )
    return alpha, beta, flag


def lanczos_grad(kind, x, ell, sigma, U, m):
    """(alpha (P, m), beta (P, m), dalpha (2, P, m), dbeta (2, P, m), breakdown): block Lanczos on
    K + (sigma^2+1e-10) I with its forward-mode derivative w.r.t. (lengthscale, sigma)."""
    ctx = context()
    x, U = _f64(x), _f64(U)
    N, D = x.shape
    _require(U.dim() == 2 and U.shape[0] == N, f"lanczos_grad: U has {tuple(U.shape)} but x has {N} rows")
    P = U.shape[1]
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    dalpha = torch.empty((2, P, m), dtype=torch.float64, device=x.device)
    dbeta = torch.empty((2, P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    check(
        ctx.lib.gpxb_lanczos_grad(ctx.handle, _kind(kind), ptr(x), N, D, float(ell), float(sigma), ptr(U), U.stride(0),
                                  P, int(m), ptr(alpha), ptr(beta), ptr(dalpha), ptr(dbeta), ptr(flag), stream_ptr()),
        "gpxb_lanczos_grad",
    )
    return alpha, beta, dalpha, dbeta, flag


def kmatvec_dl(kind, x, V, ell):
    """(d K(x, x)/d lengthscale) V, matrix-free."""
    ctx = context()
    x, V = _f64(x), _f64(V)
    if V.dim() == 1:
        V = V.reshape(-1, 1)
    N, D = x.shape
    _require(V.shape[0] == N, f"kmatvec_dl: V has {V.shape[0]} rows but x has {N}")
    P = V.shape[1]
    W = torch.empty((N, P), dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_kmatvec_dl(ctx.handle, _kind(kind), ptr(x), N, D, float(ell), ptr(V), V.stride(0), P, ptr(W),
                                W.stride(0), stream_ptr()),
        "gpxb_kmatvec_dl",
    )
    return W


def gpr_fit_iter(kind, x, y, ell, sigma, n_pivots, key, mean_mode=MEAN_DATA, partitionable=True, tol=1e-5,
                 atol=1e-7, maxiter=0):
    """(c (N, 1), mu (1,), iterations) -- PCG with the RPCholesky preconditioner."""
    import ctypes

    ctx = context()
    x, y = _f64(x), _f64(y)
    _check_xy(x, y, "gpr_fit_iter")
    N, D = x.shape
    _require(y.numel() == N, "the iterative path supports a single output column")
    c = torch.empty((N, 1), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    iters = ctypes.c_int(0)
    check(
        ctx.lib.gpxb_gpr_fit_iter(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, float(ell), float(sigma),
                                  int(mean_mode), int(n_pivots), int(key[0]), int(key[1]), int(bool(partitionable)),
                                  float(tol), float(atol), int(maxiter), ptr(c), ptr(mu), ctypes.byref(iters),
                                  stream_ptr()),
        "gpxb_gpr_fit_iter",
    )
    return c, mu, iters.value


def gpr_lml_iter(kind, x, y, ell, sigma, n_pivots, key, U, m, mean_mode=MEAN_DATA, partitionable=True, tol=1e-5,
                 atol=1e-7, maxiter=0, restart_key=None):
    """The device parts of the iterative LML in one sweep (gpx/models/_gpr.py:772-821): PCG fit + m-step Lanczos on the
    probes U, the PCG vector riding in the Lanczos block mat-vecs.  -> (c, mu, iterations, alpha, beta, breakdown)."""
    import ctypes

    ctx = context()
    x, y, U = _f64(x), _f64(y), _f64(U)
    _check_xy(x, y, "gpr_lml_iter")
    N, D = x.shape
    _require(y.numel() == N, "the iterative path supports a single output column")
    _require(U.dim() == 2 and U.shape[0] == N, f"gpr_lml_iter: U has {tuple(U.shape)} but x has {N} rows")
    P = U.shape[1]
    c = torch.empty((N, 1), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    iters = ctypes.c_int(0)
    with _lanczos_restart(restart_key, partitionable):
        check(
            ctx.lib.gpxb_gpr_lml_iter(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, float(ell), float(sigma),
                                      int(mean_mode), int(n_pivots), int(key[0]), int(key[1]),
                                      int(bool(partitionable)), float(tol), float(atol), int(maxiter), ptr(U),
                                      U.stride(0), P, int(m), ptr(c), ptr(mu), ctypes.byref(iters), ptr(alpha),
                                      ptr(beta), ptr(flag), stream_ptr()),
            "gpxb_gpr_lml_iter",
        )
    return c, mu, iters.value, alpha, beta, flag


def sgpr_lml(kind, x, y, x_locs, ell, sigma, mean_mode=MEAN_DATA):
    ctx = context()
    x, y, x_locs = _f64(x), _f64(y), _f64(x_locs)
    _check_xy(x, y, "sgpr")
    _require(x_locs.dim() == 2 and x_locs.shape[1] == x.shape[1],
             f"sgpr: x_locs has {tuple(x_locs.shape)} but x has {x.shape[1]} features")
    N, D = x.shape
    T = y.shape[1] if y.dim() > 1 else 1
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_sgpr_lml(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, T, ptr(x_locs), x_locs.shape[0],
                              float(ell), float(sigma), int(mean_mode), ptr(out), stream_ptr()),
        "gpxb_sgpr_lml",
    )
    return out


def sgpr_lml_grad(kind, x, y, x_locs, ell, sigma, mean_mode=MEAN_DATA, want_grad=True):
    """Device tensor [lml/N, d/d ell, d/d sigma] of the SGPR LML with fixed landmarks."""
    ctx = context()
    x, y, x_locs = _f64(x), _f64(y), _f64(x_locs)
    _check_xy(x, y, "sgpr")
    _require(x_locs.dim() == 2 and x_locs.shape[1] == x.shape[1],
             f"sgpr: x_locs has {tuple(x_locs.shape)} but x has {x.shape[1]} features")
    N, D = x.shape
    _require(y.numel() == N, "the SGPR gradient supports a single output column")
    out = torch.empty(3, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_sgpr_lml_grad(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, ptr(x_locs), x_locs.shape[0],
                                   float(ell), float(sigma), int(mean_mode), int(bool(want_grad)), ptr(out),
                                   stream_ptr()),
        "gpxb_sgpr_lml_grad",
    )
    return out


def sgpr_fit_iter(kind, x, y, x_locs, ell, sigma, mean_mode=MEAN_DATA, tol=1e-5, atol=0.0, maxiter=0):
    """(c (M, 1), mu (1,), iterations): CG on the M x M SGPR normal equations, matrix-free."""
    import ctypes

    ctx = context()
    x, y, x_locs = _f64(x), _f64(y), _f64(x_locs)
    _check_xy(x, y, "sgpr")
    _require(x_locs.dim() == 2 and x_locs.shape[1] == x.shape[1],
             f"sgpr: x_locs has {tuple(x_locs.shape)} but x has {x.shape[1]} features")
    N, D = x.shape
    _require(y.numel() == N, "the iterative SGPR path supports a single output column")
    M = x_locs.shape[0]
    c = torch.empty((M, 1), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    iters = ctypes.c_int(0)
    check(
        ctx.lib.gpxb_sgpr_fit_iter(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, ptr(x_locs), M, float(ell),
                                   float(sigma), int(mean_mode), float(tol), float(atol), int(maxiter), ptr(c), ptr(mu),
                                   ctypes.byref(iters), stream_ptr()),
        "gpxb_sgpr_fit_iter",
    )
    return c, mu, iters.value


def sgpr_lml_iter_parts(kind, x, y, x_locs, ell, sigma, U, m, mean_mode=MEAN_DATA, tol=1e-5, atol=0.0, maxiter=0,
                        restart_key=None, restart_partitionable=True):
    """(quad (1,), iterations, alpha (P, m), beta (P, m), breakdown) on the Nystrom operator H + (sigma^2+1e-10) I."""
    import ctypes

    ctx = context()
    x, y, x_locs, U = _f64(x), _f64(y), _f64(x_locs), _f64(U)
    _check_xy(x, y, "sgpr_lml_iter_parts")
    _require(x_locs.dim() == 2 and x_locs.shape[1] == x.shape[1], "sgpr_lml_iter_parts: x_locs feature count differs")
    _require(U.dim() == 2 and U.shape[0] == x.shape[0], "sgpr_lml_iter_parts: U must have one row per data row")
    N, D = x.shape
    _require(y.numel() == N, "the iterative SGPR path supports a single output column")
    P = U.shape[1]
    quad = torch.empty(1, dtype=torch.float64, device=x.device)
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    iters = ctypes.c_int(0)
    with _lanczos_restart(restart_key, restart_partitionable):
        check(
            ctx.lib.gpxb_sgpr_lml_iter_parts(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, ptr(x_locs), x_locs.shape[0],
                                             float(ell), float(sigma), int(mean_mode), float(tol), float(atol), int(maxiter),
                                             ptr(quad), ctypes.byref(iters), ptr(U), U.stride(0), P, int(m), ptr(alpha),
                                             ptr(beta), ptr(flag), stream_ptr()),
            "gpxb_sgpr_lml_iter_parts",
        )
    return quad, iters.value, alpha, beta, flag


def sgpr_lml_iter_grad_parts(kind, x, y, x_locs, ell, sigma, U, m, mean_mode=MEAN_DATA, tol=1e-5, atol=0.0, maxiter=0):
    """sgpr_lml_iter_parts plus what value_and_grad needs: gquad (3,) = [(dK_mn c).g, g.(dK_mm g), c.c] and the
    tangents (2, P, m) of the tridiagonals w.r.t. (lengthscale, sigma).
    -> (quad, gquad, iterations, alpha, beta, dalpha, dbeta, breakdown)."""
    import ctypes

    ctx = context()
    x, y, x_locs, U = _f64(x), _f64(y), _f64(x_locs), _f64(U)
    _check_xy(x, y, "sgpr")
    _require(x_locs.dim() == 2 and x_locs.shape[1] == x.shape[1],
             f"sgpr: x_locs has {tuple(x_locs.shape)} but x has {x.shape[1]} features")
    N, D = x.shape
    _require(y.numel() == N, "the iterative SGPR path supports a single output column")
    _require(U.dim() == 2 and U.shape[0] == N, f"sgpr: U has {tuple(U.shape)} but x has {N} rows")
    P = U.shape[1]
    quad = torch.empty(1, dtype=torch.float64, device=x.device)
    gquad = torch.empty(3, dtype=torch.float64, device=x.device)
    alpha = torch.empty((P, m), dtype=torch.float64, device=x.device)
    beta = torch.empty((P, m), dtype=torch.float64, device=x.device)
    dalpha = torch.empty((2, P, m), dtype=torch.float64, device=x.device)
    dbeta = torch.empty((2, P, m), dtype=torch.float64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    iters = ctypes.c_int(0)
    check(
        ctx.lib.gpxb_sgpr_lml_iter_grad_parts(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, ptr(x_locs),
                                              x_locs.shape[0], float(ell), float(sigma), int(mean_mode), float(tol),
                                              float(atol), int(maxiter), ptr(quad), ptr(gquad), ctypes.byref(iters),
                                              ptr(U), U.stride(0), P, int(m), ptr(alpha), ptr(beta), ptr(dalpha),
                                              ptr(dbeta), ptr(flag), stream_ptr()),
        "gpxb_sgpr_lml_iter_grad_parts",
    )
    return quad, gquad, iters.value, alpha, beta, dalpha, dbeta, flag


def sgpr_lml_grad_locs(kind, x, y, x_locs, ell, sigma, mean_mode=MEAN_DATA):
    """-> ([lml/N, d/d ell, d/d sigma], d(lml/N)/d x_locs (M, D)) for trainable landmark coordinates."""
    ctx = context()
    x, y, x_locs = _f64(x), _f64(y), _f64(x_locs)
    _check_xy(x, y, "sgpr")
    _require(x_locs.dim() == 2 and x_locs.shape[1] == x.shape[1],
             f"sgpr: x_locs has {tuple(x_locs.shape)} but x has {x.shape[1]} features")
    N, D = x.shape
    _require(y.numel() == N, "the SGPR gradient supports a single output column")
    out = torch.empty(3, dtype=torch.float64, device=x.device)
    g = torch.empty_like(x_locs)
    check(
        ctx.lib.gpxb_sgpr_lml_grad_locs(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, ptr(x_locs), x_locs.shape[0],
                                        float(ell), float(sigma), int(mean_mode), ptr(out), ptr(g), stream_ptr()),
        "gpxb_sgpr_lml_grad_locs",
    )
    return out, g


def sgpr_fit(kind, x, y, x_locs, ell, sigma, mean_mode=MEAN_DATA):
    ctx = context()
    x, y, x_locs = _f64(x), _f64(y), _f64(x_locs)
    _check_xy(x, y, "sgpr")
    _require(x_locs.dim() == 2 and x_locs.shape[1] == x.shape[1],
             f"sgpr: x_locs has {tuple(x_locs.shape)} but x has {x.shape[1]} features")
    N, D = x.shape
    T = y.shape[1] if y.dim() > 1 else 1
    M = x_locs.shape[0]
    c = torch.empty((M, T), dtype=torch.float64, device=x.device)
    mu = torch.empty(1, dtype=torch.float64, device=x.device)
    check(
        ctx.lib.gpxb_sgpr_fit(ctx.handle, _kind(kind), ptr(x), ptr(y), N, D, T, ptr(x_locs), M, float(ell),
                              float(sigma), int(mean_mode), ptr(c), ptr(mu), stream_ptr()),
        "gpxb_sgpr_fit",
    )
    return c, mu


def sgpr_predict(kind, x_locs, x, c, mu, ell, sigma, full_covariance=False):
    ctx = context()
    x_locs, x, c = _f64(x_locs), _f64(x), _f64(c)
    M, D = x_locs.shape
    _require(x.dim() == 2 and x.shape[1] == D, f"sgpr_predict: x has {tuple(x.shape)} but the landmarks have {D} features")
    _require(c.shape[0] == M, f"sgpr_predict: c has {c.shape[0]} rows but there are {M} landmarks")
    ns = x.shape[0]
    T = c.shape[1] if c.dim() > 1 else 1
    mean = torch.empty((ns, T), dtype=torch.float64, device=x.device)
    cov = torch.empty((ns, ns), dtype=torch.float64, device=x.device) if full_covariance else None
    check(
        ctx.lib.gpxb_sgpr_predict(ctx.handle, _kind(kind), ptr(x_locs), M, D, ptr(x), ns, ptr(c), T, ptr(mu),
                                  float(ell), float(sigma), ptr(mean), ptr(cov), stream_ptr()),
        "gpxb_sgpr_predict",
    )
    return (mean, cov) if full_covariance else mean


def gather_rows(x, idx):
    ctx = context()
    x = _f64(x)
    _require(idx.dtype == torch.int64 and idx.is_cuda, "gather_rows: idx must be a CUDA int64 tensor")
    M, D = idx.shape[0], x.shape[1]
    out = torch.empty((M, D), dtype=torch.float64, device=x.device)
    check(ctx.lib.gpxb_gather_rows(ctx.handle, ptr(x), D, ptr(idx), M, ptr(out), stream_ptr()), "gpxb_gather_rows")
    return out


def shard_range(N, world, rank):
    """Row range [begin, end) that `rank` of `world` owns (host-only; no GPU needed)."""
    import ctypes

    lib = _lib.load()
    b, e = ctypes.c_longlong(0), ctypes.c_longlong(0)
    check(lib.gpxb_shard_range(int(N), int(world), int(rank), ctypes.byref(b), ctypes.byref(e)), "gpxb_shard_range")
    return b.value, e.value


def broadcast_bytes(raw: bytes, n: int, src: int = 0) -> bytes:
    """Broadcast n bytes over torch.distributed (CPU tensor for gloo, CUDA tensor for nccl)."""
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(raw[:n].ljust(n, b"\0")), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def comm_init_from_torch_distributed():
    """Bootstrap the library's NCCL communicator from an initialised torch.distributed group
    (the unique id travels through a broadcast of 128 bytes)."""
    import ctypes

    import torch.distributed as dist

    ctx = context()
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = (ctypes.c_char * 128)()
    if rank == 0:
        check(ctx.lib.gpxb_comm_unique_id(buf), "gpxb_comm_unique_id")
    raw = broadcast_bytes(bytes(buf), 128, 0)
    cbuf = ctypes.create_string_buffer(raw, 128)
    check(ctx.lib.gpxb_comm_init(ctx.handle, cbuf, rank, world), "gpxb_comm_init")
    return rank, world


def prof_enable(on) -> None:
    """False / True: bracket every DMMA GEMM launch with CUDA events; 2: additionally serialise the look-ahead
    work so that the recorded durations do not overlap."""
    ctx = context()
    check(ctx.lib.gpxb_prof_enable(ctx.handle, int(on)), "gpxb_prof_enable")


def prof_collect():
    """(total ms, total algorithmic flops, launches) of the DMMA GEMM launches since prof_enable(True)."""
    import ctypes

    ctx = context()
    ms, fl, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
    check(ctx.lib.gpxb_prof_collect(ctx.handle, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)),
          "gpxb_prof_collect")
    return ms.value, fl.value, n.value


PHASES = ("matvec", "allgather", "dots", "allreduce", "vecops", "rp_pivot", "rp_column", "sgpr_chunk", "potrf_repl",
          "precond")


def phase_enable(on) -> None:
    """Bracket the phases of the row-sharded paths with CUDA events (measurement aid, see include/gpxb200.h)."""
    ctx = context()
    check(ctx.lib.gpxb_phase_enable(ctx.handle, int(bool(on))), "gpxb_phase_enable")


def phase_collect() -> dict:
    """{phase name: (summed ms, brackets)} since phase_enable(True); synchronises the device."""
    import ctypes

    ctx = context()
    n = 16
    ms, cnt = (ctypes.c_double * n)(), (ctypes.c_int * n)()
    check(ctx.lib.gpxb_phase_collect(ctx.handle, ms, cnt, n), "gpxb_phase_collect")
    return {name: (ms[i], cnt[i]) for i, name in enumerate(PHASES) if cnt[i]}


def launches() -> int:
    return _lib.context().launches()
