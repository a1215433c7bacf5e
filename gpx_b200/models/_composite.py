"""Exact GPR with a Sum / Prod composite kernel (gpx/kernels/kernels.py:1794-1942 under the dense paths of
gpx/models/_gpr.py:262-282, 434-463, 737-769 and jit(value_and_grad(loss))).  Composition of device calls:
component matrices by the fused Gram kernel, combined by gpxb_elementwise; factorisation, solves and inverse
by the blocked FP64 Cholesky; the gradient w.r.t. every leaf lengthscale by gpxb_gram_dl_contract with the
weights  W = alpha alpha^T - A^-1  (Sum)  or  W o (product of the other factors)  (Prod) -- no dK is formed."""
from __future__ import annotations

import numpy as np

from .. import ops
from ..kernels import Linear, _Composite, _to_device
from ..mean_functions import mean_mode


def is_composite(kernel) -> bool:
    """Kernels served by this generic dense path: Sum / Prod trees and the parameter-free Linear kernel."""
    return isinstance(kernel, (_Composite, Linear))


def _slice_leaf(k, xd):
    return xd if k.active_dims is None else xd[:, k.active_dims].contiguous()


def _center(state, y):
    yd = _to_device(y)
    if yd.dim() == 1:
        yd = yd.reshape(-1, 1)
    mu = float(yd.mean().cpu()) if mean_mode(state.mean_function) == 1 else 0.0
    return (yd - mu).T.contiguous(), mu  # (T, N)


def _factor(state, xd):
    kernel = state.kernel
    sigma = float(np.asarray(state.params["sigma"].value))
    A = kernel.k_device(xd, xd, state.params["kernel_params"])
    A.diagonal().add_(sigma * sigma + 1e-10)
    L, inv = ops.potrf(A)
    return L, inv, sigma


def lml_and_grad(state, x, y, want_grad):
    """(lml, {param path: d lml / d value})."""
    xd = _to_device(x)
    rT, _ = _center(state, y)
    N = xd.shape[0]
    L, inv, sigma = _factor(state, xd)
    logdet_half = float(ops.logdet(L).cpu()[0])
    cy = ops.trsm(L, inv, rT.clone(), trans=True)
    lml = (-0.5 * float((cy * cy).sum().cpu()) - logdet_half - N * 0.5 * np.log(2.0 * np.pi)) / N
    if not want_grad:
        return lml, {}
    alphaT = ops.trsm(L, inv, cy, trans=False)                       # (T, N)
    Wn = ops.potri(L, inv)                                             # A^-1, lower
    tr_inv = float(Wn.diagonal().sum().cpu())
    ops.gemm(alphaT, alphaT, transa=True, alpha=-1.0, beta=1.0, C=Wn, lower_only=True)   # Wn = A^-1 - sum_t a a^T
    grads = {("sigma",): sigma * (float((alphaT * alphaT).sum().cpu()) - tr_inv) / N}
    kp = state.params["kernel_params"]
    leaves = state.kernel.leaves(kp) if isinstance(state.kernel, _Composite) else []
    for path, k, lp in leaves:
        if "lengthscale" not in lp:  # Linear: no parameters, no gradient entry
            continue
        xl = _slice_leaf(k, xd)
        Y = Wn
        w = _other_factors(state.kernel, kp, path, xd)
        if w is not None:
            Y = ops.elementwise("mul", Wn, w)
        c = float(ops.gram_dl_contract(k.kind, xl, xl, k.lengthscale(lp), Y, sym_lower=True).cpu()[0])
        grads[("kernel_params",) + path + ("lengthscale",)] = -0.5 * c / N
    return lml, grads


def _other_factors(kernel, params, path, xd):
    """d K / d (leaf at `path`) = (d leaf) o F: returns F (device matrix) or None when F = 1."""
    F = None
    node, p = kernel, params
    for name in path:
        other = "kernel2" if name == "kernel1" else "kernel1"
        if node.op == "prod":
            M = getattr(node, other).k_device(xd, xd, p[other])
            F = M if F is None else ops.elementwise("mul", F, M, out=F)
        node, p = getattr(node, name), p[name]
    return F


def fit(state, x, y):
    xd = _to_device(x)
    rT, mu = _center(state, y)
    L, inv, _ = _factor(state, xd)
    c = ops.trsm(L, inv, ops.trsm(L, inv, rT.clone(), trans=True), trans=False)
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), c=c.T.contiguous().cpu().numpy(),
                             mu=np.float64(mu), is_fitted=True, is_fitted_derivs=False))


def predict(state, x, full_covariance=False):
    xt, xs = _to_device(state.x_train), _to_device(x)
    kp = state.params["kernel_params"]
    Ks = state.kernel.k_device(xs, xt, kp)                              # (ns, N)
    c = _to_device(state.c)
    if c.dim() == 1:
        c = c.reshape(-1, 1)
    mean = ops.gemm(Ks, c) + float(state.mu)
    if not full_covariance:
        return mean.cpu().numpy()
    L, inv, _ = _factor(state, xt)
    G = ops.trsm(L, inv, Ks, trans=True)                               # Ks L^-T
    cov = state.kernel.k_device(xs, xs, kp)
    ops.gemm(G, G, transb=True, alpha=-1.0, beta=1.0, C=cov)
    return mean.cpu().numpy(), cov.cpu().numpy()


# ---- composite kernels in the matrix-free paths (one GPU) ---------------------------------------------------------
def _leaf_desc(state, xd):
    """(kinds, lengthscales, per-leaf sliced inputs, postfix program) of the state's kernel for the device calls."""
    kernel = state.kernel
    kp = state.params["kernel_params"]
    if isinstance(kernel, _Composite):
        leaves, prog = kernel.program(kp)
    else:  # a lone Linear kernel
        leaves, prog = [(kernel, kp)], [0]
    kinds = [k.kind for k, _ in leaves]
    ells = [1.0 if k.kind == "linear" else k.lengthscale(p) for k, p in leaves]
    xl = [_slice_leaf(k, xd) for k, _ in leaves]
    if not ops.gram_composite_fits([t.shape[1] for t in xl]):
        raise NotImplementedError("the iterative / RPCholesky paths of a composite kernel need the operand tiles of all "
                                  "leaves in shared memory (sum over leaves of 2*128*(D+pad)*8 bytes <= 227 KB)")
    return kinds, ells, xl, prog


def fit_iter(state, x, y, n_pivots, key_precond, tol=1e-5, atol=1e-7, return_iters=False):
    """_fit_iter (gpx/models/_gpr.py:285-310) with a Sum / Prod kernel: PCG on the row-chunked composite operator with
    the composite RPCholesky preconditioner (gpxb_gpr_iter_composite)."""
    from ..defaults import gpxargs, threefry_partitionable
    from ..random import as_key

    if n_pivots is None:
        raise ValueError("iterative fitting needs n_pivots (RPCholesky preconditioner size)")
    xd, yd = _to_device(x), _to_device(y)
    kinds, ells, xl, prog = _leaf_desc(state, xd)
    sigma = float(np.asarray(state.params["sigma"].value))
    key = as_key(key_precond if key_precond is not None else gpxargs.key_precond)
    r = ops.gpr_iter_composite(kinds, ells, xl, prog, sigma, y=yd.reshape(-1), n_pivots=int(n_pivots), key=key,
                               mean_mode=mean_mode(state.mean_function), partitionable=threefry_partitionable(),
                               tol=tol, atol=atol)
    c, mu = r["c"].cpu().numpy(), np.float64(r["mu"].cpu().numpy()[0])
    return (c, mu, r["iters"]) if return_iters else (c, mu)


def lml_iter(state, x, y, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond, tol=1e-5, atol=1e-7):
    """_lml_iter (gpx/models/_gpr.py:772-821) with a Sum / Prod kernel."""
    from ..defaults import threefry_partitionable
    from ..random import as_key, split
    from ._iterative import slq_logdet

    xd, yd = _to_device(x), _to_device(y)
    N = xd.shape[0]
    c, mu = fit_iter(state, x, y, n_pivots, key_precond, tol, atol)
    kinds, ells, xl, prog = _leaf_desc(state, xd)
    sigma = float(np.asarray(state.params["sigma"].value))
    part = threefry_partitionable()
    subkey, carried = split(as_key(key_lanczos), 2, part)
    U = ops.rademacher_probes(subkey, N, int(num_evals), part)
    r = ops.gpr_iter_composite(kinds, ells, xl, prog, sigma, U=U, m=int(num_lanczos), restart_key=carried,
                               restart_partitionable=part)
    if int(r["flag"].cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): use fewer Lanczos steps than data points")
    res = yd.cpu().numpy().reshape(-1, 1) - mu
    mll = -0.5 * float(np.sum(res.T @ c))
    mll -= 0.5 * slq_logdet(r["alpha"].cpu().numpy(), r["beta"].cpu().numpy(), N)
    mll -= N * 0.5 * np.log(2.0 * np.pi)
    return mll / N


def fit_state_iter(state, x, y, n_pivots, key_precond):
    c, mu = fit_iter(state, x, y, n_pivots, key_precond)
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), c=c, mu=mu, is_fitted=True,
                             is_fitted_derivs=False))


def select_locs(state, xd, n_locs, key):
    """RPCholesky landmark selection with a Sum / Prod kernel (gpx/models/_sgpr_rpchol.py:36-67 with
    gpx/kernels/approximations.py:30-129): -> (x_locs device (M, D) raw rows, pivots)."""
    from ..defaults import threefry_partitionable
    from ..random import as_key

    kinds, ells, xl, prog = _leaf_desc(state, xd)
    _, piv = ops.rpcholesky_composite(kinds, ells, xl, prog, int(n_locs), as_key(key), threefry_partitionable(),
                                      want_factor=False)
    return ops.gather_rows(xd, piv), piv
