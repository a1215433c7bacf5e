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
