"""SGPR (Projected Processes) with a Sum / Prod / Linear kernel: LML, its gradient, fit and predict of
gpx/models/_sgpr.py:39-61, 319-339, 434-467, 659-697 for the composite kernels of gpx/kernels/kernels.py:1794-1942.
Composition of device calls only (fused Gram kernel per leaf, gpxb_elementwise, DMMA GEMM / SYRK, blocked Cholesky,
triangular solves, gpxb_gram_dl_contract); the single-kernel SGPR keeps its fused entry points (csrc/sgpr.cu).

Woodbury form of the reference's N x N Cholesky, with K' = K_mm + 1e-10 I = L_m L_m^T, G = L_m^-1 K_mn,
B = s2 I + G G^T = L_B L_B^T, s2 = sigma^2 + 1e-10, C = G^T G + s2 I (never formed):

    lml N = -1/2 (r.r - |L_B^-1 G r|^2) / s2 - 1/2 [(N - M) log s2 + 2 sum log L_B,ii] - N/2 log 2 pi

Gradient: with P = K'^-1 K_mn, beta = C^-1 r, W = beta beta^T - C^-1,  d(lml N) = 1/2 tr(W dC) and
dC = dK_nm P + P^T dK_mn - P^T dK_mm P + 2 sigma dsigma I  give the adjoints

    Y_nm = beta (P beta)^T - G^T B^-1 L_m^-1          (N x M;  P C^-1 = L_m^-T B^-1 G)
    Y_mm = -1/2 [ (P beta)(P beta)^T - P (G^T B^-1 L_m^-1) ]
    d(lml N)/d sigma = sigma (beta.beta - (N - tr(B^-1 G G^T)) / s2)

and every leaf lengthscale gets  sum Y_nm o F_nm o d leaf_nm + sum Y_mm o F_mm o d leaf_mm  from the contraction
epilogue of the Gram kernel (F = product of the other factors along the path, 1 for sums); no dK is formed.
Memory: a few N x M matrices (the composite path is meant for M << N that fit; the fused single-kernel path streams
row chunks instead)."""
from __future__ import annotations

import numpy as np

from .. import ops
from ..kernels import _Composite, _to_device
from ..mean_functions import mean_mode
from ._composite import _center, _slice_leaf


def _sigma(state):
    return float(np.asarray(state.params["sigma"].value))


def _locs_raw(state):
    return _to_device(np.asarray(state.params["x_locs"].value))


def _woodbury(state, xd, xl):
    """L_m, G^T (N x M), L_B and their leaf inverses."""
    kp = state.params["kernel_params"]
    s2 = _sigma(state) ** 2 + 1e-10
    Kmm = state.kernel.k_device(xl, xl, kp)
    Kmm.diagonal().add_(1e-10)
    Lm, invm = ops.potrf(Kmm)
    GT = ops.trsm(Lm, invm, state.kernel.k_device(xd, xl, kp), trans=True)      # K_nm L_m^-T = G^T
    B = ops.gemm(GT, GT, transa=True, lower_only=True)                          # G G^T (lower)
    B.diagonal().add_(s2)
    LB, invB = ops.potrf(B)
    return Lm, invm, GT, LB, invB, s2


def lml_and_grad(state, x, y, want_grad, x_locs=None):
    """(lml, {param path: d lml / d value}) for fixed landmarks."""
    xd, xl = _to_device(x), (_locs_raw(state) if x_locs is None else x_locs)
    rT, _ = _center(state, y)                                                    # (T, N)
    N, M = xd.shape[0], xl.shape[0]
    Lm, invm, GT, LB, invB, s2 = _woodbury(state, xd, xl)
    tT = ops.trsm(LB, invB, ops.gemm(rT, GT), trans=True)                        # (L_B^-1 G r)^T, (T, M)
    rr = float((rT * rT).sum().cpu())
    quad = (rr - float((tT * tT).sum().cpu())) / s2
    logdet = (N - M) * np.log(s2) + 2.0 * float(ops.logdet(LB).cpu()[0])
    lml = (-0.5 * quad - 0.5 * logdet - N * 0.5 * np.log(2.0 * np.pi)) / N
    if not want_grad:
        return lml, {}
    sigma = _sigma(state)
    uT = ops.trsm(LB, invB, tT.clone(), trans=False)                             # (B^-1 G r)^T
    betaT = ops.gemm(uT, GT, transb=True, alpha=-1.0 / s2, beta=1.0 / s2, C=rT.clone())   # beta^T = (r - G^T u)^T / s2
    pT = ops.trsm(Lm, invm, ops.gemm(betaT, GT), trans=False)                    # (P beta)^T, (T, M)
    GBi = ops.trsm(LB, invB, ops.trsm(LB, invB, GT.clone(), trans=True), trans=False)     # G^T B^-1, (N, M)
    tr_BiGGt = float((GBi * GT).sum().cpu())
    H = ops.trsm(Lm, invm, GBi, trans=False)                                     # G^T B^-1 L_m^-1 (in place)
    PT = ops.trsm(Lm, invm, GT.clone(), trans=False)                             # P^T = G^T L_m^-1, (N, M)
    Ymm = ops.gemm(PT, H, transa=True)                                           # P H
    ops.gemm(pT, pT, transa=True, alpha=-0.5, beta=0.5, C=Ymm)                   # -1/2 (p p^T - P H)
    Ynm = ops.gemm(betaT, pT, transa=True, alpha=1.0, beta=-1.0, C=H)            # beta p^T - H
    grads = {("sigma",): sigma * (float((betaT * betaT).sum().cpu()) - (N - tr_BiGGt) / s2) / N}
    kp = state.params["kernel_params"]
    leaves = state.kernel.leaves(kp) if isinstance(state.kernel, _Composite) else []
    for path, k, lp in leaves:
        if "lengthscale" not in lp:
            continue
        ell = k.lengthscale(lp)
        a, b = _slice_leaf(k, xd), _slice_leaf(k, xl)
        Fnm, Fmm = _other_factors(state.kernel, kp, path, xd, xl), _other_factors(state.kernel, kp, path, xl, xl)
        Wnm = Ynm if Fnm is None else ops.elementwise("mul", Ynm, Fnm)
        Wmm = Ymm if Fmm is None else ops.elementwise("mul", Ymm, Fmm)
        g = float(ops.gram_dl_contract(k.kind, a, b, ell, Wnm).cpu()[0]) + \
            float(ops.gram_dl_contract(k.kind, b, b, ell, Wmm, sym_lower=True).cpu()[0])  # Y_mm is symmetric
        grads[("kernel_params",) + path + ("lengthscale",)] = g / N
    return lml, grads


def _other_factors(kernel, params, path, x1d, x2d):
    """d K(x1, x2) / d (leaf at `path`) = (d leaf) o F: returns F (device matrix) or None when F = 1."""
    F = None
    node, p = kernel, params
    for name in path:
        other = "kernel2" if name == "kernel1" else "kernel1"
        if node.op == "prod":
            M = getattr(node, other).k_device(x1d, x2d, p[other])
            F = M if F is None else ops.elementwise("mul", F, M, out=F)
        node, p = getattr(node, name), p[name]
    return F


def _fit_matrix(state, xd, xl):
    """Cholesky factor of  sigma^2 K_mm + K_mn K_nm + 1e-10 I  (_sgpr.py:39-61) and K_nm."""
    kp = state.params["kernel_params"]
    sigma = _sigma(state)
    Knm = state.kernel.k_device(xd, xl, kp)
    C = state.kernel.k_device(xl, xl, kp)
    ops.gemm(Knm, Knm, transa=True, alpha=1.0, beta=sigma * sigma, C=C)
    C.diagonal().add_(1e-10)
    LC, invC = ops.potrf(C)
    return LC, invC, Knm


def fit(state, x, y, x_locs=None):
    """_sgpr.py:319-339 (the reference's LU solve is a Cholesky solve here)."""
    xd, xl = _to_device(x), (_locs_raw(state) if x_locs is None else x_locs)
    rT, mu = _center(state, y)
    LC, invC, Knm = _fit_matrix(state, xd, xl)
    cT = ops.gemm(rT, Knm)                                                       # (K_mn r)^T, (T, M)
    ops.trsm(LC, invC, cT, trans=True)
    ops.trsm(LC, invC, cT, trans=False)
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), c=cT.T.contiguous().cpu().numpy(),
                             mu=np.float64(mu), is_fitted=True, is_fitted_derivs=False))


def predict(state, x, full_covariance=False, x_locs=None):
    """_sgpr.py:434-467: mu + K_*m c;  K_** - G*^T G* + sigma^2 H*^T H*, where -- as the reference writes it
    (:457, `_A_lhs(x1=x_locs, x2=x)`) -- H* = chol(sigma^2 K_mm + K_m* K_*m + 1e-10 I)^-1 K_m* is built from the
    PREDICTION inputs, not from the training set."""
    xs, xl = _to_device(x), (_locs_raw(state) if x_locs is None else x_locs)
    kp = state.params["kernel_params"]
    Ksm = state.kernel.k_device(xs, xl, kp)                                      # (ns, M)
    c = _to_device(state.c)
    if c.dim() == 1:
        c = c.reshape(-1, 1)
    mean = ops.gemm(Ksm, c) + float(state.mu)
    if not full_covariance:
        return mean.cpu().numpy()
    sigma = _sigma(state)
    Kmm = state.kernel.k_device(xl, xl, kp)
    Kmm.diagonal().add_(1e-10)
    Lm, invm = ops.potrf(Kmm)
    LC, invC, _ = _fit_matrix(state, xs, xl)
    Gs = ops.trsm(Lm, invm, Ksm.clone(), trans=True)
    Hs = ops.trsm(LC, invC, Ksm.clone(), trans=True)
    cov = state.kernel.k_device(xs, xs, kp)
    ops.gemm(Gs, Gs, transb=True, alpha=-1.0, beta=1.0, C=cov)
    ops.gemm(Hs, Hs, transb=True, alpha=sigma * sigma, beta=1.0, C=cov)
    return mean.cpu().numpy(), cov.cpu().numpy()
