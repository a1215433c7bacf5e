"""Iterative (matrix-free) GPR paths: PCG fit and the Lanczos / SLQ log-determinant.

Mirror of gpx/models/_gpr.py:285-310 (_fit_iter), 772-821 (_lml_iter) and
gpx/lanczos.py:136-183 (lanczos_logdet).  The heavy parts are device calls (gpxb_gpr_fit_iter,
gpxb_rademacher_probes, gpxb_lanczos); the host only eigendecomposes the P tiny m x m
tridiagonal matrices (LAPACK, as jnp.linalg.eigh does on the CPU backend) and combines scalars."""
from __future__ import annotations

import numpy as np

from .. import ops
from ..defaults import gpxargs, threefry_partitionable
from ..mean_functions import mean_mode
from ..random import as_key, split


def fit_iter(state, xd, yd, ell, sigma, n_pivots, key_precond, return_iters=False, tol=1e-5, atol=1e-7):
    if n_pivots is None:
        raise ValueError("iterative fitting needs n_pivots (RPCholesky preconditioner size)")
    key = as_key(key_precond if key_precond is not None else gpxargs.key_precond)
    c, mu, iters = ops.gpr_fit_iter(state.kernel.kind, xd, yd, ell, sigma, int(n_pivots), key,
                                    mean_mode(state.mean_function), threefry_partitionable(), tol=tol, atol=atol)
    c, mu = c.cpu().numpy(), np.float64(mu.cpu().numpy()[0])
    return (c, mu, iters) if return_iters else (c, mu)


def slq_logdet(alpha: np.ndarray, beta: np.ndarray, dim_mat: int) -> float:
    """(dim/P) * sum_p sum_k tau_k^2 log(lambda_k)   (gpx/lanczos.py:171-183)."""
    P, m = alpha.shape
    acc = 0.0
    for p in range(P):
        T = np.diag(alpha[p])
        if m > 1:
            T += np.diag(beta[p, : m - 1], 1) + np.diag(beta[p, : m - 1], -1)
        evals, evecs = np.linalg.eigh(T)
        evals = np.maximum(evals, 1e-36)
        acc += float(np.sum(np.square(evecs[0, :]) * np.log(evals)))
    return (dim_mat / P) * acc


def lanczos_logdet(kind, xd, ell, diag_add, num_evals, num_lanczos, key):
    subkey, carried = split(key, 2, threefry_partitionable())
    N = xd.shape[0]
    U = ops.rademacher_probes(subkey, N, int(num_evals), threefry_partitionable())
    alpha, beta, flag = ops.lanczos(kind, xd, ell, diag_add, U, int(num_lanczos), restart_key=carried,
                                    partitionable=threefry_partitionable())
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): the Krylov space is exhausted; "
                           "use fewer Lanczos steps than data points")
    return slq_logdet(alpha.cpu().numpy(), beta.cpu().numpy(), N)


def lml_iter(state, xd, yd, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond, tol=1e-5,
             atol=1e-7):
    """_lml_iter (gpx/models/_gpr.py:772-821).  One device sweep (gpxb_gpr_lml_iter): the PCG search direction rides as
    an extra column of the SLQ block mat-vecs, so fit and log-det share the passes over the recomputed kernel tiles."""
    if n_pivots is None:
        raise ValueError("iterative fitting needs n_pivots (RPCholesky preconditioner size)")
    m = yd.shape[0]
    part = threefry_partitionable()
    subkey, carried = split(as_key(key_lanczos), 2, part)
    U = ops.rademacher_probes(subkey, m, int(num_evals), part)
    kp = as_key(key_precond if key_precond is not None else gpxargs.key_precond)
    c, mu, _, alpha, beta, flag = ops.gpr_lml_iter(state.kernel.kind, xd, yd, ell, sigma, int(n_pivots), kp, U,
                                                   int(num_lanczos), mean_mode(state.mean_function), part, tol=tol,
                                                   atol=atol, restart_key=carried)
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): the Krylov space is exhausted; "
                           "use fewer Lanczos steps than data points")
    c, mu = c.cpu().numpy(), np.float64(mu.cpu().numpy()[0])
    r = yd.cpu().numpy() - mu
    mll = -0.5 * float(np.sum(r.T @ c))
    mll -= 0.5 * slq_logdet(alpha.cpu().numpy(), beta.cpu().numpy(), m)
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def _tridiag(alpha_p, beta_p):
    m = alpha_p.shape[0]
    T = np.diag(alpha_p)
    if m > 1:
        T = T + np.diag(beta_p[: m - 1], 1) + np.diag(beta_p[: m - 1], -1)
    return T


def slq_logdet_grad(alpha, beta, dalpha, dbeta, dim_mat):
    """SLQ log-determinant and its derivatives: d/d theta of sum_k Q0k^2 log lam_k is
    e1^T Dlog(T)[dT] e1 = sum_kl Q0k Q0l (Q^T dT Q)_kl (log lam_k - log lam_l)/(lam_k - lam_l)
    (1/lam_k on the diagonal) -- the closed form of the eigh JVP jax applies (gpx/lanczos.py:171-183)."""
    P, m = alpha.shape
    acc, dacc = 0.0, np.zeros(dalpha.shape[0])
    for p in range(P):
        lam, Q = np.linalg.eigh(_tridiag(alpha[p], beta[p]))
        lam = np.maximum(lam, 1e-36)
        q0 = Q[0, :]
        acc += float(np.sum(np.square(q0) * np.log(lam)))
        dl = lam[:, None] - lam[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            dd = np.where(np.abs(dl) > 1e-14 * np.abs(lam[:, None]),
                          (np.log(lam)[:, None] - np.log(lam)[None, :]) / dl, 1.0 / lam[:, None])
        for q in range(dalpha.shape[0]):
            M = Q.T @ _tridiag(dalpha[q, p], dbeta[q, p]) @ Q
            dacc[q] += float(q0 @ (M * dd) @ q0)
    return (dim_mat / P) * acc, (dim_mat / P) * dacc


def lml_iter_grad(state, xd, yd, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond, tol=1e-5,
                  atol=1e-7):
    """(lml, d lml/d lengthscale, d lml/d sigma) of _lml_iter as jit(value_and_grad) propagates it
    (gpx/models/_gpr.py:772-821 under gpx/optimizers/scipy_optimize.py:72-78): the CG solve is a
    lax.custom_linear_solve, so d(r.c) = -c^T dA c with c the computed PCG solution; the SLQ term is
    differentiated through the Lanczos recursion (gpxb_lanczos_grad) and the eigendecomposition."""
    import torch

    kind = state.kernel.kind
    N = yd.shape[0]
    c, mu = fit_iter(state, xd, yd, ell, sigma, n_pivots, key_precond, tol=tol, atol=atol)
    r = yd.cpu().numpy() - mu
    cd = torch.as_tensor(c, dtype=torch.float64, device=xd.device).reshape(-1, 1)
    c_dK_c = float((cd * ops.kmatvec_dl(kind, xd, cd, ell)).sum().cpu())
    subkey, _ = split(as_key(key_lanczos), 2, threefry_partitionable())
    U = ops.rademacher_probes(subkey, N, int(num_evals), threefry_partitionable())
    alpha, beta, dalpha, dbeta, flag = ops.lanczos_grad(kind, xd, ell, sigma, U, int(num_lanczos))
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): the Krylov space is exhausted; "
                           "use fewer Lanczos steps than data points")
    ld, dld = slq_logdet_grad(alpha.cpu().numpy(), beta.cpu().numpy(), dalpha.cpu().numpy(), dbeta.cpu().numpy(), N)
    mll = (-0.5 * float(np.sum(r.T @ c)) - 0.5 * ld - N * 0.5 * np.log(2.0 * np.pi)) / N
    g_ell = (0.5 * c_dK_c - 0.5 * dld[0]) / N
    g_sigma = (sigma * float(np.sum(c * c)) - 0.5 * dld[1]) / N
    return mll, g_ell, g_sigma


def fit_derivs_iter(kind, xd, Jd, yd, ell, sigma, n_pivots, key_precond, return_iters=False):
    """_fit_derivs_iter (gpx/models/_gpr.py:347-390): PCG on the matrix-free Hessian operator."""
    if n_pivots is None:
        raise ValueError("iterative fitting needs n_pivots (RPCholesky preconditioner size)")
    key = as_key(key_precond if key_precond is not None else gpxargs.key_precond)
    c, jc, iters = ops.gpr_fit_derivs_iter(kind, xd, Jd, yd, ell, sigma, int(n_pivots), key,
                                           threefry_partitionable())
    return (c, jc, iters) if return_iters else (c, jc)


def lml_derivs_iter(kind, xd, Jd, yd, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond):
    """_lml_derivs_iter (gpx/models/_gpr.py:873-935), zero mean."""
    n = yd.numel()
    c, _ = fit_derivs_iter(kind, xd, Jd, yd, ell, sigma, n_pivots, key_precond)
    quad = float((yd.reshape(-1) * c).sum().cpu())
    subkey, carried = split(as_key(key_lanczos), 2, threefry_partitionable())
    U = ops.rademacher_probes(subkey, n, int(num_evals), threefry_partitionable())
    alpha, beta, flag = ops.lanczos_derivs(kind, xd, Jd, ell, sigma, U, int(num_lanczos), restart_key=carried,
                                           restart_partitionable=threefry_partitionable())
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): use fewer Lanczos steps than rows")
    logdet = slq_logdet(alpha.cpu().numpy(), beta.cpu().numpy(), n)
    return (-0.5 * quad - 0.5 * logdet - n * 0.5 * np.log(2.0 * np.pi)) / n


def lml_derivs_iter_grad(kind, xd, Jd, yd, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond):
    """(lml, d/d lengthscale, d/d sigma) of _lml_derivs_iter as jit(value_and_grad) propagates it: same structure
    as lml_iter_grad on the matrix-free Hessian operator (gpxb_hess_matvec_dl, gpxb_lanczos_grad_derivs)."""
    n = yd.numel()
    c, _ = fit_derivs_iter(kind, xd, Jd, yd, ell, sigma, n_pivots, key_precond)
    quad = float((yd.reshape(-1) * c).sum().cpu())
    c_dA_c = float((c * ops.hess_matvec_dl(kind, xd, Jd, c, ell)).sum().cpu())
    subkey, _ = split(as_key(key_lanczos), 2, threefry_partitionable())
    U = ops.rademacher_probes(subkey, n, int(num_evals), threefry_partitionable())
    alpha, beta, dalpha, dbeta, flag = ops.lanczos_grad_derivs(kind, xd, Jd, ell, sigma, U, int(num_lanczos))
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): use fewer Lanczos steps than rows")
    ld, dld = slq_logdet_grad(alpha.cpu().numpy(), beta.cpu().numpy(), dalpha.cpu().numpy(), dbeta.cpu().numpy(), n)
    mll = (-0.5 * quad - 0.5 * ld - n * 0.5 * np.log(2.0 * np.pi)) / n
    g_ell = (0.5 * c_dA_c - 0.5 * dld[0]) / n
    g_sigma = (sigma * float((c * c).sum().cpu()) - 0.5 * dld[1]) / n
    return mll, g_ell, g_sigma
