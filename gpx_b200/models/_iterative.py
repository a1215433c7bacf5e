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


def fit_iter(state, xd, yd, ell, sigma, n_pivots, key_precond, return_iters=False):
    if n_pivots is None:
        raise ValueError("iterative fitting needs n_pivots (RPCholesky preconditioner size)")
    key = as_key(key_precond if key_precond is not None else gpxargs.key_precond)
    c, mu, iters = ops.gpr_fit_iter(state.kernel.kind, xd, yd, ell, sigma, int(n_pivots), key,
                                    mean_mode(state.mean_function), threefry_partitionable())
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
    subkey, _ = split(key, 2, threefry_partitionable())
    N = xd.shape[0]
    U = ops.rademacher_probes(subkey, N, int(num_evals), threefry_partitionable())
    alpha, beta, flag = ops.lanczos(kind, xd, ell, diag_add, U, int(num_lanczos))
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): the Krylov space is exhausted; "
                           "use fewer Lanczos steps than data points")
    return slq_logdet(alpha.cpu().numpy(), beta.cpu().numpy(), N)


def lml_iter(state, xd, yd, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond):
    m = yd.shape[0]
    c, mu = fit_iter(state, xd, yd, ell, sigma, n_pivots, key_precond)
    r = yd.cpu().numpy() - mu
    mll = -0.5 * float(np.sum(r.T @ c))
    mll -= 0.5 * lanczos_logdet(state.kernel.kind, xd, ell, sigma**2 + 1e-10, num_evals, num_lanczos,
                                as_key(key_lanczos))
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m
