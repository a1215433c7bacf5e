"""Host control flow around the optimiser (mirror of gpx/models/utils.py:77-167)."""
from __future__ import annotations

import numpy as np

from ..optimizers import scipy_minimize


def _rng_from_key(key):
    k = np.asarray(key, dtype=np.uint64)
    return np.random.default_rng(int((k[0] << np.uint64(32)) | k[1]))


def randomized_minimization(key, state, x, y, loss_and_grad_fn, num_restarts=0, return_history=False,
                            opt_kwargs=None):
    """gpx/models/utils.py:98-167: optimise once from the current state, then num_restarts
    times from parameters sampled from their priors; keep the best."""
    opt_kwargs = {} if opt_kwargs is None else dict(opt_kwargs)
    if num_restarts > 0 and key is None:
        raise ValueError("randomized restarts require a valid key")
    states, optres_all, losses = [], [], []
    st, res = scipy_minimize(state, x, y, loss_and_grad_fn, **opt_kwargs)
    states.append(st), optres_all.append(res), losses.append(res.fun)
    rng = _rng_from_key(key) if num_restarts > 0 else None
    for _ in range(num_restarts):
        st0 = state.randomize(rng)
        st, res = scipy_minimize(st0, x, y, loss_and_grad_fn, **opt_kwargs)
        states.append(st), optres_all.append(res), losses.append(res.fun)
    best = int(np.nanargmin(np.asarray(losses, dtype=np.float64)))
    if return_history:
        return states[best], optres_all[best], states, losses
    return states[best], optres_all[best]


def sample(key, mean, cov_dev, n_samples=1):
    """Draws from N(mean, cov) the way gpx/models/utils.py:52-75 does, quirks included: for a 2-D `mean` (n, dims)
    the key is split once per column, each column is drawn with the CARRIED half of the split, and only the draw of
    the LAST column is returned ((n_samples, n)); a 1-D mean gives one draw of shape (n,), ignoring n_samples.
    jax.random.multivariate_normal (default method "cholesky") = mean + normal(key, (n_samples, n)) @ chol(cov)^T:
    the factorisation and the product run on the device (gpxb_potrf, gpxb_gemm); the normal deviates are a small
    host draw with jax.random's bit layout (gpx_b200/random.py)."""
    import torch

    from .. import ops
    from ..defaults import threefry_partitionable
    from ..random import as_key, normal, split

    part = threefry_partitionable()
    key = as_key(key)
    mean = np.asarray(mean, dtype=np.float64)
    L, _ = ops.potrf(cov_dev.clone())
    L = torch.tril(L)  # potrf leaves the strict upper triangle untouched

    def mvn(k, mu, shape):
        z = normal(k, tuple(shape) + (mu.shape[-1],), part).reshape(-1, mu.shape[-1])
        zd = torch.as_tensor(np.ascontiguousarray(z), dtype=torch.float64, device=L.device)
        out = ops.gemm(zd, L, transb=True).cpu().numpy() + mu
        return out.reshape(tuple(shape) + (mu.shape[-1],))

    if mean.ndim > 1:
        out = None
        for dim in range(mean.shape[1]):
            key = split(key, 2, part)[1]  # `subkey, key = random.split(key)`; the draw uses `key`
            out = mvn(key, mean[:, dim], (int(n_samples),))
        return out
    return mvn(key, mean, ())
