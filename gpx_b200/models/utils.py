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
