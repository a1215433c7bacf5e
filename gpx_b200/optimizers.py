"""SciPy L-BFGS-B driver over the unconstrained trainables (mirror of
gpx/optimizers/scipy_optimize.py:35-87 and gpx/optimizers/utils.py:73-131).  The reference
obtains (loss, grad) from jit(value_and_grad(loss)); here the pair comes from one fused device
call and the bijector chain rule is applied on the host."""
from __future__ import annotations

from typing import Callable

import numpy as np
from scipy.optimize import minimize


def ravel_backward_trainables(state):
    paths, x0 = [], []
    for path, p in state.flat_params():
        if p.trainable:
            if p.value.shape != ():
                raise NotImplementedError("only scalar trainable hyper-parameters are supported")
            paths.append(path)
            x0.append(float(p.bijector.backward(p.value)))
    return np.asarray(x0, dtype=np.float64), paths


def unravel_forward(state, paths, xt):
    for path, t in zip(paths, xt):
        p = dict(state.flat_params())[path]
        state = state.with_param_value(path, p.bijector.forward(t))
    return state


def scipy_minimize(state, x, y, loss_and_grad_fn: Callable, callback=None, **opt_kwargs):
    """loss_and_grad_fn(state, x, y) -> (loss, {path: d loss / d constrained value})."""
    x0, paths = ravel_backward_trainables(state)

    def fun(xt):
        st = unravel_forward(state, paths, xt)
        loss, grads = loss_and_grad_fn(st, x, y)
        flat = dict(st.flat_params())
        g = np.array([grads[path] * float(flat[path].bijector.grad_forward(t)) for path, t in zip(paths, xt)])
        return float(loss), g

    optres = minimize(fun, x0=x0, method="L-BFGS-B", jac=True, callback=callback, **opt_kwargs)
    return unravel_forward(state, paths, optres.x), optres
