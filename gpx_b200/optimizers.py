"""SciPy L-BFGS-B driver over the unconstrained trainables (mirror of
gpx/optimizers/scipy_optimize.py:35-87 and gpx/optimizers/utils.py:73-131).  The reference
obtains (loss, grad) from jit(value_and_grad(loss)); here the pair comes from one fused device
call and the bijector chain rule is applied on the host."""
from __future__ import annotations

from typing import Callable

import numpy as np
from scipy.optimize import minimize


def ravel_backward_trainables(state):
    """Flat vector of the unconstrained trainables (gpx/optimizers/utils.py:73-94, ravel_pytree order) and
    the (path, shape) list needed to undo it."""
    paths, x0 = [], []
    for path, p in state.flat_params():
        if p.trainable:
            v = np.asarray(p.bijector.backward(p.value), dtype=np.float64)
            paths.append((path, v.shape))
            x0.append(v.reshape(-1))
    return (np.concatenate(x0) if x0 else np.zeros(0)), paths


def _split(paths, xt):
    o = 0
    for path, shape in paths:
        n = int(np.prod(shape)) if shape else 1
        yield path, np.asarray(xt[o:o + n]).reshape(shape)
        o += n


def unravel_forward(state, paths, xt):
    flat = dict(state.flat_params())
    for path, t in _split(paths, xt):
        state = state.with_param_value(path, flat[path].bijector.forward(t))
    return state


def scipy_minimize(state, x, y, loss_and_grad_fn: Callable, callback=None, **opt_kwargs):
    """loss_and_grad_fn(state, x, y) -> (loss, {path: d loss / d constrained value})."""
    x0, paths = ravel_backward_trainables(state)

    def fun(xt):
        st = unravel_forward(state, paths, xt)
        loss, grads = loss_and_grad_fn(st, x, y)
        flat = dict(st.flat_params())
        g = [np.asarray(grads[path] * flat[path].bijector.grad_forward(t), dtype=np.float64).reshape(-1)
             for path, t in _split(paths, xt)]
        return float(loss), np.concatenate(g)

    optres = minimize(fun, x0=x0, method="L-BFGS-B", jac=True, callback=callback, **opt_kwargs)
    return unravel_forward(state, paths, optres.x), optres


class StateCheckpointer:
    """Callback for `scipy_minimize` that writes the model state after every L-BFGS-B iteration (mirror of
    gpx/optimizers/scipy_optimize.py:175-229): `StateCheckpointer(state, "run.npz")` saves `run.000.npz`,
    `run.001.npz`, ... in the `.npz` layout of `ModelState.save`, so that `model.load(...)` can inspect or restart a
    long optimisation.  Pass it as `model.fit(x, y, opt_kwargs=dict(callback=StateCheckpointer(model.state)))`."""

    def __init__(self, state, chk_file: str = "optim_chk.npz") -> None:
        import os

        self.state = state
        self.counter = 0
        name, ext = os.path.splitext(chk_file)
        self.fmt_chk_file = name + ".{:03d}" + ext
        _, self._paths = ravel_backward_trainables(state)

    def __call__(self, x) -> None:
        st = unravel_forward(self.state, self._paths, np.asarray(x, dtype=np.float64))
        st.save(self.fmt_chk_file.format(self.counter))
        self.counter += 1
