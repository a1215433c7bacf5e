"""SGPR with RPCholesky landmark selection (mirror of gpx/models/sgpr_rpchol.py:63-106,
300-350, 700-811).  The pivots are re-drawn from ``state.key`` with the CURRENT
hyper-parameters on every LML / fit call, as the reference does
(gpx/models/_sgpr_rpchol.py:36-67, 178-208)."""
from __future__ import annotations

import time
from typing import Callable

import numpy as np

from .. import ops
from ..defaults import gpxargs, threefry_partitionable
from ..mean_functions import data_mean, mean_mode
from ..parameters import ModelState
from ..random import PRNGKey, as_key
from . import _composite, _composite_sparse
from . import sgpr as _sgpr
from .base import BaseGP
from .gpr import _hyper, _xy


def _is_comp(state):
    return _composite.is_composite(state.kernel)


def _select_locs(state, xd):
    if _is_comp(state):  # Sum / Prod kernel: the pivot column is evaluated pair by pair (gpxb_rpcholesky_composite)
        return _composite.select_locs(state, xd, state.n_locs, state.key)
    kernel, ell, _ = _hyper(state)
    _, piv = ops.rpcholesky(kernel.kind, xd, int(state.n_locs), ell, as_key(state.key), threefry_partitionable(),
                            want_factor=False)
    return ops.gather_rows(xd, piv), piv


def log_marginal_likelihood(state, x, y, iterative=False, num_evals=gpxargs.num_evals,
                            num_lanczos=gpxargs.num_lanczos, key_lanczos=gpxargs.key_lanczos, **_ignored):
    """gpx/models/sgpr_rpchol.py:63-106 -> _sgpr_rpchol._lml_dense (178-208)."""
    if _is_comp(state):
        if iterative:
            raise NotImplementedError("composite kernels in SGPR_RPChol: the dense paths only")
        from ..kernels import _to_device

        x_locs, _ = _select_locs(state, _to_device(x))
        return _composite_sparse.lml_and_grad(state, x, y, want_grad=False, x_locs=x_locs)[0]
    kernel, ell, sigma = _hyper(state)
    xd, yd = _xy(state, x, y)
    x_locs, _ = _select_locs(state, xd)
    if iterative:
        return _sgpr._lml_iter_with_locs(state, xd, yd, x_locs, num_evals, num_lanczos, key_lanczos)
    out = ops.sgpr_lml(kernel.kind, xd, yd, x_locs, ell, sigma, mean_mode(state.mean_function))
    return float(out.cpu().numpy()[0])


def neg_log_marginal_likelihood(state, x, y, **kwargs):
    return -log_marginal_likelihood(state, x, y, **kwargs)


def loss_and_grad(state, x, y, iterative=False, **kwargs):
    """Pivots are re-drawn with the current hyper-parameters and treated as constants (integer
    indices carry no gradient in the reference either), then the SGPR gradient applies."""
    if _is_comp(state):
        if iterative:
            raise NotImplementedError("composite kernels in SGPR_RPChol: the dense paths only")
        from ..kernels import _to_device

        x_locs, _ = _select_locs(state, _to_device(x))
        lml, g = _composite_sparse.lml_and_grad(state, x, y, want_grad=True, x_locs=x_locs)
        flat = dict(state.flat_params())
        return -lml, {path: -v for path, v in g.items() if flat[path].trainable}
    xd, yd = _xy(state, x, y)
    x_locs, _ = _select_locs(state, xd)
    if iterative:  # the pivots are constants here too; _sgpr._lml_iter under value_and_grad
        return _sgpr._loss_and_grad_iter_with_locs(state, xd, yd, x_locs, kwargs.get("num_evals", gpxargs.num_evals),
                                                   kwargs.get("num_lanczos", gpxargs.num_lanczos),
                                                   kwargs.get("key_lanczos", gpxargs.key_lanczos))
    return _sgpr._loss_and_grad_with_locs(state, xd, yd, x_locs)


def fit(state, x, y, iterative=False, **_ignored):
    """gpx/models/sgpr_rpchol.py:309-348 -> _sgpr_rpchol._fit_dense (36-67)."""
    if _is_comp(state):
        if iterative:
            raise NotImplementedError("composite kernels in SGPR_RPChol: the dense paths only")
        from ..kernels import _to_device

        x_locs, piv = _select_locs(state, _to_device(x))
        st = _composite_sparse.fit(state, x, y, x_locs=x_locs)
        return st.update(dict(x_locs=x_locs.cpu().numpy(), pivots=piv.cpu().numpy()))
    kernel, ell, sigma = _hyper(state)
    xd, yd = _xy(state, x, y)
    x_locs, piv = _select_locs(state, xd)
    if iterative:
        c, mu, _ = ops.sgpr_fit_iter(kernel.kind, xd, yd, x_locs, ell, sigma, mean_mode(state.mean_function))
    else:
        c, mu = ops.sgpr_fit(kernel.kind, xd, yd, x_locs, ell, sigma, mean_mode(state.mean_function))
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), c=c.cpu().numpy(),
                             mu=np.float64(mu.cpu().numpy()[0]), is_fitted=True, is_fitted_derivs=False,
                             x_locs=x_locs.cpu().numpy(), pivots=piv.cpu().numpy()))


def predict(state, x, full_covariance=False, iterative=False):
    if not state.is_fitted:
        raise RuntimeError("Model is not fitted. Run `fit` to fit the model before prediction.")
    if full_covariance and iterative:
        raise RuntimeError("'full_covariance=True' is not compatible with 'iterative=True'")
    from ..kernels import _to_device

    if _is_comp(state):
        return _composite_sparse.predict(state, x, full_covariance, x_locs=_to_device(state.x_locs))
    return _sgpr._predict_with_locs(state, _to_device(state.x_locs), x, full_covariance, iterative)


def init(kernel: Callable, n_locs, key=None, mean_function: Callable = data_mean, kernel_params=None, sigma=None,
         loss_fn: Callable = neg_log_marginal_likelihood) -> ModelState:
    """gpx/models/sgpr_rpchol.py:700-746."""
    if not callable(kernel):
        raise TypeError(f"kernel must be callable, you provided {type(kernel)}")
    if kernel_params is None:
        kernel_params = kernel.default_params()
    if sigma is None:
        sigma = _sgpr.default_params()["sigma"]
    if key is None:
        key = PRNGKey(int(time.time()))
    params = {"kernel_params": kernel_params, "sigma": sigma}
    opt = dict(loss_fn=loss_fn, is_fitted=False, is_fitted_derivs=False, c=None, mu=None, n_locs=n_locs, key=key,
               x_train=None, y_train=None, x_locs=None, pivots=None)
    return ModelState(kernel, mean_function, params, **opt)


class SGPR_RPChol(BaseGP):
    """gpx/models/sgpr_rpchol.py:756-811."""

    _init_default = dict(is_fitted=False, is_fitted_derivs=False, c=None, y_mean=None)

    def __init__(self, kernel, n_locs, key=None, mean_function=data_mean, kernel_params=None, sigma=None,
                 loss_fn=neg_log_marginal_likelihood) -> None:
        self.state = init(kernel=kernel, n_locs=n_locs, key=key, mean_function=mean_function,
                          kernel_params=kernel_params, sigma=sigma, loss_fn=loss_fn)

    _default_params_fun = staticmethod(_sgpr.default_params)
    _init_fun = staticmethod(init)
    _lml_fun = staticmethod(log_marginal_likelihood)
    _loss_and_grad_fun = staticmethod(loss_and_grad)
    _fit_fun = staticmethod(fit)
    _predict_fun = staticmethod(predict)
