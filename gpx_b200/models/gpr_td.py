"""GPR trained on targets AND derivatives (mirror of gpx/models/gpr_td.py:848-1170, the dense paths with one
derivative block).  Numerics run on the device through gpxb_td_gram / gpxb_gpr_td_lml_grad / gpxb_gpr_td_fit
(include/gpxb200.h): the block kernel matrix [[K + s_t I, d1kj], [d0kj, d01kj + s_d I]] is assembled by the
fused Gram, d0kj and Hessian kernels and factorised by the blocked FP64 Cholesky."""
from __future__ import annotations

import warnings
from typing import Callable, Dict

import numpy as np

from .. import ops
from ..bijectors import Softplus
from ..kernels import Kernel, _to_device
from ..mean_functions import mean_mode, zero_mean
from ..parameters import ModelState, Parameter
from ..priors import GammaPrior
from .utils import randomized_minimization


def _hyper(state):
    kernel: Kernel = state.kernel
    if kernel.active_dims is not None:
        raise NotImplementedError("active_dims with derivative training is not supported")
    ell = kernel.lengthscale(state.params["kernel_params"])
    return (kernel, ell, float(np.asarray(state.params["sigma_targets"].value)),
            float(np.asarray(state.params["sigma_derivs"].value)))


def _single_block(jacobian_2, y_derivs_2, iterative):
    if jacobian_2 is not None or y_derivs_2 is not None:
        raise NotImplementedError("the second derivative block (jacobian_2) of GPR_TD is not implemented")
    if iterative:
        raise NotImplementedError("the iterative GPR_TD paths (gpr_td.py:162-505, 562-747) are not implemented")


def log_marginal_likelihood(state, x, y, jacobian, y_derivs, jacobian_2=None, y_derivs_2=None, iterative=False,
                            **_ignored):
    """gpx/models/gpr_td.py:848-911 -> _lml_dense (507-560)."""
    _single_block(jacobian_2, y_derivs_2, iterative)
    kernel, ell, st, sd = _hyper(state)
    out = ops.gpr_td_lml_grad(kernel.kind, _to_device(x), _to_device(jacobian), _to_device(y), _to_device(y_derivs), ell,
                              st, sd, mean_mode(state.mean_function), want_grad=False)
    return float(out.cpu().numpy()[0])


def neg_log_marginal_likelihood(state, x, y, jacobian, y_derivs, **kwargs):
    """gpx/models/gpr_td.py:914-944."""
    return -log_marginal_likelihood(state, x, y, jacobian, y_derivs, **kwargs)


def loss_and_grad(state, x, y, jacobian, y_derivs, jacobian_2=None, y_derivs_2=None, iterative=False, **_ignored):
    """Negative LML and its gradient w.r.t. the constrained (lengthscale, sigma_targets, sigma_derivs) from one
    fused device call: what jit(value_and_grad(loss)) returns inside scipy_minimize_ol."""
    _single_block(jacobian_2, y_derivs_2, iterative)
    if state.loss_fn is not neg_log_marginal_likelihood:
        raise NotImplementedError("only neg_log_marginal_likelihood has a fused device gradient")
    kernel, ell, st, sd = _hyper(state)
    out = ops.gpr_td_lml_grad(kernel.kind, _to_device(x), _to_device(jacobian), _to_device(y), _to_device(y_derivs), ell,
                              st, sd, mean_mode(state.mean_function), want_grad=True)
    lml, g_ell, g_st, g_sd = (float(v) for v in out.cpu().numpy())
    grads = {}
    if state.params["kernel_params"]["lengthscale"].trainable:
        grads[("kernel_params", "lengthscale")] = -g_ell
    if state.params["sigma_targets"].trainable:
        grads[("sigma_targets",)] = -g_st
    if state.params["sigma_derivs"].trainable:
        grads[("sigma_derivs",)] = -g_sd
    return -lml, grads


def default_params() -> Dict[str, Parameter]:
    return dict(sigma_targets=Parameter(value=1.0, trainable=True, bijector=Softplus(), prior=GammaPrior()),
                sigma_derivs=Parameter(value=1.0, trainable=True, bijector=Softplus(), prior=GammaPrior()))


class GPR_TD:
    """gpx/models/gpr_td.py:947-1170 (dense paths, one derivative block)."""

    def __init__(self, kernel: Callable, mean_function: Callable = zero_mean, kernel_params=None, sigma_targets=None,
                 sigma_derivs=None, sigma_derivs2=None, loss_fn: Callable = neg_log_marginal_likelihood) -> None:
        if not callable(kernel):
            raise TypeError(f"kernel must be callable, you provided {type(kernel)}")
        if sigma_derivs2 is not None:
            raise NotImplementedError("the second derivative block (sigma_derivs2) of GPR_TD is not implemented")
        if kernel_params is None:
            kernel_params = kernel.default_params()
        d = default_params()
        params = {"kernel_params": kernel_params,
                  "sigma_targets": sigma_targets if sigma_targets is not None else d["sigma_targets"],
                  "sigma_derivs": sigma_derivs if sigma_derivs is not None else d["sigma_derivs"]}
        opt = dict(x_train=None, jacobian_train=None, jaccoef=None, y_train=None, y_derivs_train=None, is_fitted=False,
                   c=None, c_targets=None, mu=None, loss_fn=loss_fn)
        self.state = ModelState(kernel, mean_function, params, **opt)

    def log_marginal_likelihood(self, x, y, jacobian, y_derivs, return_negative=False, **kwargs):
        lml = log_marginal_likelihood(self.state, x, y, jacobian, y_derivs, **kwargs)
        return -lml if return_negative else lml

    def loss_and_grad(self, x, y, jacobian, y_derivs, **kwargs):
        return loss_and_grad(self.state, x, y, jacobian, y_derivs, **kwargs)

    def fit(self, x, y, jacobian, y_derivs, jacobian_2=None, y_derivs_2=None, iterative=False, key=None, num_restarts=0,
            minimize=True, return_history=False, loss_kwargs=None, n_pivots=None, key_precond=None, c_guess=None,
            opt_kwargs=None):
        """gpx/models/gpr_td.py:982-1108."""
        _single_block(jacobian_2, y_derivs_2, iterative)
        if minimize:
            def loss_and_grad_fn(state, xx, yy):
                return loss_and_grad(state, xx, yy, jacobian, y_derivs)

            self.state, optres, *history = randomized_minimization(
                key=key, state=self.state, x=x, y=y, loss_and_grad_fn=loss_and_grad_fn, num_restarts=num_restarts,
                return_history=return_history, opt_kwargs=opt_kwargs)
            self.optimize_results_ = optres
            if not optres.success:
                warnings.warn("optimization returned with error: {:d}. ({:s})".format(
                    optres.status, str(optres.message)), stacklevel=2)
            if return_history:
                self.states_history_, self.losses_history_ = history[0], history[1]
        kernel, ell, st, sd = _hyper(self.state)
        Jd = _to_device(jacobian)
        c, mu = ops.gpr_td_fit(kernel.kind, _to_device(x), Jd, _to_device(y), _to_device(y_derivs), ell, st, sd,
                               mean_mode(self.state.mean_function))
        ns, _, nv = Jd.shape
        c_np = c.cpu().numpy()
        jaccoef = np.einsum("sv,sfv->sf", c_np[ns:ns + ns * nv].reshape(ns, nv), np.asarray(jacobian, dtype=np.float64))
        self.state = self.state.update(dict(
            x_train=np.asarray(x), y_train=np.asarray(y), jacobian_train=np.asarray(jacobian), jaccoef=jaccoef,
            y_derivs_train=np.asarray(y_derivs), c=c_np, c_targets=c_np[:ns].reshape(-1),
            mu=np.float64(mu.cpu().numpy()[0]), is_fitted=True))
        return self

    def predict(self, x, jacobian, jacobian_2=None, iterative=False):
        """gpx/models/gpr_td.py:1110-1160 -> _predict_dense (749-797): (targets (ns, 1), derivatives (ns, nv))."""
        if not self.state.is_fitted:
            raise RuntimeError("Model is not fitted. Run `fit` to fit the model before prediction.")
        _single_block(jacobian_2, None, iterative)
        kernel, ell, _, _ = _hyper(self.state)
        Js = _to_device(jacobian)
        K_nm = ops.td_gram(kernel.kind, _to_device(x), Js, _to_device(self.state.x_train),
                           _to_device(self.state.jacobian_train), ell)   # (test rows) x (train rows) = K_mn^T
        pred = ops.gemm(K_nm, _to_device(self.state.c).reshape(-1, 1)).cpu().numpy()
        ns, _, nv = Js.shape
        pred[:ns] += float(self.state.mu)
        return pred[:ns], pred[ns:ns + ns * nv].reshape(ns, -1)

    def print(self) -> None:
        print(self.state)

    def save(self, state_file):
        return self.state.save(state_file)

    def load(self, state_file):
        self.state = self.state.load(state_file)
        return self
