"""Exact Gaussian-process regression facade (mirror of gpx/models/gpr.py:70-118, 251-273,
397-434, 542-593, 863-976).  Numerics run on the device through the C ABI:
gpxb_gpr_lml_grad / gpxb_gpr_fit / gpxb_gpr_predict (include/gpxb200.h)."""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np

from .. import ops
from ..bijectors import Softplus
from ..defaults import gpxargs
from ..kernels import Kernel, _to_device
from ..mean_functions import data_mean, mean_mode
from ..parameters import ModelState, Parameter
from ..priors import GammaPrior
from . import _composite
from .base import BaseGP


def _hyper(state):
    kernel: Kernel = state.kernel
    ell = kernel.lengthscale(state.params["kernel_params"])
    sigma = float(np.asarray(state.params["sigma"].value))
    return kernel, ell, sigma


def _xy(state, x, y):
    xd = _to_device(state.kernel._slice(x))
    yd = _to_device(y)
    if yd.dim() == 1:
        yd = yd.reshape(-1, 1)
    if xd.dim() != 2 or yd.shape[0] != xd.shape[0]:
        raise ValueError(f"x {tuple(xd.shape)} and y {tuple(yd.shape)} must have the same number of rows")
    return xd, yd


def log_marginal_likelihood(state, x, y, iterative=False, num_evals=gpxargs.num_evals,
                            num_lanczos=gpxargs.num_lanczos, key_lanczos=gpxargs.key_lanczos,
                            n_pivots=None, key_precond=None):
    """gpx/models/gpr.py:70-118 -> _lml_dense / _lml_iter (gpx/models/_gpr.py:737-821)."""
    if _composite.is_composite(state.kernel):
        if iterative:
            return _composite.lml_iter(state, x, y, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond)
        return _composite.lml_and_grad(state, x, y, want_grad=False)[0]
    kernel, ell, sigma = _hyper(state)
    xd, yd = _xy(state, x, y)
    if iterative:
        from . import _iterative

        return _iterative.lml_iter(state, xd, yd, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots,
                                   key_precond)
    out = ops.gpr_lml_grad(kernel.kind, xd, yd, ell, sigma, mean_mode(state.mean_function), want_grad=False)
    return float(out.cpu().numpy()[0])


def neg_log_marginal_likelihood(state, x, y, **kwargs):
    """gpx/models/gpr.py:251-273."""
    return -log_marginal_likelihood(state, x, y, **kwargs)


def log_prior(state):
    """log p(theta) assuming independent hyper-parameters: the sum of prior.logpdf(value) over EVERY parameter,
    trainable or not (gpx/models/gpr.py:177-184)."""
    return float(sum(np.sum(p.prior.logpdf(p.value)) for _, p in state.flat_params()))


def log_posterior(state, x, y, **kwargs):
    """gpx/models/gpr.py:187-215: log p(theta | y) = lml (normalised by n, as the reference's is) + log p(theta)."""
    return log_marginal_likelihood(state, x, y, **kwargs) + log_prior(state)


def neg_log_posterior(state, x, y, **kwargs):
    """gpx/models/gpr.py:303-325: the MAP loss (`loss_fn=neg_log_posterior`)."""
    return -log_posterior(state, x, y, **kwargs)


def _add_prior_gradient(state, loss, grads):
    """MAP loss from the LML loss: subtract log p(theta) and its derivative w.r.t. every trainable value."""
    flat = dict(state.flat_params())
    for path in list(grads):
        grads[path] = grads[path] - np.asarray(flat[path].prior.dlogpdf(flat[path].value), dtype=np.float64)
    return loss - log_prior(state), grads


def loss_and_grad(state, x, y, iterative=False, **kwargs):
    """Negative LML and its gradient w.r.t. the constrained hyper-parameters, from one fused
    device call (stands in for jit(value_and_grad(loss)), gpx/optimizers/scipy_optimize.py:72-78).
    Non-trainable parameters get no gradient entry (stop_gradient, model_state.py:93-103)."""
    if state.loss_fn is neg_log_posterior:  # MAP estimate: the LML part below, the prior part on the host
        loss, grads = loss_and_grad(state.update(dict(loss_fn=neg_log_marginal_likelihood)), x, y, iterative, **kwargs)
        return _add_prior_gradient(state, loss, grads)
    if state.loss_fn is not neg_log_marginal_likelihood:
        raise NotImplementedError("only neg_log_marginal_likelihood and neg_log_posterior have a device gradient")
    if _composite.is_composite(state.kernel):
        if iterative:
            raise NotImplementedError("the gradient of the iterative LML with a composite kernel is not implemented "
                                      "(train with iterative=False, or fit(..., iterative=True, minimize=False))")
        lml, g = _composite.lml_and_grad(state, x, y, want_grad=True)
        flat = dict(state.flat_params())
        return -lml, {path: -v for path, v in g.items() if flat[path].trainable}
    kernel, ell, sigma = _hyper(state)
    xd, yd = _xy(state, x, y)
    if iterative:
        from . import _iterative

        lml, g_ell, g_sigma = _iterative.lml_iter_grad(
            state, xd, yd, ell, sigma, kwargs.get("num_evals", gpxargs.num_evals),
            kwargs.get("num_lanczos", gpxargs.num_lanczos), kwargs.get("key_lanczos", gpxargs.key_lanczos),
            kwargs.get("n_pivots"), kwargs.get("key_precond"))
        return -lml, _grads_dict(state, g_ell, g_sigma)
    out = ops.gpr_lml_grad(kernel.kind, xd, yd, ell, sigma, mean_mode(state.mean_function), want_grad=True)
    lml, g_ell, g_sigma = (float(v) for v in out.cpu().numpy())
    return -lml, _grads_dict(state, g_ell, g_sigma)


def _grads_dict(state, g_ell, g_sigma):
    grads = {}
    if state.params["kernel_params"]["lengthscale"].trainable:
        grads[("kernel_params", "lengthscale")] = -g_ell
    if state.params["sigma"].trainable:
        grads[("sigma",)] = -g_sigma
    return grads


def fit(state, x, y, iterative=False, n_pivots=None, key_precond=None):
    """gpx/models/gpr.py:397-434 -> _fit_dense / _fit_iter (gpx/models/_gpr.py:262-310)."""
    if _composite.is_composite(state.kernel):
        if iterative:
            return _composite.fit_state_iter(state, x, y, n_pivots, key_precond)
        return _composite.fit(state, x, y)
    kernel, ell, sigma = _hyper(state)
    xd, yd = _xy(state, x, y)
    if iterative:
        from . import _iterative

        c, mu = _iterative.fit_iter(state, xd, yd, ell, sigma, n_pivots, key_precond)
    else:
        c, mu = ops.gpr_fit(kernel.kind, xd, yd, ell, sigma, mean_mode(state.mean_function))
        c, mu = c.cpu().numpy(), np.float64(mu.cpu().numpy()[0])
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), c=c, mu=mu, is_fitted=True,
                             is_fitted_derivs=False))


def predict(state, x, full_covariance=False, iterative=False):
    """gpx/models/gpr.py:542-593 -> _predict_dense (gpx/models/_gpr.py:434-463)."""
    if not state.is_fitted:
        raise RuntimeError("Model is not fitted. Run `fit` to fit the model before prediction.")
    if state.is_fitted_derivs:
        raise RuntimeError("Model is trained on derivatives. For the prediction,"
                           " run `predict_derivs` and `predict_y_derivs`")
    if full_covariance and iterative:
        raise RuntimeError("'full_covariance=True' is not compatible with 'iterative=True'")
    if _composite.is_composite(state.kernel):
        # iterative=True (_predict_iter, _gpr.py:466-486) is one rectangular mat-vec K_nm c: the same product
        return _composite.predict(state, x, full_covariance)
    kernel, ell, sigma = _hyper(state)
    import torch

    xt = _to_device(kernel._slice(state.x_train))
    xs = _to_device(kernel._slice(x))
    c = _to_device(state.c)
    if c.dim() == 1:
        c = c.reshape(-1, 1)
    mu = torch.full((1,), float(state.mu), dtype=torch.float64, device=xs.device)
    if iterative:
        # matrix-free mean: same value, no (n* x N) kernel block stored
        out = ops.kmatvec(kernel.kind, xs, xt, c, ell, diag_add=0.0) + mu
        return out.cpu().numpy()
    out = ops.gpr_predict(kernel.kind, xt, xs, c, mu, ell, sigma, full_covariance=full_covariance)
    if full_covariance:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out.cpu().numpy()


def _xyj(state, x, y, jacobian):
    if state.kernel.active_dims is not None:
        raise NotImplementedError("active_dims with derivative training is not supported")
    xd = _to_device(x)
    Jd = _to_device(jacobian)
    yd = None if y is None else _to_device(y)
    if xd.dim() != 2 or Jd.dim() != 3 or Jd.shape[0] != xd.shape[0] or Jd.shape[1] != xd.shape[1]:
        raise ValueError(f"jacobian {tuple(Jd.shape)} must be (n, nf, nv) for x {tuple(xd.shape)}")
    if yd is not None and yd.numel() != Jd.shape[0] * Jd.shape[2]:
        raise ValueError(f"y {tuple(yd.shape)} must hold n*nv = {Jd.shape[0] * Jd.shape[2]} derivative values")
    return xd, yd, Jd


def log_marginal_likelihood_derivs(state, x, y, jacobian, iterative=False, num_evals=gpxargs.num_evals,
                                   num_lanczos=gpxargs.num_lanczos, key_lanczos=gpxargs.key_lanczos, n_pivots=None,
                                   key_precond=None, **_ignored):
    """gpx/models/gpr.py:121-174 -> _lml_derivs_dense / _lml_derivs_iter (gpx/models/_gpr.py:824-935); zero mean."""
    kernel, ell, sigma = _hyper(state)
    xd, yd, Jd = _xyj(state, x, y, jacobian)
    if iterative:
        from . import _iterative

        return _iterative.lml_derivs_iter(kernel.kind, xd, Jd, yd, ell, sigma, num_evals, num_lanczos, key_lanczos,
                                          n_pivots, key_precond)
    out = ops.gpr_lml_derivs(kernel.kind, xd, Jd, yd, ell, sigma)
    return float(out.cpu().numpy()[0])


def neg_log_marginal_likelihood_derivs(state, x, y, jacobian, **kwargs):
    """gpx/models/gpr.py:276-300."""
    return -log_marginal_likelihood_derivs(state, x, y, jacobian, **kwargs)


def loss_and_grad_derivs(state, x, y, jacobian, iterative=False, **kwargs):
    """Negative derivative-LML and its gradient w.r.t. the constrained hyper-parameters from one fused
    device call -- stands in for jit(value_and_grad(loss)) with loss = neg_log_marginal_likelihood_derivs
    inside scipy_minimize_derivs (gpx/optimizers/scipy_optimize.py:72-78, 92-122)."""
    if state.loss_fn not in (neg_log_marginal_likelihood, neg_log_marginal_likelihood_derivs):
        raise NotImplementedError("only neg_log_marginal_likelihood_derivs has a fused device gradient")
    kernel, ell, sigma = _hyper(state)
    xd, yd, Jd = _xyj(state, x, y, jacobian)
    if iterative:
        from . import _iterative

        lml, g_ell, g_sigma = _iterative.lml_derivs_iter_grad(
            kernel.kind, xd, Jd, yd, ell, sigma, kwargs.get("num_evals", gpxargs.num_evals),
            kwargs.get("num_lanczos", gpxargs.num_lanczos), kwargs.get("key_lanczos", gpxargs.key_lanczos),
            kwargs.get("n_pivots"), kwargs.get("key_precond"))
        return -lml, _grads_dict(state, g_ell, g_sigma)
    out = ops.gpr_lml_grad_derivs(kernel.kind, xd, Jd, yd, ell, sigma, want_grad=True)
    lml, g_ell, g_sigma = (float(v) for v in out.cpu().numpy())
    return -lml, _grads_dict(state, g_ell, g_sigma)


def fit_derivs(state, x, y, jacobian, iterative=False, n_pivots=None, key_precond=None, **_ignored):
    """gpx/models/gpr.py:437-489 -> _fit_derivs_dense / _fit_derivs_iter (gpx/models/_gpr.py:313-390) + jaccoef."""
    kernel, ell, sigma = _hyper(state)
    xd, yd, Jd = _xyj(state, x, y, jacobian)
    if iterative:
        from . import _iterative

        c, jc = _iterative.fit_derivs_iter(kernel.kind, xd, Jd, yd, ell, sigma, n_pivots, key_precond)
    else:
        c, jc = ops.gpr_fit_derivs(kernel.kind, xd, Jd, yd, ell, sigma)
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), jacobian_train=np.asarray(jacobian),
                             jaccoef=jc.cpu().numpy(), c=c.cpu().numpy().reshape(-1, 1), mu=np.float64(0.0),
                             is_fitted=False, is_fitted_derivs=True))


def predict_derivs(state, x, jacobian, full_covariance=False, iterative=False):
    """gpx/models/gpr.py:596-660 -> _predict_derivs_dense fast path (gpx/models/_gpr.py:509-528):
    mu + sum_s d01kjc(x_train, x, jaccoef, jacobian)[s], reshaped (ns, nv)."""
    if not getattr(state, "is_fitted_derivs", False):
        raise RuntimeError("Model is not fitted on derivatives. Run `fit_derivs` first.")
    import torch

    if full_covariance and iterative:
        raise RuntimeError("'full_covariance=True' is not compatible with 'iterative=True'")
    if full_covariance:
        # the reference returns the unreshaped (ns*nv, 1) mean beside the covariance (_gpr.py:552-556)
        kernel, ell, sigma = _hyper(state)
        xs, _, Js = _xyj(state, x, None, jacobian)
        mean, cov = ops.gpr_predict_derivs_full(kernel.kind, _to_device(state.x_train), _to_device(state.jacobian_train),
                                                xs, Js, _to_device(state.c).reshape(-1), ell, sigma, float(state.mu),
                                                full_covariance=True)
        return mean.cpu().numpy().reshape(-1, 1), cov.cpu().numpy()

    kernel, ell, _ = _hyper(state)
    xt = _to_device(state.x_train)
    xs, _, Js = _xyj(state, x, None, jacobian)
    if iterative:
        # _predict_derivs_iter (gpx/models/_gpr.py:563-605): the cross operator applied to c, never formed
        ns, _, nv = Js.shape
        mean = ops.hess_matvec(kernel.kind, xs, Js, xt, _to_device(state.jacobian_train),
                               _to_device(state.c).reshape(-1), ell) + float(state.mu)
        return mean.reshape(ns, nv).cpu().numpy()
    jc = _to_device(state.jaccoef).unsqueeze(2).contiguous()  # (N, nf, 1): the contracted Jacobian
    K = ops.hess_gram(kernel.kind, xt, jc, xs, Js, ell)  # (N, ns*nv)
    ones = torch.ones((1, K.shape[0]), dtype=torch.float64, device=K.device)
    mean = ops.gemm(ones, K) + float(state.mu)
    ns, _, nv = Js.shape
    return mean.reshape(ns, nv).cpu().numpy()


def predict_y_derivs(state, x, full_covariance=False, iterative=False):
    """gpx/models/gpr.py:660-717 -> _predict_y_derivs_dense fast path (gpx/models/_gpr.py:630-638):
    mu + sum_s d0kjc(x_train, x, jaccoef)[s], shape (n, 1)."""
    if not getattr(state, "is_fitted_derivs", False):
        raise RuntimeError("Model is not fitted. Run `fit_derivs` to fit the model before prediction.")
    if full_covariance and iterative:
        raise RuntimeError("'full_covariance=True' is not compatible with 'iterative=True'")
    kernel, ell, sigma = _hyper(state)
    if kernel.active_dims is not None:
        raise NotImplementedError("active_dims with derivative training is not supported")
    if full_covariance:
        mean, cov = ops.gpr_predict_y_derivs_full(kernel.kind, _to_device(state.x_train),
                                                  _to_device(state.jacobian_train), _to_device(x),
                                                  _to_device(state.c).reshape(-1), ell, sigma, float(state.mu),
                                                  full_covariance=True)
        return mean.cpu().numpy().reshape(-1, 1), cov.cpu().numpy()
    out = ops.gpr_predict_y_derivs(kernel.kind, _to_device(state.x_train), _to_device(state.jaccoef), _to_device(x),
                                   ell, float(state.mu))
    return out.cpu().numpy().reshape(-1, 1)


def sample_prior(key, state, x, n_samples=1):
    """gpx/models/gpr.py:761-780: zero mean of the shape of x (so the key is split once per FEATURE column, as the
    reference does), covariance K(x, x) + (sigma^2 + 1e-10) I."""
    from .utils import sample

    if _composite.is_composite(state.kernel):
        xd = _to_device(x)
        cov = state.kernel.k_device(xd, xd, state.params["kernel_params"])
        cov.diagonal().add_(float(np.asarray(state.params["sigma"].value)) ** 2 + 1e-10)
    else:
        kernel, ell, sigma = _hyper(state)
        xd = _to_device(kernel._slice(x))
        cov = ops.gram(kernel.kind, xd, xd, ell, diag_add=sigma * sigma + 1e-10)
    return sample(key, np.zeros(np.asarray(x).shape), cov, n_samples)


def sample_posterior(key, state, x, n_samples=1):
    """gpx/models/gpr.py:809-832: predictive mean and covariance (+ 1e-10 I) of the fitted model."""
    from .utils import sample

    if not state.is_fitted:
        raise RuntimeError("Cannot sample from the posterior if the model is not fitted")
    mean, cov = predict(state, x, full_covariance=True)
    cov_dev = _to_device(cov)
    cov_dev.diagonal().add_(1e-10)
    return sample(key, mean, cov_dev, n_samples)


def _is_parameter_dict(d) -> bool:
    """Recursive type check of gpx/models/utils.py:42-51: composite kernels nest {"kernel1": ..., "kernel2": ...}."""
    return isinstance(d, dict) and all(_is_parameter_dict(v) if isinstance(v, dict) else isinstance(v, Parameter)
                                       for v in d.values())


def default_params() -> Dict[str, Parameter]:
    """gpx/models/gpr.py:863-870."""
    return dict(sigma=Parameter(value=1.0, trainable=True, bijector=Softplus(), prior=GammaPrior()))


def init(kernel: Callable, mean_function: Callable = data_mean, kernel_params=None, sigma=None,
         loss_fn: Callable = neg_log_marginal_likelihood) -> ModelState:
    """gpx/models/gpr.py:873-917."""
    if not callable(kernel):
        raise TypeError(f"kernel must be callable, you provided {type(kernel)}")
    if kernel_params is None:
        kernel_params = kernel.default_params()
    elif not _is_parameter_dict(kernel_params):
        raise TypeError("kernel_params must be a (nested, for Sum / Prod kernels) dict of Parameter")
    if sigma is None:
        sigma = default_params()["sigma"]
    elif not isinstance(sigma, Parameter):
        raise TypeError(f"sigma must be a Parameter, you provided {type(sigma)}")
    params = {"kernel_params": kernel_params, "sigma": sigma}
    opt = dict(x_train=None, y_train=None, loss_fn=loss_fn, is_fitted=False, is_fitted_derivs=False, c=None,
               mu=None, jacobian_train=None, jaccoef=None)
    return ModelState(kernel, mean_function, params, **opt)


class GPR(BaseGP):
    """gpx/models/gpr.py:925-976."""

    _init_default = dict(is_fitted=False, is_fitted_derivs=False, c=None, mu=None)

    def __init__(self, kernel, mean_function=data_mean, kernel_params=None, sigma=None,
                 loss_fn=neg_log_marginal_likelihood) -> None:
        self.state = init(kernel=kernel, mean_function=mean_function, kernel_params=kernel_params, sigma=sigma,
                          loss_fn=loss_fn)

    _default_params_fun = staticmethod(default_params)
    _init_fun = staticmethod(init)
    _lml_fun = staticmethod(log_marginal_likelihood)
    _loss_and_grad_fun = staticmethod(loss_and_grad)
    _fit_fun = staticmethod(fit)
    _predict_fun = staticmethod(predict)
    _lml_derivs_fun = staticmethod(log_marginal_likelihood_derivs)
    _loss_and_grad_derivs_fun = staticmethod(loss_and_grad_derivs)
    _fit_derivs_fun = staticmethod(fit_derivs)
    _predict_y_derivs_fun = staticmethod(predict_y_derivs)
    _predict_derivs_fun = staticmethod(predict_derivs)
    _sample_prior_fun = staticmethod(sample_prior)
    _sample_posterior_fun = staticmethod(sample_posterior)
