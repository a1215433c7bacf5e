"""Sparse GPR, Projected Processes (mirror of gpx/models/sgpr.py:70-118, 309-338, 540-600,
676-804).  Numerics: gpxb_sgpr_lml / gpxb_sgpr_fit / gpxb_sgpr_predict (include/gpxb200.h)."""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np

from .. import ops
from ..bijectors import Identity, Softplus
from ..defaults import gpxargs
from ..kernels import _to_device
from ..mean_functions import data_mean, mean_mode
from ..parameters import ModelState, Parameter
from ..priors import NormalPrior
from . import _composite, _composite_sparse
from .base import BaseGP
from .gpr import _hyper, _xy


def _composite_guard(state, iterative=False, what="composite kernels in SGPR"):
    """True when the kernel is a Sum / Prod / Linear one (served by _composite_sparse: dense paths, fixed landmarks)."""
    if not _composite.is_composite(state.kernel):
        return False
    if iterative:
        raise NotImplementedError(f"{what}: the dense paths only")
    return True


def _locs(state):
    return _to_device(state.kernel._slice(np.asarray(state.params["x_locs"].value)))


def log_marginal_likelihood(state, x, y, iterative=False, num_evals=gpxargs.num_evals,
                            num_lanczos=gpxargs.num_lanczos, key_lanczos=gpxargs.key_lanczos, **_ignored):
    """gpx/models/sgpr.py:70-118 -> _sgpr._lml_dense (gpx/models/_sgpr.py:659-697)."""
    if _composite_guard(state, iterative):
        return _composite_sparse.lml_and_grad(state, x, y, want_grad=False)[0]
    kernel, ell, sigma = _hyper(state)
    xd, yd = _xy(state, x, y)
    if iterative:
        return _lml_iter_with_locs(state, xd, yd, _locs(state), num_evals, num_lanczos, key_lanczos)
    out = ops.sgpr_lml(kernel.kind, xd, yd, _locs(state), ell, sigma, mean_mode(state.mean_function))
    return float(out.cpu().numpy()[0])


def _lml_iter_with_locs(state, xd, yd, x_locs_dev, num_evals, num_lanczos, key_lanczos):
    """_sgpr._lml_iter (gpx/models/_sgpr.py:701-737): CG + SLQ on the Nystrom operator H + (sigma^2+1e-10) I."""
    from ..defaults import threefry_partitionable
    from ..random import as_key, split
    from ._iterative import slq_logdet

    kernel, ell, sigma = _hyper(state)
    N = yd.shape[0]
    subkey, carried = split(as_key(key_lanczos), 2, threefry_partitionable())
    U = ops.rademacher_probes(subkey, N, int(num_evals), threefry_partitionable())
    quad, _, alpha, beta, flag = ops.sgpr_lml_iter_parts(kernel.kind, xd, yd, x_locs_dev, ell, sigma, U,
                                                         int(num_lanczos), mean_mode(state.mean_function),
                                                         restart_key=carried,
                                                         restart_partitionable=threefry_partitionable())
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): use fewer Lanczos steps")
    logdet = slq_logdet(alpha.cpu().numpy(), beta.cpu().numpy(), N)
    return (-0.5 * float(quad.cpu().numpy()[0]) - 0.5 * logdet - N * 0.5 * np.log(2.0 * np.pi)) / N


def neg_log_marginal_likelihood(state, x, y, **kwargs):
    return -log_marginal_likelihood(state, x, y, **kwargs)


def _loss_and_grad_with_locs(state, xd, yd, x_locs_dev, locs_trainable=False):
    kernel, ell, sigma = _hyper(state)
    g_locs = None
    if locs_trainable:
        if kernel.active_dims is not None:
            raise NotImplementedError("trainable x_locs with active_dims is not supported")
        out, g_locs = ops.sgpr_lml_grad_locs(kernel.kind, xd, yd, x_locs_dev, ell, sigma,
                                             mean_mode(state.mean_function))
    else:
        out = ops.sgpr_lml_grad(kernel.kind, xd, yd, x_locs_dev, ell, sigma, mean_mode(state.mean_function))
    lml, g_ell, g_sigma = (float(v) for v in out.cpu().numpy())
    grads = {}
    if state.params["kernel_params"]["lengthscale"].trainable:
        grads[("kernel_params", "lengthscale")] = -g_ell
    if state.params["sigma"].trainable:
        grads[("sigma",)] = -g_sigma
    if g_locs is not None:
        grads[("x_locs",)] = -g_locs.cpu().numpy()
    return -lml, grads


def _loss_and_grad_iter_with_locs(state, xd, yd, x_locs_dev, num_evals, num_lanczos, key_lanczos, tol=1e-5):
    """value_and_grad of _sgpr._lml_iter (gpx/models/_sgpr.py:700-737) w.r.t. (lengthscale, sigma), landmarks fixed:
    d(r.c) = -c^T dA c (the CG solve is a custom_linear_solve), SLQ differentiated through the Lanczos recursion
    (gpxb_sgpr_lml_iter_grad_parts) and the host eigendecomposition (_iterative.slq_logdet_grad)."""
    from ..defaults import threefry_partitionable
    from ..random import as_key, split
    from ._iterative import slq_logdet_grad

    kernel, ell, sigma = _hyper(state)
    N = yd.shape[0]
    subkey, _ = split(as_key(key_lanczos), 2, threefry_partitionable())
    U = ops.rademacher_probes(subkey, N, int(num_evals), threefry_partitionable())
    quad, gq, _, alpha, beta, dalpha, dbeta, flag = ops.sgpr_lml_iter_grad_parts(
        kernel.kind, xd, yd, x_locs_dev, ell, sigma, U, int(num_lanczos), mean_mode(state.mean_function), tol=tol)
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): use fewer Lanczos steps")
    ld, dld = slq_logdet_grad(alpha.cpu().numpy(), beta.cpu().numpy(), dalpha.cpu().numpy(), dbeta.cpu().numpy(), N)
    q = float(quad.cpu().numpy()[0])
    g0, g1, cc = (float(v) for v in gq.cpu().numpy())
    lml = (-0.5 * q - 0.5 * ld - N * 0.5 * np.log(2.0 * np.pi)) / N
    g_ell = (0.5 * (2.0 * g0 - g1) - 0.5 * dld[0]) / N
    g_sigma = (sigma * cc - 0.5 * dld[1]) / N
    grads = {}
    if state.params["kernel_params"]["lengthscale"].trainable:
        grads[("kernel_params", "lengthscale")] = -g_ell
    if state.params["sigma"].trainable:
        grads[("sigma",)] = -g_sigma
    return -lml, grads


def loss_and_grad(state, x, y, iterative=False, **kwargs):
    """Negative SGPR LML and its gradient w.r.t. (lengthscale, sigma) and, when the `x_locs` Parameter is
    trainable, the landmark coordinates, from one fused device call (stands in for
    jit(value_and_grad(loss)))."""
    if iterative:
        if _composite_guard(state) or state.params["x_locs"].trainable:
            raise NotImplementedError("the gradient of the iterative SGPR LML covers fixed landmarks and plain kernels")
        xd, yd = _xy(state, x, y)
        return _loss_and_grad_iter_with_locs(state, xd, yd, _locs(state), kwargs.get("num_evals", gpxargs.num_evals),
                                             kwargs.get("num_lanczos", gpxargs.num_lanczos),
                                             kwargs.get("key_lanczos", gpxargs.key_lanczos))
    if _composite_guard(state):
        if state.params["x_locs"].trainable:
            raise NotImplementedError("trainable x_locs with a composite kernel is not supported")
        lml, g = _composite_sparse.lml_and_grad(state, x, y, want_grad=True)
        flat = dict(state.flat_params())
        return -lml, {path: -v for path, v in g.items() if flat[path].trainable}
    xd, yd = _xy(state, x, y)
    return _loss_and_grad_with_locs(state, xd, yd, _locs(state), bool(state.params["x_locs"].trainable))


def fit(state, x, y, iterative=False, **_ignored):
    """gpx/models/sgpr.py:309-338 -> _sgpr._fit_dense (gpx/models/_sgpr.py:319-339).  The extra
    keywords BaseGP.fit passes (n_pivots, key_precond) are accepted and ignored."""
    if _composite_guard(state, iterative):
        return _composite_sparse.fit(state, x, y)
    kernel, ell, sigma = _hyper(state)
    xd, yd = _xy(state, x, y)
    if iterative:  # _sgpr._fit_iter (gpx/models/_sgpr.py:343-363)
        c, mu, _ = ops.sgpr_fit_iter(kernel.kind, xd, yd, _locs(state), ell, sigma, mean_mode(state.mean_function))
    else:
        c, mu = ops.sgpr_fit(kernel.kind, xd, yd, _locs(state), ell, sigma, mean_mode(state.mean_function))
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), c=c.cpu().numpy(),
                             mu=np.float64(mu.cpu().numpy()[0]), is_fitted=True, is_fitted_derivs=False))


def _predict_with_locs(state, x_locs_dev, x, full_covariance, iterative=False):
    import torch

    kernel, ell, sigma = _hyper(state)
    xs = _to_device(kernel._slice(x))
    c = _to_device(state.c)
    if c.dim() == 1:
        c = c.reshape(-1, 1)
    if iterative:  # _sgpr._predict_iter (gpx/models/_sgpr.py:530-546): mu + K_nm c, matrix-free
        return (ops.kmatvec(kernel.kind, xs, x_locs_dev, c, ell) + float(state.mu)).cpu().numpy()
    mu = torch.full((1,), float(state.mu), dtype=torch.float64, device=xs.device)
    out = ops.sgpr_predict(kernel.kind, x_locs_dev, xs, c, mu, ell, sigma, full_covariance=full_covariance)
    if full_covariance:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out.cpu().numpy()


def predict(state, x, full_covariance=False, iterative=False):
    """gpx/models/sgpr.py:540-600 -> _sgpr._predict_dense (gpx/models/_sgpr.py:434-467)."""
    if not state.is_fitted:
        raise RuntimeError("Model is not fitted. Run `fit` to fit the model before prediction.")
    if full_covariance and iterative:
        raise RuntimeError("'full_covariance=True' is not compatible with 'iterative=True'")
    if _composite_guard(state, iterative):
        return _composite_sparse.predict(state, x, full_covariance)
    return _predict_with_locs(state, _locs(state), x, full_covariance, iterative)


def _locs_jac(state):
    if "jacobian_locs" not in state.params:
        raise RuntimeError("training on derivatives needs the `jacobian_locs` Parameter (gpx/models/sgpr.py:721-730)")
    if state.kernel.active_dims is not None:
        raise NotImplementedError("active_dims with derivative training is not supported")
    xl = _to_device(np.asarray(state.params["x_locs"].value))
    Jl = _to_device(np.asarray(state.params["jacobian_locs"].value))
    return xl, Jl


def fit_derivs(state, x, y, jacobian, iterative=False, **_ignored):
    """gpx/models/sgpr.py:340-390 -> _sgpr._fit_derivs_dense (gpx/models/_sgpr.py:366-396) with _A_derivs_lhs
    (:65-93): c = (sigma^2 K_mm + K_mn K_nm + 1e-10 I)^-1 K_mn y with K = d01kj, zero mean; the Hessian-kernel
    blocks, the SYRK, the Cholesky factorisation and the solves are device calls (the reference's LU solve is a
    Cholesky solve here)."""
    import torch

    kernel, ell, sigma = _hyper(state)
    xl, Jl = _locs_jac(state)
    xd, Jd = _to_device(x), _to_device(jacobian)
    if iterative:  # _fit_derivs_iter (gpx/models/_sgpr.py:399-428): CG on the matrix-free Hessian operators
        c, _ = ops.sgpr_fit_derivs_iter(kernel.kind, xd, Jd, _to_device(y).reshape(-1), xl, Jl, ell, sigma)
        ml, _, nvl = Jl.shape
        jaccoef = torch.einsum("sv,sfv->sf", c.reshape(ml, nvl), Jl)
        return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), jacobian_train=np.asarray(jacobian),
                                 jaccoef=jaccoef.cpu().numpy(), c=c.cpu().numpy(), mu=np.float64(0.0), is_fitted=False,
                                 is_fitted_derivs=True))
    yd = _to_device(y).reshape(1, -1)
    K_mn = ops.hess_gram(kernel.kind, xl, Jl, xd, Jd, ell)                      # (M nv_l, N nv)
    C = ops.hess_gram(kernel.kind, xl, Jl, xl, Jl, ell)                         # K_mm (full)
    C = ops.gemm(K_mn, K_mn, transb=True, alpha=1.0, beta=sigma**2, C=C)        # sigma^2 K_mm + K_mn K_nm
    C.diagonal().add_(1e-10)
    rhs = ops.gemm(yd, K_mn, transb=True)                                       # (K_mn y)^T, 1 x M'
    L, inv = ops.potrf(C)
    ops.trsm(L, inv, rhs, trans=True)
    ops.trsm(L, inv, rhs, trans=False)
    c = rhs.reshape(-1, 1)
    ml, _, nvl = Jl.shape
    jaccoef = torch.einsum("sv,sfv->sf", c.reshape(ml, nvl), Jl)
    return state.update(dict(x_train=np.asarray(x), y_train=np.asarray(y), jacobian_train=np.asarray(jacobian),
                             jaccoef=jaccoef.cpu().numpy(), c=c.cpu().numpy(), mu=np.float64(0.0), is_fitted=False,
                             is_fitted_derivs=True))


def predict_derivs(state, x, jacobian, full_covariance=False, iterative=False):
    """gpx/models/sgpr.py -> _sgpr._predict_derivs_dense fast path (gpx/models/_sgpr.py:489-496):
    mu + sum_s d01kjc(x_locs, x, jaccoef, jacobian)[s], reshaped (ns, nv)."""
    if not getattr(state, "is_fitted_derivs", False):
        raise RuntimeError("Model is not fitted on derivatives. Run `fit_derivs` first.")
    if full_covariance or iterative:
        raise NotImplementedError("only the contracted-jacobian mean of the SGPR derivative model is implemented")
    import torch

    kernel, ell, _ = _hyper(state)
    xl, _ = _locs_jac(state)
    jc = _to_device(state.jaccoef).unsqueeze(2).contiguous()
    xs, Js = _to_device(x), _to_device(jacobian)
    K = ops.hess_gram(kernel.kind, xl, jc, xs, Js, ell)
    ones = torch.ones((1, K.shape[0]), dtype=torch.float64, device=K.device)
    mean = ops.gemm(ones, K) + float(state.mu)
    ns, _, nv = Js.shape
    return mean.reshape(ns, nv).cpu().numpy()


def predict_y_derivs(state, x, full_covariance=False, iterative=False):
    """_sgpr._predict_y_derivs_dense fast path (gpx/models/_sgpr.py:581-625): mu + sum_s d0kjc(x_locs, x, jaccoef)."""
    if not getattr(state, "is_fitted_derivs", False):
        raise RuntimeError("Model is not fitted. Run `fit_derivs` to fit the model before prediction.")
    if full_covariance:
        raise NotImplementedError("only the contracted-jacobian mean of the SGPR derivative model is implemented")
    kernel, ell, _ = _hyper(state)
    xl, _ = _locs_jac(state)
    out = ops.gpr_predict_y_derivs(kernel.kind, xl, _to_device(state.jaccoef), _to_device(x), ell, float(state.mu))
    return out.cpu().numpy().reshape(-1, 1)


def default_params() -> Dict[str, Parameter]:
    """gpx/models/sgpr.py:676-684."""
    return dict(sigma=Parameter(value=1.0, trainable=True, bijector=Softplus(), prior=NormalPrior(loc=0.0, scale=1.0)))


def init(kernel: Callable, mean_function: Callable = data_mean, x_locs: Parameter = None, jacobian_locs=None,
         kernel_params=None, sigma=None, loss_fn: Callable = neg_log_marginal_likelihood) -> ModelState:
    """gpx/models/sgpr.py:687-747."""
    if not callable(kernel):
        raise TypeError(f"kernel must be callable, you provided {type(kernel)}")
    if not isinstance(x_locs, Parameter):
        raise TypeError(f"x_locs must be a Parameter, you provided {type(x_locs)}")
    if kernel_params is None:
        kernel_params = kernel.default_params()
    if sigma is None:
        sigma = default_params()["sigma"]
    params = {"kernel_params": kernel_params, "sigma": sigma, "x_locs": x_locs}
    if jacobian_locs is not None:  # gpx/models/sgpr.py:721-730
        if not isinstance(jacobian_locs, Parameter):
            raise TypeError(f"jacobian_locs must be a Parameter, you provided {type(jacobian_locs)}")
        params["jacobian_locs"] = jacobian_locs
    opt = dict(x_train=None, y_train=None, loss_fn=loss_fn, is_fitted=False, is_fitted_derivs=False, c=None, mu=None)
    return ModelState(kernel, mean_function, params, **opt)


def locs_parameter(x_locs, trainable=False) -> Parameter:
    """Convenience: wrap landmark coordinates as the Parameter the constructor expects."""
    x_locs = np.asarray(x_locs, dtype=np.float64)
    return Parameter(x_locs, trainable, Identity(), NormalPrior(shape=x_locs.shape))


class SGPR(BaseGP):
    """gpx/models/sgpr.py:750-804."""

    _init_default = dict(is_fitted=False, is_fitted_derivs=False, c=None, mu=None)

    def __init__(self, kernel, mean_function=data_mean, x_locs=None, jacobian_locs=None, kernel_params=None,
                 sigma=None, loss_fn=neg_log_marginal_likelihood) -> None:
        self.state = init(kernel=kernel, mean_function=mean_function, x_locs=x_locs, jacobian_locs=jacobian_locs,
                          kernel_params=kernel_params, sigma=sigma, loss_fn=loss_fn)

    _default_params_fun = staticmethod(default_params)
    _init_fun = staticmethod(init)
    _lml_fun = staticmethod(log_marginal_likelihood)
    _loss_and_grad_fun = staticmethod(loss_and_grad)
    _fit_fun = staticmethod(fit)
    _predict_fun = staticmethod(predict)
    _fit_derivs_fun = staticmethod(fit_derivs)
    _predict_derivs_fun = staticmethod(predict_derivs)
    _predict_y_derivs_fun = staticmethod(predict_y_derivs)
