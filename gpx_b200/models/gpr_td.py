"""GPR trained on targets AND derivatives (mirror of gpx/models/gpr_td.py:848-1170, the dense paths, with the
optional second derivative block).  Numerics run on the device through gpxb_td_gram / gpxb_gpr_td[2]_lml_grad /
gpxb_gpr_td[2]_fit (include/gpxb200.h): the block kernel matrix [[K + s_t I, d1kj], [d0kj, d01kj + s_d I]] is
assembled by the fused Gram, d0kj and Hessian kernels and factorised by the blocked FP64 Cholesky.

Second block (jacobian_2, y_derivs_2, sigma_derivs2; gpr_td.py:114-158): the device works on the two jacobians
concatenated along the variable axis, which is the reference's 3 x 3 block system with the derivative rows ordered
(sample, variable) instead of block after block -- a symmetric permutation, so the LML and its gradient are
unchanged; `c` is permuted to the reference's order [targets; block 1; block 2] before it is stored."""
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
from ..defaults import gpxargs, threefry_partitionable
from ..random import as_key, split
from ._iterative import slq_logdet
from .utils import randomized_minimization


def _hyper(state):
    kernel: Kernel = state.kernel
    if kernel.active_dims is not None:
        raise NotImplementedError("active_dims with derivative training is not supported")
    ell = kernel.lengthscale(state.params["kernel_params"])
    return (kernel, ell, float(np.asarray(state.params["sigma_targets"].value)),
            float(np.asarray(state.params["sigma_derivs"].value)))


def _cat_blocks(state, jacobian, y_derivs, jacobian_2, y_derivs_2):
    """(J, y_derivs) of both derivative blocks concatenated along the variable axis, the number of variables of the
    first block and sigma_derivs2 (None / nv / 0.0 without a second block)."""
    J = np.asarray(jacobian, dtype=np.float64)
    yd = None if y_derivs is None else np.asarray(y_derivs, dtype=np.float64).reshape(J.shape[0], J.shape[2])
    if jacobian_2 is None:
        if y_derivs_2 is not None:
            raise ValueError("y_derivs_2 needs jacobian_2")
        return J, yd, J.shape[2], 0.0
    if "sigma_derivs2" not in state.params:
        raise ValueError("a second derivative block needs the sigma_derivs2 Parameter (gpx/models/gpr_td.py:955, 68-69)")
    J2 = np.asarray(jacobian_2, dtype=np.float64)
    if J2.shape[:2] != J.shape[:2]:
        raise ValueError(f"jacobian_2 {J2.shape} does not match jacobian {J.shape} in (samples, features)")
    if yd is not None:
        if y_derivs_2 is None:
            raise ValueError("jacobian_2 needs y_derivs_2")
        yd = np.concatenate([yd, np.asarray(y_derivs_2, dtype=np.float64).reshape(J2.shape[0], J2.shape[2])], axis=1)
    return (np.concatenate([J, J2], axis=2), yd, J.shape[2], float(np.asarray(state.params["sigma_derivs2"].value)))


def _to_ref_order(c, ns, nv1, nv2):
    """[targets; (sample, variable of block 1 | block 2)] -> the reference's [targets; block 1; block 2]."""
    if nv2 == 0:
        return c
    d = c[ns:].reshape(ns, nv1 + nv2, -1)
    return np.concatenate([c[:ns], d[:, :nv1].reshape(ns * nv1, -1), d[:, nv1:].reshape(ns * nv2, -1)])


def _to_lib_order(c, ns, nv1, nv2):
    if nv2 == 0:
        return c
    a = c[ns:ns + ns * nv1].reshape(ns, nv1, -1)
    b = c[ns + ns * nv1:ns + ns * (nv1 + nv2)].reshape(ns, nv2, -1)
    return np.concatenate([c[:ns], np.concatenate([a, b], axis=1).reshape(ns * (nv1 + nv2), -1)])


def _ref_to_lib_perm(ns, nv1, nv2):
    """perm with v_lib = v_ref[perm]: rows [targets; block 1; block 2] -> [targets; (sample, [block 1 | block 2])]."""
    nv = nv1 + nv2
    u = np.arange(nv)[None, :]
    t = np.arange(ns)[:, None]
    d = np.where(u < nv1, ns + t * nv1 + u, ns + ns * nv1 + t * nv2 + (u - nv1))
    return np.concatenate([np.arange(ns), d.reshape(-1)])


def _fit_iter(state, x, y, jacobian, y_derivs, n_pivots, key_precond, jacobian_2=None, y_derivs_2=None, tol=1e-5,
              atol=1e-7):
    """_fit_iter (gpx/models/gpr_td.py:685-747): PCG on the matrix-free block operator with the rpcholesky_TD
    preconditioner.  -> (c (N + N nv, 1) device in the library's row order, mu device (1,), iterations)."""
    if n_pivots is None:
        raise ValueError("iterative fitting needs n_pivots (RPCholesky preconditioner size)")
    kernel, ell, st, sd = _hyper(state)
    J, yd, nv1, sd2 = _cat_blocks(state, jacobian, y_derivs, jacobian_2, y_derivs_2)
    key = as_key(key_precond if key_precond is not None else gpxargs.key_precond)
    return ops.gpr_td_fit_iter(kernel.kind, _to_device(x), _to_device(J), _to_device(y), _to_device(yd), ell, st, sd,
                               int(n_pivots), key, mean_mode(state.mean_function), threefry_partitionable(), tol=tol,
                               atol=atol, nv_first=nv1, sigma_derivs2=sd2)


def _lml_iter(state, x, y, jacobian, y_derivs, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond,
              jacobian_2=None, y_derivs_2=None, tol=1e-5, atol=1e-7):
    """_lml_iter (gpx/models/gpr_td.py:562-637): quadratic form from the PCG solution, log-determinant by stochastic
    Lanczos quadrature on the same operator (host eigh of the P tridiagonals as in _iterative.slq_logdet).  The
    Rademacher probes are drawn in the reference's row order and permuted to the library's, so that the estimate is
    the reference's number and not just an equally good one."""
    kernel, ell, st, sd = _hyper(state)
    J, yd, nv1, sd2 = _cat_blocks(state, jacobian, y_derivs, jacobian_2, y_derivs_2)
    ns, _, nv = J.shape
    m = ns + ns * nv
    c, mu, _ = _fit_iter(state, x, y, jacobian, y_derivs, n_pivots, key_precond, jacobian_2, y_derivs_2, tol, atol)
    mu_h = float(mu.cpu().numpy()[0])
    ym = np.concatenate([np.asarray(y, dtype=np.float64).reshape(-1, 1) - mu_h, yd.reshape(-1, 1)])  # library order
    part = threefry_partitionable()
    subkey, carried = split(as_key(key_lanczos), 2, part)
    U = ops.rademacher_probes(subkey, m, int(num_evals), part)
    if nv1 < nv:
        import torch

        U = U[torch.as_tensor(_ref_to_lib_perm(ns, nv1, nv - nv1), device=U.device)].contiguous()
    # breakdown restart (lanczos.py:92-108) only with one block: with two, the reference's normal deviates follow its
    # own row order
    alpha, beta, flag = ops.lanczos_td(kernel.kind, _to_device(x), _to_device(J), ell, st, sd, U, int(num_lanczos),
                                       nv_first=nv1, sigma_derivs2=sd2, restart_key=carried if nv1 == nv else None,
                                       restart_partitionable=part)
    if int(flag.cpu().numpy()[0]) != 0:
        raise RuntimeError("Lanczos breakdown (beta <= 1e-12): the Krylov space is exhausted; "
                           "use fewer Lanczos steps than rows")
    mll = -0.5 * float(np.sum(ym.T @ c.cpu().numpy()))
    mll -= 0.5 * slq_logdet(alpha.cpu().numpy(), beta.cpu().numpy(), m)
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def _lml_grad(state, x, y, jacobian, y_derivs, jacobian_2, y_derivs_2, want_grad):
    """[lml, d/d lengthscale, d/d sigma_targets, d/d sigma_derivs, d/d sigma_derivs2] (host floats)."""
    kernel, ell, st, sd = _hyper(state)
    J, yd, nv1, sd2 = _cat_blocks(state, jacobian, y_derivs, jacobian_2, y_derivs_2)
    if J.shape[2] == nv1:
        out = ops.gpr_td_lml_grad(kernel.kind, _to_device(x), _to_device(J), _to_device(y), _to_device(yd), ell, st, sd,
                                  mean_mode(state.mean_function), want_grad=want_grad)
        return [float(v) for v in out.cpu().numpy()] + [0.0]
    out = ops.gpr_td2_lml_grad(kernel.kind, _to_device(x), _to_device(J), _to_device(y), _to_device(yd), nv1, ell, st, sd,
                               sd2, mean_mode(state.mean_function), want_grad=want_grad)
    return [float(v) for v in out.cpu().numpy()]


def log_marginal_likelihood(state, x, y, jacobian, y_derivs, jacobian_2=None, y_derivs_2=None, iterative=False,
                            num_evals=gpxargs.num_evals, num_lanczos=gpxargs.num_lanczos,
                            key_lanczos=gpxargs.key_lanczos, n_pivots=None, key_precond=None, **_ignored):
    """gpx/models/gpr_td.py:848-911 -> _lml_dense (507-560) / _lml_iter (562-637)."""
    if iterative:
        return _lml_iter(state, x, y, jacobian, y_derivs, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond,
                         jacobian_2, y_derivs_2)
    return _lml_grad(state, x, y, jacobian, y_derivs, jacobian_2, y_derivs_2, want_grad=False)[0]


def neg_log_marginal_likelihood(state, x, y, jacobian, y_derivs, **kwargs):
    """gpx/models/gpr_td.py:914-944."""
    return -log_marginal_likelihood(state, x, y, jacobian, y_derivs, **kwargs)


def loss_and_grad(state, x, y, jacobian, y_derivs, jacobian_2=None, y_derivs_2=None, iterative=False, **_ignored):
    """Negative LML and its gradient w.r.t. the constrained (lengthscale, sigma_targets, sigma_derivs[, sigma_derivs2])
    from one fused device call: what jit(value_and_grad(loss)) returns inside scipy_minimize_ol."""
    if iterative:
        raise NotImplementedError("the gradient of the iterative GPR_TD LML is not implemented: train with "
                                  "iterative=False, or fit(..., iterative=True, minimize=False)")
    if state.loss_fn is not neg_log_marginal_likelihood:
        raise NotImplementedError("only neg_log_marginal_likelihood has a fused device gradient")
    lml, g_ell, g_st, g_sd, g_sd2 = _lml_grad(state, x, y, jacobian, y_derivs, jacobian_2, y_derivs_2, want_grad=True)
    grads = {}
    if jacobian_2 is not None and state.params["sigma_derivs2"].trainable:
        grads[("sigma_derivs2",)] = -g_sd2
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
    """gpx/models/gpr_td.py:947-1170 (dense paths; one or two derivative blocks)."""

    def __init__(self, kernel: Callable, mean_function: Callable = zero_mean, kernel_params=None, sigma_targets=None,
                 sigma_derivs=None, sigma_derivs2=None, loss_fn: Callable = neg_log_marginal_likelihood) -> None:
        if not callable(kernel):
            raise TypeError(f"kernel must be callable, you provided {type(kernel)}")
        if kernel_params is None:
            kernel_params = kernel.default_params()
        d = default_params()
        params = {"kernel_params": kernel_params,
                  "sigma_targets": sigma_targets if sigma_targets is not None else d["sigma_targets"],
                  "sigma_derivs": sigma_derivs if sigma_derivs is not None else d["sigma_derivs"]}
        if sigma_derivs2 is not None:  # the reference keeps a None entry otherwise (gpr_td.py:959-964)
            params["sigma_derivs2"] = sigma_derivs2
        opt = dict(x_train=None, jacobian_train=None, jacobian_train_2=None, jaccoef=None, jaccoef_2=None, y_train=None,
                   y_derivs_train=None, y_derivs_train_2=None, is_fitted=False, c=None, c_targets=None, mu=None,
                   loss_fn=loss_fn)
        self.state = ModelState(kernel, mean_function, params, **opt)

    def log_marginal_likelihood(self, x, y, jacobian, y_derivs, return_negative=False, **kwargs):
        """kwargs: jacobian_2, y_derivs_2 (second derivative block), iterative."""
        lml = log_marginal_likelihood(self.state, x, y, jacobian, y_derivs, **kwargs)
        return -lml if return_negative else lml

    def loss_and_grad(self, x, y, jacobian, y_derivs, **kwargs):
        return loss_and_grad(self.state, x, y, jacobian, y_derivs, **kwargs)

    def fit(self, x, y, jacobian, y_derivs, jacobian_2=None, y_derivs_2=None, iterative=False, key=None, num_restarts=0,
            minimize=True, return_history=False, loss_kwargs=None, n_pivots=None, key_precond=None, c_guess=None,
            opt_kwargs=None):
        """gpx/models/gpr_td.py:982-1108.  iterative=True: the coefficients come from the PCG solve (_fit_iter); the
        hyperparameters can only be trained on the dense LML (minimize=True needs iterative=False)."""
        if iterative:
            if minimize:
                raise NotImplementedError("the gradient of the iterative GPR_TD LML is not implemented: pass "
                                          "minimize=False (or train with iterative=False first)")
        if minimize:
            def loss_and_grad_fn(state, xx, yy):
                return loss_and_grad(state, xx, yy, jacobian, y_derivs, jacobian_2=jacobian_2, y_derivs_2=y_derivs_2)

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
        J, yd, nv1, sd2 = _cat_blocks(self.state, jacobian, y_derivs, jacobian_2, y_derivs_2)
        ns, _, nv = J.shape
        nv2 = nv - nv1
        if iterative:
            c, mu, self.cg_iterations_ = _fit_iter(self.state, x, y, jacobian, y_derivs, n_pivots, key_precond,
                                                   jacobian_2, y_derivs_2)
        elif nv2 == 0:
            c, mu = ops.gpr_td_fit(kernel.kind, _to_device(x), _to_device(J), _to_device(y), _to_device(yd), ell, st, sd,
                                   mean_mode(self.state.mean_function))
        else:
            c, mu = ops.gpr_td2_fit(kernel.kind, _to_device(x), _to_device(J), _to_device(y), _to_device(yd), nv1, ell,
                                    st, sd, sd2, mean_mode(self.state.mean_function))
        c_np = _to_ref_order(c.cpu().numpy(), ns, nv1, nv2)
        jaccoef = np.einsum("sv,sfv->sf", c_np[ns:ns + ns * nv1].reshape(ns, nv1), J[:, :, :nv1])
        jaccoef_2 = None
        if nv2:  # gpx/models/gpr_td.py:1076-1083
            jaccoef_2 = np.einsum("sv,sfv->sf", c_np[ns + ns * nv1:ns + ns * nv].reshape(ns, nv2), J[:, :, nv1:])
        self.state = self.state.update(dict(
            x_train=np.asarray(x), y_train=np.asarray(y), jacobian_train=np.asarray(jacobian),
            jacobian_train_2=None if jacobian_2 is None else np.asarray(jacobian_2), jaccoef=jaccoef,
            jaccoef_2=jaccoef_2, y_derivs_train=np.asarray(y_derivs),
            y_derivs_train_2=None if y_derivs_2 is None else np.asarray(y_derivs_2), c=c_np,
            c_targets=c_np[:ns].reshape(-1), mu=np.float64(mu.cpu().numpy()[0]), is_fitted=True))
        return self

    def predict(self, x, jacobian, jacobian_2=None, iterative=False):
        """gpx/models/gpr_td.py:1110-1160 -> _predict_dense (749-797): (targets (ns, 1), derivatives (ns, nv_1)
        [, derivatives of the second block (ns, nv_2)])."""
        if not self.state.is_fitted:
            raise RuntimeError("Model is not fitted. Run `fit` to fit the model before prediction.")
        # iterative=True (gpr_td.py:799-845) is ONE rectangular mat-vec K_nm c: the blocks below are that product,
        # evaluated in chunks of test samples so that the (test rows) x (train rows) block never exceeds ~1 GB
        kernel, ell, _, _ = _hyper(self.state)
        Jt2 = getattr(self.state, "jacobian_train_2", None)
        if (jacobian_2 is None) != (Jt2 is None):
            raise ValueError("jacobian_2 must be given exactly when the model was fitted with a second derivative block")
        Js, _, nv1, _ = _cat_blocks(self.state, jacobian, None, jacobian_2, None)
        Jt, _, nt1, _ = _cat_blocks(self.state, self.state.jacobian_train, None, Jt2, None)
        nt = Jt.shape[0]
        c_lib = _to_lib_order(np.asarray(self.state.c, dtype=np.float64).reshape(-1, 1), nt, nt1, Jt.shape[2] - nt1)
        ns, _, nv = Js.shape
        xs = np.asarray(x, dtype=np.float64)
        cols = nt * (1 + Jt.shape[2])
        step = max(1, min(ns, (1 << 27) // max(cols * (1 + nv), 1)))
        xt_d, Jt_d, c_d = _to_device(self.state.x_train), _to_device(Jt), _to_device(c_lib)
        pred = np.empty((ns + ns * nv, c_lib.shape[1]))
        for a in range(0, ns, step):
            b = min(ns, a + step)
            K_nm = ops.td_gram(kernel.kind, _to_device(xs[a:b]), _to_device(Js[a:b]), xt_d, Jt_d, ell)  # K_mn^T rows
            blk = ops.gemm(K_nm, c_d).cpu().numpy()
            pred[a:b] = blk[:b - a]
            pred[ns + a * nv:ns + b * nv] = blk[b - a:]
        pred[:ns] += float(self.state.mu)
        d = pred[ns:ns + ns * nv].reshape(ns, nv)
        if jacobian_2 is None:
            return pred[:ns], d
        return pred[:ns], d[:, :nv1].copy(), d[:, nv1:].copy()

    def print(self) -> None:
        print(self.state)

    def save(self, state_file):
        return self.state.save(state_file)

    def load(self, state_file):
        self.state = self.state.load(state_file)
        return self
