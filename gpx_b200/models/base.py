"""BaseGP facade (mirror of gpx/models/base.py:43-709 for the GP hot path: fit / predict /
log_marginal_likelihood, state accessors, save/load, randomize)."""
from __future__ import annotations

import warnings

from ..defaults import gpxargs
from .utils import randomized_minimization


class BaseGP:
    def __repr__(self):
        return repr(self.print())

    @property
    def c_(self):
        return self.state.c

    @property
    def mu_(self):
        return self.state.mu

    @property
    def x_train(self):
        return self.state.x_train

    @property
    def y_train(self):
        return self.state.y_train

    @classmethod
    def from_state(cls, state):
        self = cls.__new__(cls)
        self.state = state
        return self

    def save(self, state_file):
        return self.state.save(state_file)

    def load(self, state_file):
        self.state = self.state.load(state_file)
        return self

    def init(self, *args, **kwargs):
        return self._init_fun(*args, **kwargs)

    def default_params(self):
        return self._default_params_fun()

    def randomize(self, key, reset=True):
        from .utils import _rng_from_key

        new_state = self.state.randomize(_rng_from_key(key), opt=self._init_default if reset else None)
        return self.from_state(new_state)

    def print(self):
        return self.state.print_params()

    def is_fitted(self):
        return hasattr(self.state, "c")  # gpx/models/base.py:375-387 (sic)

    def _check_is_fitted(self):
        if not self.is_fitted():
            raise RuntimeError(
                f"{self.__class__.__name__} is not fitted yet. Call 'fit' before using this model for prediction."
            )

    def sample(self, key, x, n_samples=1, kind="prior"):
        """Draws samples from the GP prior or posterior (gpx/models/base.py:99-128)."""
        if kind not in ("prior", "posterior"):
            raise ValueError(f"kind can be either 'prior' or 'posterior', you provided {kind}")
        fun = getattr(self, "_sample_prior_fun" if kind == "prior" else "_sample_posterior_fun", None)
        if fun is None:
            raise NotImplementedError(f"{self.__class__.__name__}.sample(kind={kind!r}) is not implemented")
        return fun(key, state=self.state, x=x, n_samples=n_samples)

    # ---- hot path -------------------------------------------------------------------
    def predict(self, x, full_covariance=False, iterative=False):
        """gpx/models/base.py:167-187."""
        self._check_is_fitted()
        return self._predict_fun(self.state, x=x, full_covariance=full_covariance, iterative=iterative)

    def log_marginal_likelihood(self, x, y, return_negative=False, iterative=False,
                                num_evals=gpxargs.num_evals, num_lanczos=gpxargs.num_lanczos,
                                key_lanczos=gpxargs.key_lanczos, n_pivots=None, key_precond=None):
        """gpx/models/base.py:266-302.  n_pivots / key_precond are accepted here (the reference's
        class method forgets to forward them)."""
        lml = self._lml_fun(self.state, x=x, y=y, iterative=iterative, num_evals=num_evals,
                            num_lanczos=num_lanczos, key_lanczos=key_lanczos, n_pivots=n_pivots,
                            key_precond=key_precond if key_precond is not None else gpxargs.key_precond)
        return -lml if return_negative else lml

    def log_marginal_likelihood_derivs(self, x, y, jacobian, return_negative=False, iterative=False, **kwargs):
        """gpx/models/base.py:304-345."""
        lml = self._lml_derivs_fun(self.state, x=x, y=y, jacobian=jacobian, iterative=iterative, **kwargs)
        return -lml if return_negative else lml

    def fit_derivs(self, x, y, jacobian, minimize=True, num_restarts=0, key=None, return_history=False,
                   iterative=False, loss_kwargs=None, opt_kwargs=None, n_pivots=None):
        """gpx/models/base.py:487-589: L-BFGS-B on the derivative LML (scipy_minimize_derivs binds the
        jacobian into the loss, gpx/optimizers/scipy_optimize.py:92-122), then the linear fit."""
        loss_kwargs = {} if loss_kwargs is None else dict(loss_kwargs)
        loss_kwargs["iterative"] = iterative
        loss_kwargs["n_pivots"] = n_pivots
        loss_kwargs["key_precond"] = gpxargs.key_precond
        if minimize:
            if not hasattr(self, "_loss_and_grad_derivs_fun"):
                raise NotImplementedError(f"{self.__class__.__name__} has no derivative-LML gradient")

            def loss_and_grad_fn(state, xx, yy):
                return self._loss_and_grad_derivs_fun(state, xx, yy, jacobian, **loss_kwargs)

            self.state, optres, *history = randomized_minimization(
                key=key, state=self.state, x=x, y=y, loss_and_grad_fn=loss_and_grad_fn,
                num_restarts=num_restarts, return_history=return_history, opt_kwargs=opt_kwargs)
            self.optimize_results_ = optres
            if not optres.success:
                warnings.warn("optimization returned with error: {:d}. ({:s})".format(
                    optres.status, str(optres.message)), stacklevel=2)
            if return_history:
                self.states_history_, self.losses_history_ = history[0], history[1]
        self.state = self._fit_derivs_fun(self.state, x=x, y=y, jacobian=jacobian, iterative=iterative,
                                          n_pivots=n_pivots, key_precond=gpxargs.key_precond)
        return self

    def predict_derivs(self, x, jacobian, full_covariance=False, iterative=False):
        """gpx/models/base.py:189-225."""
        return self._predict_derivs_fun(self.state, x=x, jacobian=jacobian, full_covariance=full_covariance,
                                        iterative=iterative)

    def predict_y_derivs(self, x, full_covariance=False, iterative=False):
        """gpx/models/base.py:217-240."""
        if not hasattr(self, "_predict_y_derivs_fun"):
            raise NotImplementedError(f"{self.__class__.__name__} has no predict_y_derivs")
        return self._predict_y_derivs_fun(self.state, x=x, full_covariance=full_covariance, iterative=iterative)

    @property
    def jacobian_train(self):
        return self.state.jacobian_train

    def loss_and_grad(self, x, y, **loss_kwargs):
        """(loss, {param path: d loss/d value}) -- what jit(value_and_grad(loss)) returns in the
        reference's optimiser loop (gpx/optimizers/scipy_optimize.py:72-78)."""
        return self._loss_and_grad_fun(self.state, x, y, **loss_kwargs)

    def fit(self, x, y, minimize=True, num_restarts=0, key=None, return_history=False, iterative=False,
            loss_kwargs=None, opt_kwargs=None, n_pivots=None):
        """gpx/models/base.py:389-485."""
        loss_kwargs = {} if loss_kwargs is None else dict(loss_kwargs)
        loss_kwargs["iterative"] = iterative
        loss_kwargs["n_pivots"] = n_pivots
        loss_kwargs["key_precond"] = gpxargs.key_precond

        def loss_and_grad_fn(state, xx, yy):
            return self._loss_and_grad_fun(state, xx, yy, **loss_kwargs)

        if minimize:
            self.state, optres, *history = randomized_minimization(
                key=key, state=self.state, x=x, y=y, loss_and_grad_fn=loss_and_grad_fn,
                num_restarts=num_restarts, return_history=return_history, opt_kwargs=opt_kwargs)
            self.optimize_results_ = optres
            if not optres.success:
                warnings.warn("optimization returned with error: {:d}. ({:s})".format(
                    optres.status, str(optres.message)), stacklevel=2)
            if return_history:
                self.states_history_, self.losses_history_ = history[0], history[1]
        self.state = self._fit_fun(self.state, x=x, y=y, iterative=iterative, n_pivots=n_pivots,
                                   key_precond=gpxargs.key_precond)
        return self
