"""Parameter / ModelState (mirror of gpx/parameters/parameter.py:32-110 and
gpx/parameters/model_state.py:40-404, without the JAX pytree machinery)."""
from __future__ import annotations

import copy
from typing import Any, Callable, Dict

import numpy as np

from .bijectors import Bijector
from .priors import Prior


class Parameter:
    def __init__(self, value, trainable: bool, bijector: Bijector, prior: Prior) -> None:
        self.value = np.asarray(value, dtype=np.float64)
        self.trainable = bool(trainable)
        self.bijector = bijector
        self.prior = prior
        if tuple(self.value.shape) != tuple(getattr(prior, "shape", ())):
            raise ValueError(
                f"Shape of value ({self.value.shape}) and prior ({prior.shape}) do not match."
            )

    def __repr__(self) -> str:
        return (f"Parameter(value={self.value}, trainable={self.trainable}, "
                f"bijector={self.bijector}, prior={self.prior})")

    def update(self, update_dict: Dict) -> "Parameter":
        d = dict(value=self.value, trainable=self.trainable, bijector=self.bijector, prior=self.prior)
        d.update(update_dict)
        return Parameter(**d)

    def copy(self) -> "Parameter":
        return self.update({})

    def sample_prior(self, rng) -> "Parameter":
        return self.update(dict(value=self.bijector.forward(self.prior.sample(rng))))


def _flatten_params(params, prefix=()):
    for k, v in params.items():
        if isinstance(v, dict):
            yield from _flatten_params(v, prefix + (k,))
        else:
            yield prefix + (k,), v


def _set_path(params, path, value):
    out = dict(params)
    if len(path) == 1:
        out[path[0]] = value
    else:
        out[path[0]] = _set_path(params[path[0]], path[1:], value)
    return out


class ModelState:
    def __init__(self, kernel: Callable, mean_function: Callable, params: Dict[str, Any], **kwargs) -> None:
        self.kernel = kernel
        self.mean_function = mean_function
        self.params = params
        self._register = []
        for name, value in kwargs.items():
            setattr(self, name, value)
            self._register.append(name)

    def update(self, update_dict: Dict) -> "ModelState":
        d = {name: getattr(self, name) for name in self._register}
        base = dict(kernel=self.kernel, mean_function=self.mean_function, params=self.params)
        for k, v in update_dict.items():
            (base if k in base else d)[k] = v
        return ModelState(base["kernel"], base["mean_function"], base["params"], **d)

    def copy(self) -> "ModelState":
        return self.update({})

    def flat_params(self):
        return list(_flatten_params(self.params))

    def with_param_value(self, path, value) -> "ModelState":
        p = dict(_flatten_params(self.params))[tuple(path)]
        return self.update(dict(params=_set_path(self.params, tuple(path), p.update(dict(value=value)))))

    def randomize(self, rng, opt=None) -> "ModelState":
        params = self.params
        for path, p in _flatten_params(self.params):
            if p.trainable:
                params = _set_path(params, path, p.sample_prior(rng))
        upd = dict(params=params)
        if opt:
            upd.update(opt)
        return self.update(upd)

    def print_params(self) -> str:
        from tabulate import tabulate

        rows = [(":".join(path), p.trainable, p.bijector, p.prior, p.value.dtype, p.value.shape, p.value)
                for path, p in _flatten_params(self.params)]
        tab = tabulate(rows, headers=["name", "trainable", "bijector", "prior", "dtype", "shape", "value"])
        print(tab)
        return tab

    # plain .npz checkpoint with "params:kernel_params:lengthscale"-style keys
    # (gpx/parameters/model_state.py:246-347)
    def save(self, state_file: str) -> Dict:
        d = {"params:" + ":".join(path): p.value for path, p in _flatten_params(self.params)}
        for name in self._register:
            v = getattr(self, name)
            if v is not None and not callable(v) and not isinstance(v, dict):
                try:
                    d[name] = np.asarray(v)
                except Exception:
                    pass
        np.savez(state_file, **d)
        return d

    def load(self, state_file: str, allow_pickle: bool = False) -> "ModelState":
        """Loads a checkpoint written by `save` here or by the reference (gpx/parameters/model_state.py:246-347).
        The reference stores unset auxiliary entries (None) as pickled 0-d object arrays; those carry no data
        and are skipped unless allow_pickle=True is passed (never unpickle files you do not trust)."""
        z = np.load(state_file, allow_pickle=allow_pickle)
        st = self
        upd = {}
        for k in z.files:
            try:
                v = z[k]
            except ValueError:  # object array and allow_pickle=False: a None placeholder of the reference
                continue
            if k.startswith("params:"):
                st = st.with_param_value(k.split(":")[1:], np.asarray(v, dtype=np.float64))
            elif k in self._register:
                if v.dtype == object:
                    upd[k] = None if v.ndim == 0 and v.item() is None else v
                else:
                    upd[k] = bool(v) if v.dtype == bool and v.shape == () else v
        return st.update(upd)

    def __repr__(self):
        return f"ModelState(kernel={self.kernel}, params={self.params})"


def deepcopy_state(state: ModelState) -> ModelState:
    return copy.copy(state).update({})
