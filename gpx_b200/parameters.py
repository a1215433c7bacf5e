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
        raw = np.asarray(value)
        self.trainable = bool(trainable)
        self.bijector = bijector
        self.prior = prior
        # value and prior must agree in shape and dtype (gpx/parameters/parameter.py:45-47, utils.py:26-37): a Python
        # int is int64 under jax_enable_x64 and is rejected against the default float64 prior, as in the reference
        if tuple(raw.shape) != tuple(getattr(prior, "shape", ())):
            raise ValueError(f"value.shape={raw.shape} and prior.shape={getattr(prior, 'shape', ())} do not match.")
        if raw.dtype != np.dtype(getattr(prior, "dtype", np.float64)):
            raise ValueError(f"value.dtype={raw.dtype} and prior.dtype={np.dtype(prior.dtype)} do not match.")
        self.value = np.array(raw, dtype=np.float64)  # own copy; the numerical path is FP64 throughout

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
        """Shallow copy: kernel and Parameter objects are shared, the (nested) parameter dictionaries are new, so
        re-assigning an entry of one state leaves the other untouched (tests/gpx/parameters/test_model_state.py:91-110)."""
        def dicts(d):
            return {k: dicts(v) if isinstance(v, dict) else v for k, v in d.items()}

        return self.update(dict(params=dicts(self.params)))

    def flat_params(self):
        return list(_flatten_params(self.params))

    def with_param_value(self, path, value) -> "ModelState":
        p = dict(_flatten_params(self.params))[tuple(path)]
        return self.update(dict(params=_set_path(self.params, tuple(path), p.update(dict(value=value)))))

    def randomize(self, rng, opt=None) -> "ModelState":
        """New state with every trainable parameter re-drawn from its prior (gpx/parameters/model_state.py:349-404).
        `rng`: a numpy Generator, or a PRNG key of two uint32 words as the reference takes."""
        if not isinstance(rng, np.random.Generator):
            k = np.asarray(rng, dtype=np.uint64)
            rng = np.random.default_rng(int((k[0] << np.uint64(32)) | k[1]))
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
        # the reference raises "Wrong number of parameters in dumped file" when the structures differ
        # (model_state.py:321-327); here the keys are compared, so the message names what is wrong
        have = {"params:" + ":".join(path): pp for path, pp in _flatten_params(self.params)}
        in_file = {k for k in z.files if k.startswith("params:")}
        if in_file != set(have):
            raise ValueError("Wrong number of parameters in dumped file: missing "
                             f"{sorted(set(have) - in_file)}, unknown {sorted(in_file - set(have))}")
        for k in in_file:
            if tuple(np.shape(z[k])) != tuple(have[k].value.shape):
                raise ValueError(f"parameter {k}: the file holds shape {tuple(np.shape(z[k]))}, the model expects "
                                 f"{tuple(have[k].value.shape)}")
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
