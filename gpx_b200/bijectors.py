"""Bijectors constraining hyper-parameters (mirror of gpx/bijectors.py:27-71, 160-287).

Host-side scalar math only; ``grad_forward`` supplies the chain-rule factor the reference
obtains by differentiating ``forward`` (gpx/optimizers/utils.py:97-131)."""
from __future__ import annotations

import numpy as np


class Bijector:
    def forward(self, x):
        raise NotImplementedError

    def backward(self, x):
        raise NotImplementedError

    def grad_forward(self, x):
        raise NotImplementedError

    def __repr__(self):
        return f"{self.__class__.__name__}()"


class Identity(Bijector):
    def forward(self, x):
        return np.asarray(x, dtype=np.float64)

    def backward(self, x):
        return np.asarray(x, dtype=np.float64)

    def grad_forward(self, x):
        return np.ones_like(np.asarray(x, dtype=np.float64))


class Softplus(Bijector):
    """y = log(1 + exp(x)) + 1e-300 (gpx/bijectors.py:27-47)."""

    def forward(self, x):
        return np.logaddexp(np.asarray(x, dtype=np.float64), 0.0) + 1e-300

    def backward(self, x):
        x = np.asarray(x, dtype=np.float64)
        return x + np.log(-np.expm1(-x))

    def grad_forward(self, x):
        return 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))
