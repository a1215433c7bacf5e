"""Priors on hyper-parameters (mirror of the constructor surface of gpx/priors.py:190-300).
Only what the GP hot path touches: shape/dtype bookkeeping, sampling for random restarts
and logpdf for MAP losses are host-side NumPy."""
from __future__ import annotations

import numpy as np
from scipy import stats


class Prior:
    shape = ()
    dtype = np.float64

    def sample(self, rng):
        raise NotImplementedError

    def logpdf(self, x):
        raise NotImplementedError

    def dlogpdf(self, x):
        """d logpdf / d x (what autodiff of the MAP loss yields in the reference)."""
        raise NotImplementedError

    def __repr__(self):
        return f"{self.__class__.__name__}()"


class NormalPrior(Prior):
    def __init__(self, loc=0.0, scale=1.0, shape=(), dtype=np.float64):
        self.loc, self.scale, self.shape, self.dtype = loc, scale, tuple(shape), dtype

    def sample(self, rng):
        return self.loc + self.scale * rng.standard_normal(self.shape)

    def logpdf(self, x):
        return stats.norm.logpdf(x, self.loc, self.scale)

    def dlogpdf(self, x):
        return -(np.asarray(x, dtype=np.float64) - self.loc) / self.scale**2


class GammaPrior(Prior):
    def __init__(self, concentration=1.0, rate=1.0, shape=(), dtype=np.float64):
        self.concentration, self.rate, self.shape, self.dtype = concentration, rate, tuple(shape), dtype

    def sample(self, rng):
        return rng.gamma(self.concentration, 1.0 / self.rate, self.shape)

    def logpdf(self, x):
        return stats.gamma.logpdf(x, self.concentration, scale=1.0 / self.rate)

    def dlogpdf(self, x):
        return (self.concentration - 1.0) / np.asarray(x, dtype=np.float64) - self.rate
