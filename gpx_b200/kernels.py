"""Stationary kernels (mirror of the class surface of gpx/kernels/kernels.py:1433-1488,
1623-1754): ``Kernel.k(x1, x2, params)`` and ``default_params()``.  Every evaluation runs
in the fused sm_100a Gram kernel (csrc/gram.cu) through the C ABI."""
from __future__ import annotations

from typing import Dict

import numpy as np

from .bijectors import Softplus
from .parameters import Parameter
from .priors import GammaPrior


def _to_device(a):
    import torch

    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


class Kernel:
    kind: str = ""

    def __init__(self, active_dims=None) -> None:
        self.active_dims = None if active_dims is None else np.asarray(active_dims, dtype=np.int64)

    def __call__(self, x1, x2, params: Dict[str, Parameter]):
        return self.k(x1, x2, params)

    def _slice(self, x):
        return x if self.active_dims is None else x[:, self.active_dims]

    def lengthscale(self, params) -> float:
        ell = np.asarray(params["lengthscale"].value)
        if ell.shape != ():
            raise NotImplementedError("gpx_b200 supports the reference's default scalar lengthscale only")
        return float(ell)

    def k(self, x1, x2, params, as_numpy: bool = True):
        """Kernel matrix k(x1, x2) (gpx/kernels/kernels.py:760-769 et al.)."""
        from . import ops

        same = x2 is x1
        d1 = _to_device(self._slice(x1))
        d2 = d1 if same else _to_device(self._slice(x2))
        out = ops.gram(self.kind, d1, d2, self.lengthscale(params))
        return out.cpu().numpy() if as_numpy else out

    def default_params(self) -> Dict[str, Parameter]:
        # gpx/kernels/kernels.py:1650-1658
        return dict(lengthscale=Parameter(value=1.0, trainable=True, bijector=Softplus(), prior=GammaPrior()))

    def __repr__(self):
        return f"{self.__class__.__name__}()"


class SquaredExponential(Kernel):
    """exp(-|x - x'|^2 / l^2)  (gpx/kernels/kernels.py:1623-1658; note: no 1/2)."""
    kind = "se"


class Matern12(Kernel):
    kind = "m12"


class Matern32(Kernel):
    kind = "m32"


class Matern52(Kernel):
    kind = "m52"
