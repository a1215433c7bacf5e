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
        """Kernel matrix k(x1, x2) (gpx/kernels/kernels.py:760-769 et al.).  A length-D lengthscale (one per active
        feature) is accepted HERE ONLY, as in the reference, where it broadcasts in x / lengthscale for `.k` while the
        derivative kernels and the models assume a scalar: the inputs are pre-scaled."""
        from . import ops

        same = x2 is x1
        d1 = _to_device(self._slice(x1))
        d2 = d1 if same else _to_device(self._slice(x2))
        ell = np.asarray(params["lengthscale"].value, dtype=np.float64)
        if ell.shape == ():
            out = ops.gram(self.kind, d1, d2, float(ell))
        else:
            if ell.shape != (d1.shape[1],):
                raise ValueError(f"lengthscale of shape {ell.shape} does not match {d1.shape[1]} active features")
            inv = _to_device(1.0 / ell)
            s1 = (d1 * inv).contiguous()
            out = ops.gram(self.kind, s1, s1 if same else (d2 * inv).contiguous(), 1.0)
        return out.cpu().numpy() if as_numpy else out

    def k_device(self, x1d, x2d, params):
        """Device kernel matrix from raw (unsliced) device inputs; `x2d is x1d` marks the symmetric case."""
        from . import ops

        a = x1d if self.active_dims is None else x1d[:, self.active_dims].contiguous()
        b = a if x2d is x1d else (x2d if self.active_dims is None else x2d[:, self.active_dims].contiguous())
        return ops.gram(self.kind, a, b, self.lengthscale(params))

    def default_params(self) -> Dict[str, Parameter]:
        # gpx/kernels/kernels.py:1650-1658
        return dict(lengthscale=Parameter(value=1.0, trainable=True, bijector=Softplus(), prior=GammaPrior()))

    # ---- derivative kernels (gpx/kernels/kernels.py:1433-1488; shapes as the kernelizers return them,
    # gpx/kernels/kernelizers.py:106-392).  All of them are blocks of the two device kernels gpxb_hess_gram
    # (d01kj) and gpxb_td_gram ([[k, d1kj], [d0kj, d01kj]]); the jacobian-free forms use identity jacobians.
    # SquaredExponential and Matern52 only (the kinds csrc/hess.cu implements).  With `active_dims` the jacobian
    # is restricted to the active feature rows (the kernel does not depend on the others).
    @staticmethod
    def _identity_jacobian(x):
        """(n, nf, nf) identity jacobians over ALL features: the jacobian-free forms d0k / d1k / d01k are the *kj forms
        with these; under active_dims the inactive feature ROWS are dropped by _deriv_operands, which leaves exact
        zeros for the inactive VARIABLES in the output (tests/gpx/kernels/test_kernels.py:551-580)."""
        x = np.asarray(x)
        return np.tile(np.eye(x.shape[1])[None], (x.shape[0], 1, 1))

    def _deriv_operands(self, x, jacobian):
        if self.kind not in ("se", "m52"):
            raise NotImplementedError(f"derivative kernels of {self.__class__.__name__} are not implemented "
                                      "(SquaredExponential and Matern52 only)")
        xd = _to_device(self._slice(np.asarray(x, dtype=np.float64)))
        J = np.asarray(jacobian, dtype=np.float64)
        if self.active_dims is not None:
            J = J[:, self.active_dims, :]
        return xd, _to_device(J)

    @staticmethod
    def _zero_jacobian(xd, nv=3):
        import torch

        return torch.zeros((xd.shape[0], xd.shape[1], nv), dtype=torch.float64, device=xd.device)

    def d01kj(self, x1, x2, params, jacobian1, jacobian2, as_numpy: bool = True):
        """Hessian kernel contracted with both jacobians: (n1 nv1, n2 nv2), row index sample * nv + variable."""
        from . import ops

        a, Ja = self._deriv_operands(x1, jacobian1)
        b, Jb = self._deriv_operands(x2, jacobian2)
        out = ops.hess_gram(self.kind, a, Ja, b, Jb, self.lengthscale(params))
        return out.cpu().numpy() if as_numpy else out

    def d01kjc(self, x1, x2, params, jaccoef, jacobian, as_numpy: bool = True):
        """d01kj with the first jacobian already contracted with the coefficients: (n1, n2 nv)."""
        jc = np.asarray(jaccoef, dtype=np.float64)[:, :, None]
        return self.d01kj(x1, x2, params, jc, jacobian, as_numpy)

    def d0kj(self, x1, x2, params, jacobian, as_numpy: bool = True):
        """Derivative w.r.t. the first argument contracted with its jacobian: (n1 nv, n2)."""
        from . import ops

        a, Ja = self._deriv_operands(x1, jacobian)
        b = _to_device(self._slice(np.asarray(x2, dtype=np.float64)))
        full = ops.td_gram(self.kind, a, Ja, b, self._zero_jacobian(b), self.lengthscale(params))
        out = full[a.shape[0]:, : b.shape[0]].contiguous()
        return out.cpu().numpy() if as_numpy else out

    def d1kj(self, x1, x2, params, jacobian, as_numpy: bool = True):
        """Derivative w.r.t. the second argument contracted with its jacobian: (n1, n2 nv)."""
        from . import ops

        a = _to_device(self._slice(np.asarray(x1, dtype=np.float64)))
        b, Jb = self._deriv_operands(x2, jacobian)
        full = ops.td_gram(self.kind, a, self._zero_jacobian(a), b, Jb, self.lengthscale(params))
        out = full[: a.shape[0], b.shape[0]:].contiguous()
        return out.cpu().numpy() if as_numpy else out

    def d0kjc(self, x1, x2, params, jaccoef, as_numpy: bool = True):
        """d0kj with the jacobian already contracted with the coefficients: (n1, n2)."""
        return self.d0kj(x1, x2, params, np.asarray(jaccoef, dtype=np.float64)[:, :, None], as_numpy)

    def d0k(self, x1, x2, params, as_numpy: bool = True):
        """(n1 nf, n2): d k / d x1."""
        return self.d0kj(x1, x2, params, self._identity_jacobian(x1), as_numpy)

    def d1k(self, x1, x2, params, as_numpy: bool = True):
        """(n1, n2 nf): d k / d x2."""
        return self.d1kj(x1, x2, params, self._identity_jacobian(x2), as_numpy)

    def d01k(self, x1, x2, params, as_numpy: bool = True):
        """(n1 nf, n2 nf): d^2 k / d x1 d x2."""
        return self.d01kj(x1, x2, params, self._identity_jacobian(x1), self._identity_jacobian(x2), as_numpy)

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


class Linear(Kernel):
    """x1 . x2 on the active dimensions, no parameters (gpx/kernels/kernels.py:226-236, class at :1535-1576).
    One DMMA GEMM (gpxb_gemm); available as a component of Sum / Prod kernels and in the dense exact-GPR paths."""
    kind = "linear"

    def lengthscale(self, params):
        raise NotImplementedError("the Linear kernel has no lengthscale")

    def default_params(self) -> Dict[str, Parameter]:
        return {}

    def k_device(self, x1d, x2d, params):
        from . import ops

        a = x1d if self.active_dims is None else x1d[:, self.active_dims].contiguous()
        b = a if x2d is x1d else (x2d if self.active_dims is None else x2d[:, self.active_dims].contiguous())
        return ops.gemm(a, b, transb=True)

    def k(self, x1, x2, params, as_numpy: bool = True):
        d1 = _to_device(x1)
        d2 = d1 if x2 is x1 else _to_device(x2)
        out = self.k_device(d1, d2, params)
        return out.cpu().numpy() if as_numpy else out


class _Composite(Kernel):
    """Sum / Schur product of two stationary kernels (gpx/kernels/kernels.py:1794-1942, kernels/operations.py:103-426).
    Parameters: {"kernel1": params1, "kernel2": params2}.  Each component matrix comes from the fused Gram
    kernel; they are combined on the device (gpxb_elementwise).  Supported by the dense exact-GPR paths
    (LML, gradient, fit, predict); nesting is allowed."""

    op = ""

    def __init__(self, kernel1: Kernel, kernel2: Kernel) -> None:
        super().__init__(None)
        self.kernel1, self.kernel2 = kernel1, kernel2

    @property
    def kind(self):
        raise NotImplementedError(f"{self.__class__.__name__} kernels are supported by the dense exact-GPR paths only")

    def lengthscale(self, params):
        raise NotImplementedError(f"{self.__class__.__name__} has one lengthscale per component")

    def default_params(self):
        return {"kernel1": self.kernel1.default_params(), "kernel2": self.kernel2.default_params()}

    def leaves(self, params, prefix=()):
        """[(path prefix, leaf kernel, leaf params)] in evaluation order."""
        out = []
        for name, k in (("kernel1", self.kernel1), ("kernel2", self.kernel2)):
            if isinstance(k, _Composite):
                out += k.leaves(params[name], prefix + (name,))
            else:
                out.append((prefix + (name,), k, params[name]))
        return out

    def program(self, params):
        """The tree in postfix order for gpxb_gram_composite: ([(leaf kernel, leaf params)], [leaf index | OP_SUM | OP_PROD])."""
        from . import ops

        leaves, prog = [], []

        def walk(node, p):
            if isinstance(node, _Composite):
                walk(node.kernel1, p["kernel1"])
                walk(node.kernel2, p["kernel2"])
                prog.append(ops.OP_SUM if node.op == "sum" else ops.OP_PROD)
            else:
                prog.append(len(leaves))
                leaves.append((node, p))

        walk(self, params)
        return leaves, prog

    def k_device(self, x1d, x2d, params):
        """Device kernel matrix from raw (unsliced) device inputs; `x2d is x1d` marks the symmetric case.  One fused
        tile-level pass over all leaves (gpxb_gram_composite) when their operand tiles fit in shared memory and the
        tree is at most 8 leaves / 4 stack levels; else component matrices combined by gpxb_elementwise."""
        from . import ops

        leaves, prog = self.program(params)
        kinds = [k.kind for k, _ in leaves]
        if all(kd in ("se", "m12", "m32", "m52", "linear") for kd in kinds) and len(leaves) <= 8 and _stack_depth(prog) <= 4:
            sym = x2d is x1d
            x1l = [x1d if k.active_dims is None else x1d[:, k.active_dims].contiguous() for k, _ in leaves]
            if ops.gram_composite_fits([t.shape[1] for t in x1l]):
                x2l = None if sym else [x2d if k.active_dims is None else x2d[:, k.active_dims].contiguous()
                                        for k, _ in leaves]
                ells = [1.0 if k.kind == "linear" else k.lengthscale(p) for k, p in leaves]
                return ops.gram_composite(kinds, ells, x1l, x2l, prog)
        mats = [k.k_device(x1d, x2d, params[name])
                for name, k in (("kernel1", self.kernel1), ("kernel2", self.kernel2))]
        return ops.elementwise("axpby" if self.op == "sum" else "mul", mats[0], mats[1], out=mats[0])

    def k(self, x1, x2, params, as_numpy: bool = True):
        d1 = _to_device(x1)
        d2 = d1 if x2 is x1 else _to_device(x2)
        out = self.k_device(d1, d2, params)
        return out.cpu().numpy() if as_numpy else out

    def __repr__(self):
        return f"{self.__class__.__name__}({self.kernel1!r}, {self.kernel2!r})"


def _stack_depth(prog) -> int:
    sp = deep = 0
    for t in prog:
        sp += 1 if t >= 0 else -1
        deep = max(deep, sp)
    return deep


class Sum(_Composite):
    """k1 + k2  (gpx/kernels/kernels.py:1794-1848)."""
    op = "sum"


class Prod(_Composite):
    """k1 * k2, element-wise  (gpx/kernels/kernels.py:1851-1942)."""
    op = "prod"


def _kernel_add(self, other):
    return Sum(self, other)


def _kernel_mul(self, other):
    return Prod(self, other)


Kernel.__add__ = _kernel_add   # gpx/kernels/kernels.py:1490-1500
Kernel.__mul__ = _kernel_mul
