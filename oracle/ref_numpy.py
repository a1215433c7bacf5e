"""CPU oracle: NumPy/SciPy FP64 restatement of the GPX Gaussian-process hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gpx_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may use it, and only as the checker / CPU baseline.

Every function cites the reference file:line (relative to the GPX repository
root) whose arithmetic it restates.  JAX is not installable in this image, so
the reference itself cannot be run here.

Parity pinning status
---------------------
* kernel-matrix entries (SE, Matern-1/2, 3/2, 5/2):  PINNED by the reference's own
  closed-form tests (tests/gpx/kernels/test_kernels.py:58-81, 104-164) restated in
  tests/test_oracle_kernels.py.
* derivative (Hessian/Jacobian) kernels: PINNED the way the reference pins them
  (tests/gpx/kernels/test_kernels.py:172-324: hand-written vs automatic
  differentiation of the ``*_base`` pair kernel) with complex-step / finite
  differences standing in for autodiff.
* Threefry-2x32-20: PINNED by the Random123 known-answer vectors.
* exact-GPR LML, its optimum under L-BFGS-B (hence the gradient's zero), the SGPR LML, and the
  legacy jax.random layout (PRNGKey -> random_bits -> uniform -> normal): PINNED to 8 digits by the
  numbers the reference's own notebooks print (examples/gpr.ipynb cells 11, 14, 19;
  examples/sgpr.ipynb cells 5, 7; tests/golden/notebook_anchors.json): with the data regenerated
  through the restated PRNG this oracle reproduces the fitted lengthscale/sigma and the LML.
* the RPCholesky pivot stream (split -> uniform -> choice -> pivot loop, legacy mode): PINNED by
  examples/sgpr_with_rpcholesky.ipynb cell 7 (the printed end point of a fit whose loss re-draws the
  pivots at every evaluation is reproduced, including the abnormal L-BFGS-B termination).
* Sum / Prod composite kernels with per-leaf active_dims, the Linear and Matern-5/2 entries and the 2-D legacy
  normal() layout: PINNED to 8 digits by the kernel-matrix corners examples/kernel_operations.ipynb prints
  (cells 6, 10, 13, 16); the sign / scaling conventions of d0kj, d1kj, d01kj (the GPR_TD block matrix): PINNED by the
  entries examples/derivatives_solak.ipynb cell 4 prints; the Hessian-kernel LML value: PINNED to 10 digits by the
  scipy OptimizeResult examples/gpr_with_derivatives_only.ipynb cell 27 prints (fun = -1.5536783384709247).
* jax.scipy.sparse.linalg.cg (third-party, restated in `cg`): cross-checked against scipy.sparse.linalg.cg, an
  independent implementation of the same textbook PCG and stopping rule -- identical iteration counts and
  solutions with and without the RPCholesky preconditioner (tests/test_oracle_cpu.py).  Not a pin against JAX
  itself, which cannot run here.
* exact-GPR LML, its gradient, fit coefficients, predictive mean and covariance (all four kernels): cross-checked
  to ~1e-12 against scikit-learn's GaussianProcessRegressor with the equivalent kernels (an independent third-party
  implementation; Matern-1/2 agrees to 1e-6 only, because GPX's 1e-12 jitter under the square root is GPX's own
  semantics) -- tests/test_oracle_cpu.py::test_oracle_exact_gpr_against_scikit_learn.
* predictions and fit coefficients of the sparse / derivative models, Lanczos/SLQ and the partitionable PRNG mode:
  **parity unpinned** -- the reference has no test, golden vector or
  printed output for them and cannot be executed here.  They are anchored
  by self-consistency (analytic gradient vs finite differences, Woodbury vs
  the reference's N x N form, SLQ vs exact log-det).
* ``jax.random`` bit layouts (split / random_bits) are restated from the
  published JAX algorithm for both values of ``jax_threefry_partitionable``
  (default True since JAX 0.5.0); the reference does not pin a JAX version.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

# =============================================================================
# jax.random restatement (third-party dependency of the reference: jax, unpinned;
# call sites gpx/kernels/approximations.py:100-101, gpx/lanczos.py:94,103,168-169)
# =============================================================================

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al., Random123).  Inputs broadcast;
    arithmetic is done in uint64 and masked to 32 bits."""
    k0 = np.asarray(k0, dtype=np.uint64)
    k1 = np.asarray(k1, dtype=np.uint64)
    x0 = np.asarray(x0, dtype=np.uint64).copy()
    x1 = np.asarray(x1, dtype=np.uint64).copy()
    ks = (k0, k1, (k0 ^ k1 ^ np.uint64(0x1BD11BDA)) & _M32)

    def rotl(v, r):
        return ((v << np.uint64(r)) | (v >> np.uint64(32 - r))) & _M32

    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M32
            x1 = rotl(x1, r)
            x1 = x1 ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M32
        x1 = (x1 + ks[(i + 2) % 3] + np.uint64(i + 1)) & _M32
    return x0.astype(np.uint32), x1.astype(np.uint32)


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey: [hi32, lo32] of the 64-bit seed."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def _threefry_2x32_flat(key, counts):
    """jax._src.prng.threefry_2x32 on a flat uint32 counter array (odd sizes are
    zero padded, counters split into first/second half, outputs concatenated)."""
    counts = np.asarray(counts, dtype=np.uint32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, np.uint32)])
    h = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:h], counts[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num: int = 2, partitionable: bool = True) -> np.ndarray:
    """jax.random.split -> (num, 2) uint32."""
    key = np.asarray(key, dtype=np.uint32)
    if partitionable:  # _threefry_split_foldlike
        lo = np.arange(num, dtype=np.uint64)
        b1, b2 = threefry2x32(key[0], key[1], lo >> np.uint64(32), lo & _M32)
        return np.stack([b1, b2], axis=1)
    out = _threefry_2x32_flat(key, np.arange(2 * num, dtype=np.uint32))
    return out.reshape(num, 2)


def random_bits64(key, shape, partitionable: bool = True) -> np.ndarray:
    """jax._src.prng.threefry_random_bits, bit_width=64."""
    key = np.asarray(key, dtype=np.uint32)
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if partitionable:
        idx = np.arange(size, dtype=np.uint64)
        hi, lo = threefry2x32(key[0], key[1], idx >> np.uint64(32), idx & _M32)
    else:
        bits = _threefry_2x32_flat(key, np.arange(2 * size, dtype=np.uint32))
        hi, lo = bits[:size], bits[size:]
    out = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return out.reshape(shape)


def uniform(key, shape=(), partitionable: bool = True) -> np.ndarray:
    """jax.random.uniform(key, shape, float64) in [0, 1)."""
    bits = random_bits64(key, shape, partitionable)
    fb = (bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)
    return fb.view(np.float64) - 1.0


def rademacher(key, shape, partitionable: bool = True) -> np.ndarray:
    """jax.random.rademacher: 2*bernoulli(0.5)-1 with bernoulli = uniform < p."""
    return 2 * (uniform(key, shape, partitionable) < 0.5).astype(np.int64) - 1


def normal(key, shape, partitionable: bool = True) -> np.ndarray:
    """jax.random.normal float64: sqrt(2)*erfinv(uniform(-1+eps, 1)).  Only used on
    the Lanczos breakdown branch (gpx/lanczos.py:94-96); erfinv is SciPy's, so the
    last bits may differ from XLA's polynomial."""
    from scipy.special import erfinv

    lo = np.nextafter(-1.0, 0.0)
    u = uniform(key, shape, partitionable)
    u = np.maximum(lo, u * (1.0 - lo) + lo)
    return math.sqrt(2.0) * erfinv(u)


def choice_p(key, p: np.ndarray, partitionable: bool = True) -> int:
    """jax.random.choice(key, arange(n), p=p, shape=(), replace=True)."""
    cum = np.cumsum(p)
    r = cum[-1] * (1.0 - uniform(key, (), partitionable))
    return int(np.searchsorted(cum, r, side="left"))


# =============================================================================
# bijectors / mean functions (gpx/bijectors.py:27-71, gpx/mean_functions.py:25-30)
# =============================================================================


def softplus(t):
    return np.logaddexp(t, 0.0) + 1e-300


def inverse_softplus(x):
    return x + np.log(-np.expm1(-x))


def sigmoid(t):
    return 1.0 / (1.0 + np.exp(-t))


def data_mean(y):
    return float(np.mean(y))


def zero_mean(y):
    return 0.0


# =============================================================================
# kernels (gpx/utils.py:55-72; gpx/kernels/kernels.py:751-757, 922-927, 966-971,
# 1050-1055)
# =============================================================================

KINDS = ("se", "m12", "m32", "m52")


def squared_distances(x1, x2):
    """gpx/utils.py:55-72 -- GEMM trick with +1e-12 jitter, no clamp."""
    x1s = np.sum(np.square(x1), axis=-1)
    x2s = np.sum(np.square(x2), axis=-1)
    return x1s[:, None] - 2.0 * np.dot(x1, x2.T) + x2s + 1e-12


def kernel(kind: str, x1, x2, ell: float):
    """Batched stationary kernels on z = x / ell."""
    z1 = x1 / ell
    z2 = x2 / ell
    d2 = squared_distances(z1, z2)
    if kind == "se":  # kernels.py:751-757
        return np.exp(-d2)
    r = np.sqrt(np.maximum(d2, 1e-36))
    if kind == "m12":  # kernels.py:922-927
        return np.exp(-r)
    if kind == "m32":  # kernels.py:966-971
        d = np.sqrt(3.0) * r
        return (1.0 + d) * np.exp(-d)
    if kind == "m52":  # kernels.py:1050-1055
        d = np.sqrt(5.0) * r
        return (1.0 + d + d**2 / 3.0) * np.exp(-d)
    raise ValueError(kind)


def composite_kernel(spec, x1, x2):
    """Sum / Prod kernels (kernels.py:1794-1942; operations.py:103-426).  spec: ("leaf", kind, ell, active_dims)
    or ("sum" | "prod", spec1, spec2); kind "linear" is x1 . x2 on the active dimensions (kernels.py:226-236),
    its ell entry is ignored."""
    if spec[0] == "leaf":
        _, kind, ell, dims = spec
        a, b = (x1, x2) if dims is None else (x1[:, dims], x2[:, dims])
        if kind == "linear":
            return a @ b.T
        return kernel(kind, a, b, ell)
    K1, K2 = composite_kernel(spec[1], x1, x2), composite_kernel(spec[2], x1, x2)
    return K1 + K2 if spec[0] == "sum" else K1 * K2


def composite_lml_dense(spec, x, y, sigma, mean_function=data_mean):
    """_gpr.py:737-769 with a composite kernel."""
    m = y.shape[0]
    r = y - mean_function(y)
    A = composite_kernel(spec, x, x) + (sigma**2 + 1e-10) * np.eye(m)
    L = sla.cholesky(A, lower=True, check_finite=False)
    cy = sla.solve_triangular(L, r, lower=True, check_finite=False)
    return (-0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(L))) - m * 0.5 * np.log(2.0 * np.pi)) / m


def kernel_dl(kind: str, x1, x2, ell: float):
    """(K, dK/d ell): what reverse-mode autodiff of ``kernel`` yields.
    s = d2 - jitter is formed exactly as autodiff forms it: -ell/2 * d(d2)/d ell."""
    z1 = x1 / ell
    z2 = x2 / ell
    d2 = squared_distances(z1, z2)
    s = d2 - 1e-12
    if kind == "se":
        K = np.exp(-d2)
        return K, K * 2.0 * s / ell
    live = d2 > 1e-36  # the maximum() passes no gradient when clamped
    r = np.sqrt(np.maximum(d2, 1e-36))
    if kind == "m12":
        K = np.exp(-r)
        return K, np.where(live, K * s / (r * ell), 0.0)
    if kind == "m32":
        d = np.sqrt(3.0) * r
        e = np.exp(-d)
        return (1.0 + d) * e, np.where(live, 3.0 * s * e / ell, 0.0)
    if kind == "m52":
        d = np.sqrt(5.0) * r
        e = np.exp(-d)
        return (1.0 + d + d**2 / 3.0) * e, np.where(live, (5.0 * s / (3.0 * ell)) * (1.0 + d) * e, 0.0)
    raise ValueError(kind)


def kernel_base(kind: str, x1, x2, ell: float):
    """Single-pair kernels without jitter (kernels.py:731-736, 902-910, 945-953,
    1029-1037); the autodiff kernelizers differentiate these."""
    z = (np.asarray(x1) - np.asarray(x2)) / ell
    d2 = np.sum(z * z)
    if kind == "se":
        return np.exp(-d2)
    c = {"m12": 1.0, "m32": 3.0, "m52": 5.0}[kind]
    d = np.sqrt(np.maximum(c * d2, 1e-36))
    if kind == "m12":
        return np.exp(-d)
    if kind == "m32":
        return (1.0 + d) * np.exp(-d)
    return (1.0 + d + c * d2 / 3.0) * np.exp(-d)


def d01kj(kind: str, x1, x2, ell: float, J1, J2):
    """Hessian-Jacobian kernel, literal einsum form.
    SE: kernels.py:804-831; Matern-5/2: kernels.py:1253-1283."""
    ns1, nf = x1.shape
    ns2 = x2.shape[0]
    nv1, nv2 = J1.shape[2], J2.shape[2]
    z1, z2 = x1 / ell, x2 / ell
    d2 = squared_distances(z1, z2)
    if kind == "se":
        ed2 = np.exp(-d2)
        diff = (2.0 / ell) * (z1[:, None] - z2)
        c_ab = -ed2
        c_g = ed2 * (2.0 / ell**2)
    elif kind == "m52":
        diff = np.sqrt(5.0) * (z1[:, None] - z2)
        d = np.sqrt(5.0) * np.sqrt(np.maximum(d2, 1e-36))
        c = (5.0 / (3.0 * ell**2)) * np.exp(-d)
        c_ab = -c
        c_g = c * (1.0 + d)
    else:
        raise ValueError(kind)
    a = np.einsum("stf,sfv->stv", diff, J1)
    b = np.einsum("stf,tfu->stu", diff, J2)
    out = np.einsum("st,stv,stu->svtu", c_ab, a, b)
    G = np.einsum("sfv,tfu->svtu", J1, J2)
    out = out + c_g[:, None, :, None] * G
    return out.reshape(ns1 * nv1, ns2 * nv2)


def d01kjc(kind: str, x1, x2, ell: float, jaccoef, J2):
    """Hessian kernel contracted with jaccoef on the first side
    (kernels.py:851-877 SE, 1304-1332 Matern-5/2) -> (ns1, ns2*nv2)."""
    out = d01kj(kind, x1, x2, ell, jaccoef[:, :, None], J2)
    return out.reshape(x1.shape[0], -1)


def d0kjc(kind: str, x1, x2, ell: float, jaccoef):
    """Kernel.d0kjc (kernels.py:1474-1480: grad_kernelize(argnums=0, with_jaccoef, trace_samples=False)) of
    the pair kernels kernels.py:731-737 (SE, plain autodiff: no 1e-12 jitter in the pair form) and 1030-1037
    with its custom JVP 1068-1075 (Matern-5/2) -> (ns1, ns2): sum_f jaccoef[s,f] d k(x1_s, x2_t)/d x1_sf."""
    dz = x1[:, None, :] / ell - x2[None, :, :] / ell
    if kind == "se":
        k = np.exp(-np.sum(dz**2, axis=2))
        g = -(2.0 / ell) * dz * k[:, :, None]
    elif kind == "m52":
        diff = np.sqrt(5.0) * dz
        d = np.sqrt(np.maximum(np.sum(diff**2, axis=2), 1e-36))
        g = -(np.sqrt(5.0) / (3.0 * ell)) * ((1.0 + d) * np.exp(-d))[:, :, None] * diff
    else:
        raise ValueError(kind)
    return np.einsum("sf,stf->st", jaccoef, g)


def d0kj(kind: str, x1, x2, ell: float, J1):
    """Kernel.d0kj -> (ns1*nv, ns2).  SE: grad_kernelize(argnums=0, with_jacob) of the pair kernel
    kernels.py:731-737, 1459-1462; Matern-5/2: the hand-written kernels.py:1108-1149 (d2 from squared_distances,
    i.e. with the 1e-12 jitter)."""
    ns1, nf = x1.shape
    nv = J1.shape[2]
    z1, z2 = x1 / ell, x2 / ell
    dz = z1[:, None, :] - z2[None, :, :]
    if kind == "se":
        d0k = -(2.0 / ell) * dz * np.exp(-np.sum(dz**2, axis=2))[:, :, None]
    elif kind == "m52":
        d2 = squared_distances(z1, z2)
        d = np.sqrt(5.0) * np.sqrt(np.maximum(d2, 1e-36))
        d0k = -(np.sqrt(5.0) / (3.0 * ell)) * ((1.0 + d) * np.exp(-d))[:, :, None] * (np.sqrt(5.0) * dz)
    else:
        raise ValueError(kind)
    return np.einsum("sfv,stf->svt", J1, d0k).reshape(ns1 * nv, x2.shape[0])


def predict_derivs_dense_full(kind, x_train, J_train, x, J, c, mu, ell, sigma):
    """_gpr.py:488-560 without jaccoef, full_covariance=True: returns (mu (ns*nv, 1), C_nn)."""
    K_mn = d01kj(kind, x_train, x, ell, J_train, J)
    mean = mu + K_mn.T @ c.reshape(-1, 1)
    C_mm = d01kj(kind, x_train, x_train, ell, J_train, J_train) + (sigma**2 + 1e-10) * np.eye(K_mn.shape[0])
    L = sla.cholesky(C_mm, lower=True, check_finite=False)
    G = sla.solve_triangular(L, K_mn, lower=True, check_finite=False)
    return mean, d01kj(kind, x, x, ell, J, J) - G.T @ G


def predict_y_derivs_dense_full(kind, x_train, J_train, x, c, mu, ell, sigma):
    """_gpr.py:608-660 without jaccoef, full_covariance=True (sigma^2 I without the 1e-10, as written there)."""
    K_mn = d0kj(kind, x_train, x, ell, J_train)
    mean = mu + K_mn.T @ c.reshape(-1, 1)
    C_mm = d01kj(kind, x_train, x_train, ell, J_train, J_train) + sigma**2 * np.eye(K_mn.shape[0])
    L = sla.cholesky(C_mm, lower=True, check_finite=False)
    G = sla.solve_triangular(L, K_mn, lower=True, check_finite=False)
    return mean, kernel(kind, x, x, ell) - G.T @ G


def td_A_lhs(kind, x1, J1, x2, J2, ell, sigma_t, sigma_d, noise=True):
    """gpr_td.py:43-159 (one derivative block): [[K + s_t I, d1kj], [d0kj, d01kj + s_d I]];
    d1kj(x1, x2, J2) = d0kj(x2, x1, J2)^T."""
    K = kernel(kind, x1, x2, ell)
    D01 = d01kj(kind, x1, x2, ell, J1, J2)
    if noise:
        K = K + (sigma_t**2 + 1e-10) * np.eye(K.shape[0])
        D01 = D01 + (sigma_d**2 + 1e-10) * np.eye(D01.shape[0])
    D0 = d0kj(kind, x1, x2, ell, J1)
    D1 = d0kj(kind, x2, x1, ell, J2).T
    return np.block([[K, D1], [D0, D01]])


def td_lml_dense(kind, x, y, J, y_derivs, ell, sigma_t, sigma_d, mean_function=data_mean):
    """gpr_td.py:507-560."""
    mu = mean_function(y)
    ym = np.concatenate([(y - mu).reshape(-1, 1), y_derivs.reshape(-1, 1)])
    m = ym.shape[0]
    A = td_A_lhs(kind, x, J, x, J, ell, sigma_t, sigma_d)
    L = sla.cholesky(A, lower=True, check_finite=False)
    cy = sla.solve_triangular(L, ym, lower=True, check_finite=False)
    return (-0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(L))) - m * 0.5 * np.log(2.0 * np.pi)) / m


def td_fit_dense(kind, x, y, J, y_derivs, ell, sigma_t, sigma_d, mean_function=data_mean):
    """gpr_td.py:639-683."""
    mu = mean_function(y)
    ym = np.concatenate([(y - mu).reshape(-1, 1), y_derivs.reshape(-1, 1)])
    return np.linalg.solve(td_A_lhs(kind, x, J, x, J, ell, sigma_t, sigma_d), ym), mu


def td_predict_dense(kind, x_train, J_train, x, J, c, mu, ell):
    """gpr_td.py:749-797."""
    K_mn = td_A_lhs(kind, x_train, J_train, x, J, ell, 0.0, 0.0, noise=False)
    pred = K_mn.T @ c
    ns, _, nv = J.shape
    pred[:ns] += mu
    return pred[:ns], pred[ns:ns + ns * nv].reshape(ns, -1)


def rpcholesky_td(kind, x, J, n_pivots, ell, key, partitionable=True):
    """gpx/models/gpr_td.py:313-460 (one derivative block): randomly pivoted Cholesky of the noise-free block kernel
    [[k, d1kj], [d0kj, d01kj]] over the rows [targets; derivatives (sample, variable)]."""
    n = x.shape[0]
    nv = J.shape[2]
    ntot = n + n * nv
    F = np.zeros((ntot, n_pivots))
    pivots = np.zeros(n_pivots, dtype=np.int64)
    dk = np.array([kernel(kind, x[i:i + 1], x[i:i + 1], ell)[0, 0] for i in range(n)])
    dh = np.concatenate([np.diagonal(d01kj(kind, x[i:i + 1], x[i:i + 1], ell, J[i:i + 1], J[i:i + 1]))
                         for i in range(n)])
    diag = np.concatenate([dk, dh])
    key = np.asarray(key, dtype=np.uint32)
    for i in range(n_pivots):
        prob = diag / np.sum(diag)
        key = split(key, 2, partitionable)[0]   # `key, subkey = split(key)`, then choice(key, ...)  (:387-388)
        s = choice_p(key, prob, partitionable)
        pivots[i] = s
        if s < n:      # first_row (:397-407)
            xs = x[s:s + 1]
            row = np.concatenate([kernel(kind, xs, x, ell)[0], d0kj(kind, x, xs, ell, J).reshape(-1)])
        else:          # second_row (:409-421)
            i1, i2 = (int(s) - n) // nv, (int(s) - n) % nv
            xs, js = x[i1:i1 + 1], J[i1:i1 + 1]
            row = np.concatenate([d0kj(kind, xs, x, ell, js)[i2], d01kj(kind, xs, x, ell, js, J)[i2]])
        g = row - F @ F[s]
        F[:, i] = g / (g[s] ** 0.5 + 1e-6)
        diag = diag - F[:, i] ** 2
    return F, pivots


def td_fit_iter(kind, x, y, J, y_derivs, ell, sigma_t, sigma_d, n_pivots, key_precond, mean_function=data_mean,
                partitionable=True, cg_tol=1e-5, cg_atol=1e-7):
    """gpr_td.py:685-747 with _precond_rpcholesky_TD :463-504 (sigma_targets in the preconditioner for every row)."""
    mu = mean_function(y)
    ym = np.concatenate([(y - mu).reshape(-1, 1), y_derivs.reshape(-1, 1)])
    A = td_A_lhs(kind, x, J, x, J, ell, sigma_t, sigma_d)
    F, _ = rpcholesky_td(kind, x, J, n_pivots, ell, key_precond, partitionable)
    P = np.linalg.inv(sigma_t**2 * np.eye(n_pivots) + F.T @ F)

    def precond(z):
        res = P @ (F.T @ z)
        return -(sigma_t ** (-2)) * (F @ res) + sigma_t ** (-2) * z

    c, iters = cg(lambda v: A @ v, ym, M=precond, tol=cg_tol, atol=cg_atol)
    return c, mu, iters


def td_lml_iter(kind, x, y, J, y_derivs, ell, sigma_t, sigma_d, num_evals, num_lanczos, key_lanczos, n_pivots,
                key_precond, mean_function=data_mean, partitionable=True, cg_tol=1e-5, cg_atol=1e-7):
    """gpr_td.py:562-637."""
    c, mu, _ = td_fit_iter(kind, x, y, J, y_derivs, ell, sigma_t, sigma_d, n_pivots, key_precond, mean_function,
                           partitionable, cg_tol, cg_atol)
    ym = np.concatenate([(y - mu).reshape(-1, 1), y_derivs.reshape(-1, 1)])
    m = ym.shape[0]
    A = td_A_lhs(kind, x, J, x, J, ell, sigma_t, sigma_d)
    mll = -0.5 * np.sum(ym.T @ c)
    mll -= 0.5 * lanczos_logdet(lambda v: A @ v, num_evals, m, num_lanczos, key_lanczos, partitionable)
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def td2_A_lhs(kind, x1, J1a, J1b, x2, J2a, J2b, ell, sigma_t, sigma_d, sigma_d2, noise=True):
    """gpr_td.py:43-159 with the second derivative block, literal 3 x 3 assembly:
    [[K, d1kj_1, d1kj_2], [d0kj_1, d01kj_11, d01kj_12], [d0kj_2, d01kj_21, d01kj_22]]."""
    K = kernel(kind, x1, x2, ell)
    D11 = d01kj(kind, x1, x2, ell, J1a, J2a)
    D22 = d01kj(kind, x1, x2, ell, J1b, J2b)
    if noise:
        K = K + (sigma_t**2 + 1e-10) * np.eye(K.shape[0])
        D11 = D11 + (sigma_d**2 + 1e-10) * np.eye(D11.shape[0])
        D22 = D22 + (sigma_d2**2 + 1e-10) * np.eye(D22.shape[0])
    D12 = d01kj(kind, x1, x2, ell, J1a, J2b)
    D21 = d01kj(kind, x1, x2, ell, J1b, J2a)
    D0a, D0b = d0kj(kind, x1, x2, ell, J1a), d0kj(kind, x1, x2, ell, J1b)
    D1a, D1b = d0kj(kind, x2, x1, ell, J2a).T, d0kj(kind, x2, x1, ell, J2b).T
    return np.block([[K, D1a, D1b], [D0a, D11, D12], [D0b, D21, D22]])


def rpcholesky_td2(kind, x, Ja, Jb, n_pivots, ell, key, partitionable=True):
    """gpx/models/gpr_td.py:313-460 with jacobian_2: rows [targets; block 1 (sample, variable); block 2]."""
    n = x.shape[0]
    n1, n2 = Ja.shape[2], Jb.shape[2]
    ntot = n + n * n1 + n * n2
    F = np.zeros((ntot, n_pivots))
    pivots = np.zeros(n_pivots, dtype=np.int64)
    dk = np.array([kernel(kind, x[i:i + 1], x[i:i + 1], ell)[0, 0] for i in range(n)])
    da = np.concatenate([np.diagonal(d01kj(kind, x[i:i + 1], x[i:i + 1], ell, Ja[i:i + 1], Ja[i:i + 1])) for i in range(n)])
    db = np.concatenate([np.diagonal(d01kj(kind, x[i:i + 1], x[i:i + 1], ell, Jb[i:i + 1], Jb[i:i + 1])) for i in range(n)])
    diag = np.concatenate([dk, da, db])
    key = np.asarray(key, dtype=np.uint32)
    for i in range(n_pivots):
        prob = diag / np.sum(diag)
        key = split(key, 2, partitionable)[0]
        s = int(choice_p(key, prob, partitionable))
        pivots[i] = s
        if s < n:                       # first_row (:397-407)
            xs = x[s:s + 1]
            row = np.concatenate([kernel(kind, xs, x, ell)[0], d0kj(kind, x, xs, ell, Ja).reshape(-1),
                                  d0kj(kind, x, xs, ell, Jb).reshape(-1)])
        else:                           # second_row / third_row (:409-436)
            first = s < n + n * n1
            q = s - n if first else s - n - n * n1
            nvq, Jq = (n1, Ja) if first else (n2, Jb)
            i1, i2 = q // nvq, q % nvq
            xs, js = x[i1:i1 + 1], Jq[i1:i1 + 1]
            row = np.concatenate([d0kj(kind, xs, x, ell, js)[i2], d01kj(kind, xs, x, ell, js, Ja)[i2],
                                  d01kj(kind, xs, x, ell, js, Jb)[i2]])
        g = row - F @ F[s]
        F[:, i] = g / (g[s] ** 0.5 + 1e-6)
        diag = diag - F[:, i] ** 2
    return F, pivots


def td2_fit_iter(kind, x, y, Ja, Jb, yd_a, yd_b, ell, sigma_t, sigma_d, sigma_d2, n_pivots, key_precond,
                 mean_function=data_mean, partitionable=True, cg_tol=1e-5, cg_atol=1e-7):
    """gpr_td.py:685-747 with jacobian_2 / y_derivs_2 (c in the reference's order)."""
    mu = mean_function(y)
    ym = np.concatenate([(y - mu).reshape(-1, 1), yd_a.reshape(-1, 1), yd_b.reshape(-1, 1)])
    A = td2_A_lhs(kind, x, Ja, Jb, x, Ja, Jb, ell, sigma_t, sigma_d, sigma_d2)
    F, _ = rpcholesky_td2(kind, x, Ja, Jb, n_pivots, ell, key_precond, partitionable)
    P = np.linalg.inv(sigma_t**2 * np.eye(n_pivots) + F.T @ F)

    def precond(z):
        return -(sigma_t ** (-2)) * (F @ (P @ (F.T @ z))) + sigma_t ** (-2) * z

    c, iters = cg(lambda v: A @ v, ym, M=precond, tol=cg_tol, atol=cg_atol)
    return c, mu, iters


def td2_lml_iter(kind, x, y, Ja, Jb, yd_a, yd_b, ell, sigma_t, sigma_d, sigma_d2, num_evals, num_lanczos, key_lanczos,
                 n_pivots, key_precond, mean_function=data_mean, partitionable=True, cg_tol=1e-5, cg_atol=1e-7):
    """gpr_td.py:562-637 with jacobian_2 / y_derivs_2."""
    c, mu, _ = td2_fit_iter(kind, x, y, Ja, Jb, yd_a, yd_b, ell, sigma_t, sigma_d, sigma_d2, n_pivots, key_precond,
                            mean_function, partitionable, cg_tol, cg_atol)
    ym = np.concatenate([(y - mu).reshape(-1, 1), yd_a.reshape(-1, 1), yd_b.reshape(-1, 1)])
    m = ym.shape[0]
    A = td2_A_lhs(kind, x, Ja, Jb, x, Ja, Jb, ell, sigma_t, sigma_d, sigma_d2)
    mll = -0.5 * np.sum(ym.T @ c)
    mll -= 0.5 * lanczos_logdet(lambda v: A @ v, num_evals, m, num_lanczos, key_lanczos, partitionable)
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def td2_lml_dense(kind, x, y, Ja, Jb, yd_a, yd_b, ell, sigma_t, sigma_d, sigma_d2, mean_function=data_mean):
    """gpr_td.py:507-560 with jacobian_2 / y_derivs_2."""
    mu = mean_function(y)
    ym = np.concatenate([(y - mu).reshape(-1, 1), yd_a.reshape(-1, 1), yd_b.reshape(-1, 1)])
    m = ym.shape[0]
    A = td2_A_lhs(kind, x, Ja, Jb, x, Ja, Jb, ell, sigma_t, sigma_d, sigma_d2)
    L = sla.cholesky(A, lower=True, check_finite=False)
    cy = sla.solve_triangular(L, ym, lower=True, check_finite=False)
    return (-0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(L))) - m * 0.5 * np.log(2.0 * np.pi)) / m


def td2_fit_dense(kind, x, y, Ja, Jb, yd_a, yd_b, ell, sigma_t, sigma_d, sigma_d2, mean_function=data_mean):
    """gpr_td.py:639-683 with jacobian_2 / y_derivs_2: c in the reference's order [targets; block 1; block 2]."""
    mu = mean_function(y)
    ym = np.concatenate([(y - mu).reshape(-1, 1), yd_a.reshape(-1, 1), yd_b.reshape(-1, 1)])
    return np.linalg.solve(td2_A_lhs(kind, x, Ja, Jb, x, Ja, Jb, ell, sigma_t, sigma_d, sigma_d2), ym), mu


def td2_predict_dense(kind, x_train, Ja_train, Jb_train, x, Ja, Jb, c, mu, ell):
    """gpr_td.py:749-797 with jacobian_2: (targets (ns, 1), derivatives block 1 (ns, nv_1), block 2 (ns, nv_2))."""
    K_mn = td2_A_lhs(kind, x_train, Ja_train, Jb_train, x, Ja, Jb, ell, 0.0, 0.0, 0.0, noise=False)
    pred = K_mn.T @ c
    ns, n1, n2 = x.shape[0], Ja.shape[2], Jb.shape[2]
    pred[:ns] += mu
    return pred[:ns], pred[ns:ns + ns * n1].reshape(ns, -1), pred[ns + ns * n1:ns + ns * (n1 + n2)].reshape(ns, -1)


def predict_y_derivs_dense(kind, x_train, x, jaccoef, mu, ell):
    """_gpr.py:630-638 (contracted-jacobian fast path of _predict_y_derivs_dense)."""
    return (mu + np.sum(d0kjc(kind, x_train, x, ell, jaccoef), axis=0)).reshape(-1, 1)


# =============================================================================
# exact GPR (gpx/models/_gpr.py)
# =============================================================================


def A_lhs(kind, x1, x2, ell, sigma, noise=True):
    """_gpr.py:39-63."""
    C = kernel(kind, x1, x2, ell)
    if noise:
        C = C + (sigma**2 + 1e-10) * np.eye(x1.shape[0])
    return C


def lml_dense(kind, x, y, ell, sigma, mean_function=data_mean):
    """_gpr.py:737-769."""
    m = y.shape[0]
    mu = mean_function(y)
    r = y - mu
    A = A_lhs(kind, x, x, ell, sigma)
    L = sla.cholesky(A, lower=True, check_finite=False)
    cy = sla.solve_triangular(L, r, lower=True, check_finite=False)
    mll = -0.5 * np.sum(np.square(cy))
    mll -= np.sum(np.log(np.diag(L)))
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def kernel_dl_threaded(kind, x1, x2, ell, threads):
    """kernel_dl evaluated in row blocks on a thread pool (NumPy releases the GIL), the way XLA's CPU
    backend parallelises element-wise work; same arithmetic per entry."""
    from concurrent.futures import ThreadPoolExecutor

    n = x1.shape[0]
    K = np.empty((n, x2.shape[0]))
    dK = np.empty_like(K)
    step = max(64, -(-n // (4 * threads)))

    def work(i0):
        i1 = min(n, i0 + step)
        K[i0:i1], dK[i0:i1] = kernel_dl(kind, x1[i0:i1], x2, ell)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(0, n, step)))
    return K, dK


def lml_grad_dense(kind, x, y, ell, sigma, mean_function=data_mean, threads=1):
    """(lml, d lml/d ell, d lml/d sigma): value_and_grad of _lml_dense
    (optimizers/scipy_optimize.py:72-78) in analytic form.  threads > 1 only
    parallelises the element-wise kernel evaluation (used by the timed CPU baseline)."""
    m = y.shape[0]
    mu = mean_function(y)
    r = y - mu
    K, dK = kernel_dl(kind, x, x, ell) if threads <= 1 else kernel_dl_threaded(kind, x, x, ell, threads)
    A = K + (sigma**2 + 1e-10) * np.eye(m)
    L = sla.cholesky(A, lower=True, check_finite=False)
    cy = sla.solve_triangular(L, r, lower=True, check_finite=False)
    lml = (-0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(L))) - m * 0.5 * np.log(2.0 * np.pi)) / m
    alpha = sla.solve_triangular(L, cy, lower=True, trans="T", check_finite=False)
    Ainv, info = sla.lapack.dpotri(L, lower=1)  # A^-1 from the factor (lower triangle), then mirrored
    Ainv = Ainv + Ainv.T - np.diag(np.diag(Ainv))  # the strict upper triangle of the factor is zero
    # multi-output y (m, t): quadratic summed over outputs, log-det counted once
    W = alpha.reshape(m, -1) @ alpha.reshape(m, -1).T - Ainv
    g_ell = 0.5 * np.sum(W * dK) / m
    g_sigma = sigma * np.trace(W) / m
    return lml, g_ell, g_sigma


def lml_grad_dense_lowmem(kind, x, y, ell, sigma, mean_function=data_mean, threads=1, timings=None):
    """lml_grad_dense for sizes where its temporaries do not fit in host memory (N = 32768: K, dK, A, L,
    A^-1, its mirrored copy and W would need ~60 GB).  Same algorithm and the same arithmetic per entry
    (_gpr.py:737-769); the differences are storage only: the noise is added on K's diagonal in
    place, LAPACK's dpotrf / dpotri overwrite that buffer (through its transposed view, which is
    Fortran-contiguous), and sum(W o dK) is accumulated in row blocks over the lower triangle
    (W and dK are symmetric).  Peak memory: two N x N matrices.  Equal to lml_grad_dense to ~1e-14
    relative (tests/test_oracle_cpu.py).  `timings` (dict) receives the seconds spent per phase."""
    import time

    m = y.shape[0]
    mu = mean_function(y)
    r = y - mu
    t0 = time.perf_counter()
    K, dK = kernel_dl(kind, x, x, ell) if threads <= 1 else kernel_dl_threaded(kind, x, x, ell, threads)
    K[np.diag_indices(m)] += sigma**2 + 1e-10
    t1 = time.perf_counter()
    # K is symmetric: its transposed view is Fortran-contiguous, so LAPACK factorises in place
    L = sla.cholesky(K.T, lower=True, overwrite_a=True, check_finite=False)
    t2 = time.perf_counter()
    cy = sla.solve_triangular(L, r, lower=True, check_finite=False)
    lml = (-0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(L))) - m * 0.5 * np.log(2.0 * np.pi)) / m
    alpha = sla.solve_triangular(L, cy, lower=True, trans="T", check_finite=False).reshape(m, -1)
    t3 = time.perf_counter()
    Ainv, info = sla.lapack.dpotri(L, lower=1, overwrite_c=1)  # lower triangle of A^-1, in place
    t4 = time.perf_counter()
    acc, tr = 0.0, 0.0
    step = 1024
    for i0 in range(0, m, step):
        i1 = min(m, i0 + step)
        Wb = alpha[i0:i1] @ alpha[:i1].T - Ainv[i0:i1, :i1]  # rows i0..i1 of W, columns 0..i1
        Wb *= dK[i0:i1, :i1]
        d = np.diagonal(Wb[:, i0:i1]).copy()
        blk = np.tril(Wb[:, i0:i1], -1)
        acc += 2.0 * (np.sum(Wb[:, :i0]) + np.sum(blk)) + np.sum(d)
        tr += np.sum(np.sum(alpha[i0:i1] ** 2, axis=1) - np.diagonal(Ainv)[i0:i1])
    t5 = time.perf_counter()
    if timings is not None:
        timings.update(kernel=t1 - t0, potrf=t2 - t1, solves=t3 - t2, potri=t4 - t3, contraction=t5 - t4)
    return lml, 0.5 * acc / m, sigma * tr / m


def neg_lml_and_grad_unconstrained(kind, x, y, t_ell, t_sigma, mean_function=data_mean):
    """loss(t) and d loss/d t with Softplus bijectors on both parameters
    (optimizers/utils.py:97-131, bijectors.py:27-47)."""
    ell, sigma = softplus(t_ell), softplus(t_sigma)
    lml, gl, gs = lml_grad_dense(kind, x, y, ell, sigma, mean_function)
    return -lml, np.array([-gl * sigmoid(t_ell), -gs * sigmoid(t_sigma)])


def fit_dense(kind, x, y, ell, sigma, mean_function=data_mean):
    """_gpr.py:262-282 (LU solve)."""
    mu = mean_function(y)
    A = A_lhs(kind, x, x, ell, sigma)
    c = np.linalg.solve(A, y - mu)
    return c, mu


def predict_dense(kind, x_train, x, c, mu, ell, sigma, full_covariance=False):
    """_gpr.py:434-463."""
    K_mn = kernel(kind, x_train, x, ell)
    mean = mu + K_mn.T @ c
    if not full_covariance:
        return mean
    A = A_lhs(kind, x_train, x_train, ell, sigma)
    L = sla.cholesky(A, lower=True, check_finite=False)
    G = sla.solve_triangular(L, K_mn, lower=True, check_finite=False)
    C = kernel(kind, x, x, ell) - G.T @ G
    return mean, C


def lml_derivs_dense(kind, x, y, J, ell, sigma):
    """_gpr.py:824-870 (zero mean forced by gpr.py:160,173)."""
    yf = y.reshape(-1, 1)
    m = yf.shape[0]
    A = d01kj(kind, x, x, ell, J, J) + (sigma**2 + 1e-10) * np.eye(m)
    L = sla.cholesky(A, lower=True, check_finite=False)
    cy = sla.solve_triangular(L, yf, lower=True, check_finite=False)
    mll = -0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(L))) - m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def d01kj_dl(kind: str, x1, x2, ell: float, J1, J2):
    """d/d lengthscale of d01kj, obtained by differentiating the literal einsum form term by term
    (what jax.grad does to kernels.py:804-831 / 1253-1283): z = x/ell, so d2 - 1e-12 scales as ell^-2
    (the jitter is a constant; the Matern clamp passes no gradient), diff as ell^-2 (SE) or ell^-1 (M52)."""
    ns1, nf = x1.shape
    ns2 = x2.shape[0]
    nv1, nv2 = J1.shape[2], J2.shape[2]
    z1, z2 = x1 / ell, x2 / ell
    d2 = squared_distances(z1, z2)
    dd2 = -2.0 * (d2 - 1e-12) / ell
    if kind == "se":
        ed2 = np.exp(-d2)
        ded2 = -ed2 * dd2
        diff = (2.0 / ell) * (z1[:, None] - z2)
        ddiff = -2.0 * diff / ell
        c_ab, dc_ab = -ed2, -ded2
        c_g = ed2 * (2.0 / ell**2)
        dc_g = ded2 * (2.0 / ell**2) - 4.0 * ed2 / ell**3
    elif kind == "m52":
        diff = np.sqrt(5.0) * (z1[:, None] - z2)
        ddiff = -diff / ell
        rr = np.sqrt(np.maximum(d2, 1e-36))
        d = np.sqrt(5.0) * rr
        dd = np.where(d2 > 1e-36, np.sqrt(5.0) * dd2 / (2.0 * rr), 0.0)
        c = (5.0 / (3.0 * ell**2)) * np.exp(-d)
        dc = c * (-2.0 / ell - dd)
        c_ab, dc_ab = -c, -dc
        c_g = c * (1.0 + d)
        dc_g = dc * (1.0 + d) + c * dd
    else:
        raise ValueError(kind)
    a = np.einsum("stf,sfv->stv", diff, J1)
    b = np.einsum("stf,tfu->stu", diff, J2)
    da = np.einsum("stf,sfv->stv", ddiff, J1)
    db = np.einsum("stf,tfu->stu", ddiff, J2)
    out = np.einsum("st,stv,stu->svtu", dc_ab, a, b)
    out += np.einsum("st,stv,stu->svtu", c_ab, da, b)
    out += np.einsum("st,stv,stu->svtu", c_ab, a, db)
    G = np.einsum("sfv,tfu->svtu", J1, J2)
    out = out + dc_g[:, None, :, None] * G
    return out.reshape(ns1 * nv1, ns2 * nv2)


def lml_derivs_grad_dense(kind, x, y, J, ell, sigma):
    """(lml, d lml/d ell, d lml/d sigma) of lml_derivs_dense: the pair jit(value_and_grad(loss)) returns for
    neg_log_marginal_likelihood_derivs (gpr.py:276-300; scipy_optimize.py:72-78, 92-122), analytic:
    W = alpha alpha^T - A^-1, d lml/d theta = (1/2 sum W o dA/d theta)/m, dA/d sigma = 2 sigma I."""
    yf = y.reshape(-1, 1)
    m = yf.shape[0]
    A = d01kj(kind, x, x, ell, J, J) + (sigma**2 + 1e-10) * np.eye(m)
    L = sla.cholesky(A, lower=True, check_finite=False)
    cy = sla.solve_triangular(L, yf, lower=True, check_finite=False)
    mll = (-0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(L))) - m * 0.5 * np.log(2.0 * np.pi)) / m
    alpha = sla.solve_triangular(L, cy, lower=True, trans="T", check_finite=False)
    Ainv, _ = sla.lapack.dpotri(L, lower=1)
    Ainv = np.tril(Ainv) + np.tril(Ainv, -1).T
    W = alpha @ alpha.T - Ainv
    dA = d01kj_dl(kind, x, x, ell, J, J)
    g_ell = 0.5 * np.sum(W * dA) / m
    g_sigma = sigma * np.trace(W) / m
    return mll, g_ell, g_sigma


def fit_derivs_dense(kind, x, y, J, ell, sigma):
    """_gpr.py:313-344 + gpr.py:475-476 (jaccoef)."""
    yf = y.reshape(-1, 1)
    m = yf.shape[0]
    A = d01kj(kind, x, x, ell, J, J) + (sigma**2 + 1e-10) * np.eye(m)
    c = np.linalg.solve(A, yf)
    ns, _, nv = J.shape
    jaccoef = np.einsum("sv,sfv->sf", c.reshape(ns, nv), J)
    return c, 0.0, jaccoef


def multivariate_normal(key, mean, cov, shape=(), partitionable=True):
    """jax.random.multivariate_normal, default method "cholesky" (third-party; restated from the published
    algorithm): mean + normal(key, shape + (n,)) @ cholesky(cov)^T."""
    L = np.linalg.cholesky(cov)
    z = normal(key, tuple(shape) + (mean.shape[-1],), partitionable)
    return mean + np.einsum("...ij,...j->...i", L, z)


def sample(key, mean, cov, n_samples=1, partitionable=True):
    """gpx/models/utils.py:52-75, quirks included: a 2-D mean is drawn column by column with the carried half of
    each split and only the LAST column's draw is returned; a 1-D mean ignores n_samples."""
    if mean.ndim > 1:
        out = None
        for dim in range(mean.shape[1]):
            key = split(key, 2, partitionable)[1]
            out = multivariate_normal(key, mean[:, dim], cov, (n_samples,), partitionable)
        return out
    return multivariate_normal(key, mean, cov, (), partitionable)


def sample_prior(kind, key, x, ell, sigma, n_samples=1, partitionable=True):
    """gpx/models/gpr.py:761-780."""
    return sample(key, np.zeros(x.shape), A_lhs(kind, x, x, ell, sigma), n_samples, partitionable)


def sample_posterior(kind, key, x_train, x, c, mu, ell, sigma, n_samples=1, partitionable=True):
    """gpx/models/gpr.py:809-832."""
    mean, cov = predict_dense(kind, x_train, x, c, mu, ell, sigma, full_covariance=True)
    return sample(key, mean, cov + 1e-10 * np.eye(cov.shape[0]), n_samples, partitionable)


# =============================================================================
# matrix-free mat-vec, Lanczos, SLQ, PCG
# (gpx/operations.py:32-123, gpx/lanczos.py:29-183, _gpr.py:94-119, 180-216, 285-310,
#  772-821; jax.scipy.sparse.linalg.cg restated)
# =============================================================================


def make_matvec(kind, x, ell, sigma, noise=True, block=2048):
    """(K + (sigma^2+1e-10) I) z without storing K; row blocks stand in for the
    reference's one-row-per-scan-step loop (operations.py:93-107)."""
    n = x.shape[0]
    shift = sigma**2 + 1e-10 if noise else 0.0

    def matvec(z):
        out = np.empty_like(z, dtype=np.float64)
        for i0 in range(0, n, block):
            i1 = min(n, i0 + block)
            rows = kernel(kind, x[i0:i1], x, ell)
            if noise:
                rows[np.arange(i1 - i0), np.arange(i0, i1)] += shift
            out[i0:i1] = rows @ z
        return out

    return matvec


def _orthogonalize(v, vecs):
    """lanczos.py:29-37 -- modified Gram-Schmidt in column order."""
    for j in range(vecs.shape[1]):
        v = v - np.dot(v, vecs[:, j]) * vecs[:, j]
    return v


def lanczos_tridiagonal(matvec, m, v1, key=None, partitionable=True):
    """lanczos.py:46-133."""
    n = v1.shape[0]
    vecs = np.zeros((n, m))
    vecs[:, 0] = v1
    T = np.zeros((m, m))
    w = matvec(v1)
    alpha = np.dot(w, v1)
    w = w - alpha * v1
    beta = np.linalg.norm(w)
    T[0, 0] = alpha
    for j in range(m - 1):
        if key is not None:
            sk = split(key, 2, partitionable)
            subkey, key = sk[0], sk[1]
        if beta > 1e-12:
            v = w / beta
        else:
            rv = normal(subkey, (n,), partitionable)
            rv = _orthogonalize(rv, vecs)
            v = rv / np.linalg.norm(rv)
        vecs[:, j + 1] = v
        w = matvec(v)
        alpha = np.dot(w, v)
        w = w - alpha * v - beta * vecs[:, j]
        w = _orthogonalize(w, vecs)
        T[j + 1, j + 1] = alpha
        T[j, j + 1] = beta
        T[j + 1, j] = beta
        beta = np.linalg.norm(w)
    return T, vecs


def slq_from_tridiag(T):
    """lanczos.py:175-180."""
    evals, evecs = np.linalg.eigh(T)
    evals = np.maximum(evals, 1e-36)
    return float(np.sum(np.square(evecs[0, :]) * np.log(evals)))


def rademacher_probes(key, n, num_evals, partitionable=True):
    """lanczos.py:168-170: returns (normalised probes (n, P), carried key)."""
    sk = split(key, 2, partitionable)
    subkey, key = sk[0], sk[1]
    us = rademacher(subkey, (n, num_evals), partitionable).astype(np.float64)
    us = us / np.linalg.norm(us, axis=0)
    return us, key


def lanczos_logdet(matvec, num_evals, dim_mat, num_lanczos, key, partitionable=True):
    """lanczos.py:136-183."""
    us, key = rademacher_probes(key, dim_mat, num_evals, partitionable)
    acc = 0.0
    for p in range(num_evals):
        T, _ = lanczos_tridiagonal(matvec, num_lanczos, us[:, p], key, partitionable)
        acc += slq_from_tridiag(T)
    return (dim_mat / num_evals) * acc


def rpcholesky(kind, x, n_pivots, ell, key, partitionable=True):
    """gpx/kernels/approximations.py:30-129."""
    n = x.shape[0]
    F = np.zeros((n, n_pivots))
    pivots = np.zeros(n_pivots, dtype=np.int64)
    # diagonal via the batched kernel on single points (approximations.py:88-93)
    z = x / ell
    zs = np.sum(z * z, axis=1)
    d2 = zs - 2.0 * zs + zs + 1e-12
    diag = _phi(kind, d2)
    key = np.asarray(key, dtype=np.uint32)
    for i in range(n_pivots):
        prob = diag / np.sum(diag)
        key = split(key, 2, partitionable)[0]  # key, subkey = split(key); sample with key
        s = choice_p(key, prob, partitionable)
        pivots[i] = s
        g = kernel(kind, x[s : s + 1], x, ell)[0] - F @ F[s]
        F[:, i] = g / (g[s] ** 0.5 + 1e-6)
        diag = diag - F[:, i] ** 2
    return F, pivots


def rpcholesky_composite(spec, x, n_pivots, key, partitionable=True):
    """gpx/kernels/approximations.py:30-129 with a Sum / Prod kernel (kernels.py:1794-1942): the reference calls the
    kernel object, whatever its class -- the diagonal through single-point evaluations (:88-93), each pivot's row through
    kernel(x[s], x)."""
    n = x.shape[0]
    F = np.zeros((n, n_pivots))
    pivots = np.zeros(n_pivots, dtype=np.int64)
    diag = np.array([composite_kernel(spec, x[i:i + 1], x[i:i + 1])[0, 0] for i in range(n)])
    key = np.asarray(key, dtype=np.uint32)
    for i in range(n_pivots):
        prob = diag / np.sum(diag)
        key = split(key, 2, partitionable)[0]
        s = choice_p(key, prob, partitionable)
        pivots[i] = s
        g = composite_kernel(spec, x[s:s + 1], x)[0] - F @ F[s]
        F[:, i] = g / (g[s] ** 0.5 + 1e-6)
        diag = diag - F[:, i] ** 2
    return F, pivots


def composite_fit_iter(spec, x, y, sigma, n_pivots, key_precond, mean_function=data_mean, partitionable=True,
                       cg_tol=1e-5, cg_atol=1e-7):
    """_gpr.py:285-310 with a composite kernel (dense matrix as the operator: the row scan is the same product)."""
    mu = mean_function(y)
    r = y - mu
    n = x.shape[0]
    A = composite_kernel(spec, x, x) + (sigma**2 + 1e-10) * np.eye(n)
    F, _ = rpcholesky_composite(spec, x, n_pivots, key_precond, partitionable)
    P = np.linalg.inv(sigma**2 * np.eye(n_pivots) + F.T @ F)

    def precond(z):
        return -(sigma ** (-2)) * (F @ (P @ (F.T @ z))) + sigma ** (-2) * z

    c, iters = cg(lambda v: A @ v, r, M=precond, tol=cg_tol, atol=cg_atol)
    return c, mu, iters


def composite_lml_iter(spec, x, y, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond,
                       mean_function=data_mean, partitionable=True, cg_tol=1e-5, cg_atol=1e-7):
    """_gpr.py:772-821 with a composite kernel."""
    m = y.shape[0]
    c, mu, _ = composite_fit_iter(spec, x, y, sigma, n_pivots, key_precond, mean_function, partitionable, cg_tol, cg_atol)
    r = y - mu
    A = composite_kernel(spec, x, x) + (sigma**2 + 1e-10) * np.eye(m)
    mll = -0.5 * np.sum(r.T @ c)
    mll -= 0.5 * lanczos_logdet(lambda v: A @ v, num_evals, m, num_lanczos, key_lanczos, partitionable)
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def _phi(kind, d2):
    if kind == "se":
        return np.exp(-d2)
    r = np.sqrt(np.maximum(d2, 1e-36))
    if kind == "m12":
        return np.exp(-r)
    if kind == "m32":
        d = np.sqrt(3.0) * r
        return (1.0 + d) * np.exp(-d)
    d = np.sqrt(5.0) * r
    return (1.0 + d + d**2 / 3.0) * np.exp(-d)


def precond_rpcholesky(kind, x, ell, sigma, n_pivots, key, partitionable=True):
    """_gpr.py:180-216 (explicit inverse, sigma^2 without the 1e-10)."""
    F, _ = rpcholesky(kind, x, n_pivots, ell, key, partitionable)
    P = np.linalg.inv(sigma**2 * np.eye(n_pivots) + F.T @ F)

    def precond(z):
        res = F.T @ z
        res = P @ res
        res = -(sigma ** (-2)) * (F @ res)
        return res + sigma ** (-2) * z

    return precond


def cg(A, b, M=None, tol=1e-5, atol=0.0, maxiter=None):
    """jax.scipy.sparse.linalg.cg (third-party; restated).
    Returns (x, iterations)."""
    if M is None:
        M = lambda v: v  # noqa: E731
        ident = True
    else:
        ident = False
    if maxiter is None:
        maxiter = 10 * b.size
    bs = float(np.vdot(b, b))
    atol2 = max(tol**2 * bs, atol**2)
    x = np.zeros_like(b)
    r = b - A(x)
    z = M(r)
    p = z
    gamma = float(np.vdot(r, z))
    k = 0
    while k < maxiter:
        rs = gamma if ident else float(np.vdot(r, r))
        if not rs > atol2:
            break
        Ap = A(p)
        alpha = gamma / float(np.vdot(p, Ap))
        x = x + alpha * p
        r = r - alpha * Ap
        z = M(r)
        gamma_new = float(np.vdot(r, z))
        p = z + (gamma_new / gamma) * p
        gamma = gamma_new
        k += 1
    return x, k


def lanczos_tridiagonal_jvp(matvec, dmatvecs, m, v1):
    """Forward-mode derivative of lanczos_tridiagonal (lanczos.py:46-133) w.r.t. parameters theta_q of the
    operator: every statement of the recursion is differentiated (v1 is constant; the breakdown branch of
    the jnp.where is not selected and carries no derivative).  dmatvecs[q](v) = (dA/d theta_q) v.
    Returns (T, [dT_q]).  This is the derivative jit(value_and_grad) propagates through the fori_loop,
    evaluated in tangent instead of cotangent form."""
    n, nq = v1.shape[0], len(dmatvecs)
    vecs = np.zeros((n, m))
    dvecs = [np.zeros((n, m)) for _ in range(nq)]
    vecs[:, 0] = v1
    T = np.zeros((m, m))
    dT = [np.zeros((m, m)) for _ in range(nq)]
    w = matvec(v1)
    dw = [dmv(v1) for dmv in dmatvecs]
    alpha = np.dot(w, v1)
    dalpha = [np.dot(dwq, v1) for dwq in dw]
    w = w - alpha * v1
    dw = [dwq - da * v1 for dwq, da in zip(dw, dalpha)]
    beta = np.linalg.norm(w)
    dbeta = [np.dot(w, dwq) / beta for dwq in dw]
    T[0, 0] = alpha
    for q in range(nq):
        dT[q][0, 0] = dalpha[q]
    for j in range(m - 1):
        if not beta > 1e-12:
            raise FloatingPointError("Lanczos breakdown: the tangent recursion is undefined")
        v = w / beta
        dv = [(dwq - v * db) / beta for dwq, db in zip(dw, dbeta)]
        vecs[:, j + 1] = v
        for q in range(nq):
            dvecs[q][:, j + 1] = dv[q]
        w = matvec(v)
        dw = [dmatvecs[q](v) + matvec(dv[q]) for q in range(nq)]
        alpha = np.dot(w, v)
        dalpha = [np.dot(dw[q], v) + np.dot(w, dv[q]) for q in range(nq)]
        dw = [dw[q] - dalpha[q] * v - alpha * dv[q] - dbeta[q] * vecs[:, j] - beta * dvecs[q][:, j] for q in range(nq)]
        w = w - alpha * v - beta * vecs[:, j]
        for k in range(m):  # _orthogonalize, column order; zero columns are no-ops
            coef = np.dot(w, vecs[:, k])
            dcoef = [np.dot(dw[q], vecs[:, k]) + np.dot(w, dvecs[q][:, k]) for q in range(nq)]
            dw = [dw[q] - dcoef[q] * vecs[:, k] - coef * dvecs[q][:, k] for q in range(nq)]
            w = w - coef * vecs[:, k]
        T[j + 1, j + 1] = alpha
        T[j, j + 1] = T[j + 1, j] = beta
        for q in range(nq):
            dT[q][j + 1, j + 1] = dalpha[q]
            dT[q][j, j + 1] = dT[q][j + 1, j] = dbeta[q]
        beta = np.linalg.norm(w)
        dbeta = [np.dot(w, dw[q]) / beta for q in range(nq)]
    return T, dT


def slq_from_tridiag_jvp(T, dT):
    """d/d theta of slq_from_tridiag (lanczos.py:175-180) = e1^T Dlog(T)[dT] e1: with T = Q diag(lam) Q^T and
    M = Q^T dT Q, sum_kl Q0k Q0l M_kl (log lam_k - log lam_l)/(lam_k - lam_l) (1/lam_k on the diagonal) -- the
    closed form of the eigh JVP applied to sum_k Q0k^2 log lam_k."""
    lam, Q = np.linalg.eigh(T)
    lam = np.maximum(lam, 1e-36)
    M = Q.T @ dT @ Q
    dl = lam[:, None] - lam[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        dd = np.where(np.abs(dl) > 1e-14 * np.abs(lam[:, None]), (np.log(lam)[:, None] - np.log(lam)[None, :]) / dl,
                      1.0 / lam[:, None])
    q0 = Q[0, :]
    return float(q0 @ (M * dd) @ q0)


def lanczos_logdet_grad(kind, x, ell, sigma, num_evals, num_lanczos, key, partitionable=True):
    """(logdet_SLQ, d/d ell, d/d sigma) of lanczos_logdet applied to A = K + (sigma^2+1e-10) I."""
    n = x.shape[0]
    K, dK = kernel_dl(kind, x, x, ell)
    A = K + (sigma**2 + 1e-10) * np.eye(n)
    us, _ = rademacher_probes(key, n, num_evals, partitionable)
    acc, dacc = 0.0, np.zeros(2)
    for p in range(num_evals):
        T, dT = lanczos_tridiagonal_jvp(lambda v: A @ v, [lambda v: dK @ v, lambda v: 2.0 * sigma * v], num_lanczos,
                                        us[:, p])
        acc += slq_from_tridiag(T)
        dacc += np.array([slq_from_tridiag_jvp(T, dT[0]), slq_from_tridiag_jvp(T, dT[1])])
    return (n / num_evals) * acc, (n / num_evals) * dacc[0], (n / num_evals) * dacc[1]


def lml_iter_grad(kind, x, y, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond,
                  mean_function=data_mean, partitionable=True, cg_tol=1e-5, cg_atol=1e-7):
    """(lml, d/d ell, d/d sigma) of _lml_iter (_gpr.py:772-821) as jit(value_and_grad) propagates it:
    the CG solve is a lax.custom_linear_solve, so d(r.c) = -c^T dA c with c the computed PCG solution
    (nothing flows through the preconditioner or the iteration count); the SLQ term is differentiated
    through the Lanczos recursion and the eigendecomposition."""
    m = y.shape[0]
    c, mu, _ = fit_iter(kind, x, y, ell, sigma, n_pivots, key_precond, mean_function, partitionable, cg_tol, cg_atol)
    r = y - mu
    ld, ld_l, ld_s = lanczos_logdet_grad(kind, x, ell, sigma, num_evals, num_lanczos, key_lanczos, partitionable)
    _, dK = kernel_dl(kind, x, x, ell)
    mll = (-0.5 * np.sum(r.T @ c) - 0.5 * ld - m * 0.5 * np.log(2.0 * np.pi)) / m
    g_l = (0.5 * np.sum(c * (dK @ c)) - 0.5 * ld_l) / m
    g_s = (0.5 * 2.0 * sigma * np.sum(c * c) - 0.5 * ld_s) / m
    return mll, g_l, g_s


def rpcholesky_derivs(kind, x, J, n_pivots, ell, key, partitionable=True):
    """gpx/kernels/approximations.py:132-242: RPCholesky of the Hessian kernel over the (sample, variable) rows."""
    n, _ = x.shape
    jv = J.shape[2]
    F = np.zeros((n * jv, n_pivots))
    pivots = np.zeros(n_pivots, dtype=np.int64)
    diag = np.concatenate([np.diagonal(d01kj(kind, x[i:i + 1], x[i:i + 1], ell, J[i:i + 1], J[i:i + 1]))
                           for i in range(n)])
    key = np.asarray(key, dtype=np.uint32)
    for i in range(n_pivots):
        prob = diag / np.sum(diag)
        key = split(key, 2, partitionable)[0]
        s = choice_p(key, prob, partitionable)
        pivots[i] = s
        i1, i2 = int(s) // jv, int(s) % jv
        g = d01kj(kind, x[i1:i1 + 1], x, ell, J[i1:i1 + 1], J)[i2] - F @ F[s]
        F[:, i] = g / (g[s] ** 0.5 + 1e-6)
        diag = diag - F[:, i] ** 2
    return F, pivots


def fit_derivs_iter(kind, x, y, J, ell, sigma, n_pivots, key_precond, partitionable=True):
    """_fit_derivs_iter _gpr.py:347-390 with _precond_rpcholesky_derivs :219-256 (zero mean)."""
    yf = y.reshape(-1, 1)
    m = yf.shape[0]
    A = d01kj(kind, x, x, ell, J, J) + (sigma**2 + 1e-10) * np.eye(m)
    F, _ = rpcholesky_derivs(kind, x, J, n_pivots, ell, key_precond, partitionable)
    P = np.linalg.inv(sigma**2 * np.eye(n_pivots) + F.T @ F)

    def precond(z):
        res = P @ (F.T @ z)
        return -(sigma ** (-2)) * (F @ res) + sigma ** (-2) * z

    c, iters = cg(lambda v: A @ v, yf, M=precond, atol=1e-7)
    return c, iters


def lml_derivs_iter(kind, x, y, J, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond,
                    partitionable=True):
    """_lml_derivs_iter _gpr.py:873-935."""
    yf = y.reshape(-1, 1)
    m = yf.shape[0]
    c, _ = fit_derivs_iter(kind, x, y, J, ell, sigma, n_pivots, key_precond, partitionable)
    A = d01kj(kind, x, x, ell, J, J) + (sigma**2 + 1e-10) * np.eye(m)
    mll = -0.5 * np.sum(yf.T @ c)
    mll -= 0.5 * lanczos_logdet(lambda v: A @ v, num_evals, m, num_lanczos, key_lanczos, partitionable)
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


def lml_derivs_iter_grad(kind, x, y, J, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond,
                         partitionable=True):
    """(lml, d/d ell, d/d sigma) of _lml_derivs_iter as value_and_grad propagates it (see lml_iter_grad)."""
    yf = y.reshape(-1, 1)
    m = yf.shape[0]
    c, _ = fit_derivs_iter(kind, x, y, J, ell, sigma, n_pivots, key_precond, partitionable)
    A = d01kj(kind, x, x, ell, J, J) + (sigma**2 + 1e-10) * np.eye(m)
    dA = d01kj_dl(kind, x, x, ell, J, J)
    us, _ = rademacher_probes(key_lanczos, m, num_evals, partitionable)
    acc, dacc = 0.0, np.zeros(2)
    for p in range(num_evals):
        T, dT = lanczos_tridiagonal_jvp(lambda v: A @ v, [lambda v: dA @ v, lambda v: 2.0 * sigma * v], num_lanczos,
                                        us[:, p])
        acc += slq_from_tridiag(T)
        dacc += np.array([slq_from_tridiag_jvp(T, dT[0]), slq_from_tridiag_jvp(T, dT[1])])
    ld, dld = (m / num_evals) * acc, (m / num_evals) * dacc
    mll = (-0.5 * np.sum(yf.T @ c) - 0.5 * ld - m * 0.5 * np.log(2.0 * np.pi)) / m
    g_l = (0.5 * np.sum(c * (dA @ c)) - 0.5 * dld[0]) / m
    g_s = (sigma * np.sum(c * c) - 0.5 * dld[1]) / m
    return mll, g_l, g_s


def fit_iter(kind, x, y, ell, sigma, n_pivots, key_precond, mean_function=data_mean, partitionable=True,
             cg_tol=1e-5, cg_atol=1e-7):
    """_gpr.py:285-310.  cg_tol / cg_atol default to the reference's stopping rule (jax cg: tol=1e-5; atol=1e-7 at the
    call site); the tests tighten them on BOTH sides to compare converged solves at 1e-8 (a test knob, not in GPX)."""
    mu = mean_function(y)
    r = y - mu
    mv = make_matvec(kind, x, ell, sigma)
    pre = precond_rpcholesky(kind, x, ell, sigma, n_pivots, key_precond, partitionable)
    c, iters = cg(mv, r, M=pre, tol=cg_tol, atol=cg_atol)
    return c, mu, iters


def lml_iter(kind, x, y, ell, sigma, num_evals, num_lanczos, key_lanczos, n_pivots, key_precond,
             mean_function=data_mean, partitionable=True, cg_tol=1e-5, cg_atol=1e-7):
    """_gpr.py:772-821 (cg_tol / cg_atol: see fit_iter)."""
    m = y.shape[0]
    c, mu, _ = fit_iter(kind, x, y, ell, sigma, n_pivots, key_precond, mean_function, partitionable, cg_tol, cg_atol)
    r = y - mu
    mv = make_matvec(kind, x, ell, sigma)
    mll = -0.5 * np.sum(r.T @ c)
    mll -= 0.5 * lanczos_logdet(lambda v: mv(v), num_evals, m, num_lanczos, key_lanczos, partitionable)
    mll -= m * 0.5 * np.log(2.0 * np.pi)
    return mll / m


# =============================================================================
# SGPR, Projected Processes (gpx/models/_sgpr.py, _sgpr_rpchol.py)
# =============================================================================


def sgpr_A_lhs(kind, x_locs, x, ell, sigma):
    """_sgpr.py:39-61."""
    m = x_locs.shape[0]
    K_mn = kernel(kind, x_locs, x, ell)
    C = sigma**2 * kernel(kind, x_locs, x_locs, ell) + K_mn @ K_mn.T + 1e-10 * np.eye(m)
    return C, K_mn


def sgpr_fit_dense(kind, x_locs, x, y, ell, sigma, mean_function=data_mean):
    """_sgpr.py:319-339."""
    mu = mean_function(y)
    C, K_mn = sgpr_A_lhs(kind, x_locs, x, ell, sigma)
    c = np.linalg.solve(C, K_mn @ (y - mu))
    return c, mu


def sgpr_predict_dense(kind, x_locs, x, c, mu, ell, sigma, full_covariance=False):
    """_sgpr.py:434-467."""
    m = x_locs.shape[0]
    K_nm = kernel(kind, x, x_locs, ell)
    mean = mu + K_nm @ c
    if not full_covariance:
        return mean
    C, K_mn = sgpr_A_lhs(kind, x_locs, x, ell, sigma)
    K_mm = kernel(kind, x_locs, x_locs, ell) + 1e-10 * np.eye(m)
    L = sla.cholesky(K_mm, lower=True, check_finite=False)
    G = sla.solve_triangular(L, K_mn, lower=True, check_finite=False)
    L2 = sla.cholesky(C, lower=True, check_finite=False)
    H = sla.solve_triangular(L2, K_mn, lower=True, check_finite=False)
    C_nn = kernel(kind, x, x, ell) - G.T @ G + sigma**2 * (H.T @ H)
    return mean, C_nn


def sgpr_lml_dense_nxn(kind, x_locs, x, y, ell, sigma, mean_function=data_mean):
    """_sgpr.py:659-697, literal N x N form (feasible for small N only)."""
    m = x_locs.shape[0]
    mu = mean_function(y)
    r = y - mu
    n = y.shape[0]
    K_mm = kernel(kind, x_locs, x_locs, ell)
    K_mn = kernel(kind, x_locs, x, ell)
    L = sla.cholesky(K_mm + 1e-10 * np.eye(m), lower=True, check_finite=False)
    G = sla.solve_triangular(L, K_mn, lower=True, check_finite=False)
    C = G.T @ G + sigma**2 * np.eye(n) + 1e-10 * np.eye(n)
    Ln = sla.cholesky(C, lower=True, check_finite=False)
    cy = sla.solve_triangular(Ln, r, lower=True, check_finite=False)
    mll = -0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(Ln))) - n * 0.5 * np.log(2.0 * np.pi)
    return mll / n


def sgpr_lml_dense(kind, x_locs, x, y, ell, sigma, mean_function=data_mean):
    """Woodbury / matrix-determinant-lemma form of _sgpr.py:659-697;
    algebraically identical, O(N M^2)."""
    m = x_locs.shape[0]
    mu = mean_function(y)
    r = y - mu
    n = y.shape[0]
    s2 = sigma**2 + 1e-10
    K_mm = kernel(kind, x_locs, x_locs, ell)
    K_mn = kernel(kind, x_locs, x, ell)
    L = sla.cholesky(K_mm + 1e-10 * np.eye(m), lower=True, check_finite=False)
    G = sla.solve_triangular(L, K_mn, lower=True, check_finite=False)
    B = s2 * np.eye(m) + G @ G.T
    LB = sla.cholesky(B, lower=True, check_finite=False)
    t = sla.solve_triangular(LB, G @ r, lower=True, check_finite=False)
    quad = (np.sum(r * r) - np.sum(t * t)) / s2
    logdet = (n - m) * np.log(s2) + 2.0 * np.sum(np.log(np.diag(LB)))
    return (-0.5 * quad - 0.5 * logdet - n * 0.5 * np.log(2.0 * np.pi)) / n


def sgpr_composite_lml_dense_nxn(spec, x_locs, x, y, sigma, mean_function=data_mean):
    """_sgpr.py:659-697, literal N x N form, with a composite kernel (composite_kernel spec)."""
    m, n = x_locs.shape[0], y.shape[0]
    r = y - mean_function(y)
    K_mm = composite_kernel(spec, x_locs, x_locs)
    K_mn = composite_kernel(spec, x_locs, x)
    L = sla.cholesky(K_mm + 1e-10 * np.eye(m), lower=True, check_finite=False)
    G = sla.solve_triangular(L, K_mn, lower=True, check_finite=False)
    C = G.T @ G + sigma**2 * np.eye(n) + 1e-10 * np.eye(n)
    Ln = sla.cholesky(C, lower=True, check_finite=False)
    cy = sla.solve_triangular(Ln, r, lower=True, check_finite=False)
    return (-0.5 * np.sum(cy**2) - np.sum(np.log(np.diag(Ln))) - n * 0.5 * np.log(2.0 * np.pi)) / n


def sgpr_composite_fit_dense(spec, x_locs, x, y, sigma, mean_function=data_mean):
    """_sgpr.py:39-61, 319-339 with a composite kernel."""
    mu = mean_function(y)
    K_mn = composite_kernel(spec, x_locs, x)
    C = sigma**2 * composite_kernel(spec, x_locs, x_locs) + K_mn @ K_mn.T + 1e-10 * np.eye(x_locs.shape[0])
    return np.linalg.solve(C, K_mn @ (y - mu)), mu


def sgpr_composite_predict_dense(spec, x_locs, x, c, mu, sigma):
    """_sgpr.py:434-467 with a composite kernel, full covariance (C_mm from the prediction inputs, as written
    there: `_A_lhs(x1=x_locs, x2=x)`)."""
    m = x_locs.shape[0]
    mean = mu + composite_kernel(spec, x, x_locs) @ c
    K_ms = composite_kernel(spec, x_locs, x)
    K_mm = composite_kernel(spec, x_locs, x_locs)
    C = sigma**2 * K_mm + K_ms @ K_ms.T + 1e-10 * np.eye(m)
    G = sla.solve_triangular(sla.cholesky(K_mm + 1e-10 * np.eye(m), lower=True), K_ms, lower=True)
    H = sla.solve_triangular(sla.cholesky(C, lower=True), K_ms, lower=True)
    return mean, composite_kernel(spec, x, x) - G.T @ G + sigma**2 * (H.T @ H)


def sgpr_fit_derivs_dense(kind, x_locs, J_locs, x, y, J, ell, sigma):
    """_sgpr.py:366-396 with _A_derivs_lhs :65-93 (zero mean) + the contracted jacobian of sgpr.py:372-375."""
    yf = y.reshape(-1, 1)
    K_mn = d01kj(kind, x_locs, x, ell, J_locs, J)
    C = sigma**2 * d01kj(kind, x_locs, x_locs, ell, J_locs, J_locs) + K_mn @ K_mn.T + 1e-10 * np.eye(K_mn.shape[0])
    c = np.linalg.solve(C, K_mn @ yf)
    ns, _, nv = J_locs.shape
    return c, np.einsum("sv,sfv->sf", c.reshape(ns, nv), J_locs)


def sgpr_fit_iter(kind, x_locs, x, y, ell, sigma, mean_function=data_mean):
    """_sgpr.py:343-363 with _Ax_lhs_fun :96-143: cg(sigma^2 K_mm + K_mn K_nm + 1e-10 I, K_mn (y - mu)), default
    tolerances (tol 1e-5, atol 0), no preconditioner."""
    mu = mean_function(y)
    K_mm = kernel(kind, x_locs, x_locs, ell)
    K_mn = kernel(kind, x_locs, x, ell)
    m = x_locs.shape[0]
    A = sigma**2 * K_mm + K_mn @ K_mn.T + 1e-10 * np.eye(m)
    c, iters = cg(lambda v: A @ v, K_mn @ (y - mu))
    return c, mu, iters


def sgpr_lml_iter(kind, x_locs, x, y, ell, sigma, num_evals, num_lanczos, key_lanczos, mean_function=data_mean,
                  partitionable=True):
    """_sgpr.py:701-737 with _Hx_fun :204-239 (explicit inverse of K_mm + 1e-10 I)."""
    mu = mean_function(y)
    r = y - mu
    n, m = y.shape[0], x_locs.shape[0]
    Kmm_inv = np.linalg.inv(kernel(kind, x_locs, x_locs, ell) + 1e-10 * np.eye(m))
    K_nm = kernel(kind, x, x_locs, ell)
    shift = sigma**2 + 1e-10

    def mv(v):
        return K_nm @ (Kmm_inv @ (K_nm.T @ v)) + shift * v

    c, _ = cg(mv, r)
    mll = -0.5 * np.sum(r.T @ c)
    mll -= 0.5 * lanczos_logdet(mv, num_evals, n, num_lanczos, key_lanczos, partitionable)
    mll -= n * 0.5 * np.log(2.0 * np.pi)
    return mll / n


def sgpr_lml_iter_grad(kind, x_locs, x, y, ell, sigma, num_evals, num_lanczos, key_lanczos, mean_function=data_mean,
                       partitionable=True, cg_tol=1e-5):
    """(lml, d/d ell, d/d sigma) of sgpr_lml_iter as jit(value_and_grad) propagates it (_sgpr.py:700-737; landmarks
    fixed): the CG solve is a custom_linear_solve, d(r.c) = -c^T dA c; the SLQ term is differentiated through the
    Lanczos recursion.  A = K_nm (K_mm + 1e-10 I)^-1 K_mn + (sigma^2 + 1e-10) I,
    dA/d ell = dK_nm W K_mn + K_nm W dK_mn - K_nm W dK_mm W K_mn (W the explicit inverse), dA/d sigma = 2 sigma I."""
    mu = mean_function(y)
    r = y - mu
    n, m = y.shape[0], x_locs.shape[0]
    K_mm, dK_mm = kernel_dl(kind, x_locs, x_locs, ell)
    np.fill_diagonal(dK_mm, 0.0)  # the diagonal distance is the constant jitter
    W = np.linalg.inv(K_mm + 1e-10 * np.eye(m))
    K_nm, dK_nm = kernel_dl(kind, x, x_locs, ell)
    shift = sigma**2 + 1e-10

    def mv(v):
        return K_nm @ (W @ (K_nm.T @ v)) + shift * v

    def dmv(v):
        g = W @ (K_nm.T @ v)
        return dK_nm @ g + K_nm @ (W @ (dK_nm.T @ v - dK_mm @ g))

    c, _ = cg(mv, r, tol=cg_tol)
    us, _ = rademacher_probes(key_lanczos, n, num_evals, partitionable)
    acc, dacc = 0.0, np.zeros(2)
    for p in range(num_evals):
        T, dT = lanczos_tridiagonal_jvp(mv, [dmv, lambda v: 2.0 * sigma * v], num_lanczos, us[:, p])
        acc += slq_from_tridiag(T)
        dacc += np.array([slq_from_tridiag_jvp(T, dT[0]), slq_from_tridiag_jvp(T, dT[1])])
    ld, dld = (n / num_evals) * acc, (n / num_evals) * dacc
    mll = (-0.5 * np.sum(r.T @ c) - 0.5 * ld - n * 0.5 * np.log(2.0 * np.pi)) / n
    g_l = (0.5 * np.sum(c * dmv(c)) - 0.5 * dld[0]) / n
    g_s = (sigma * np.sum(c * c) - 0.5 * dld[1]) / n
    return mll, g_l, g_s


def kernel_dx_profile(kind: str, x1, x2, ell: float):
    """h with d kernel(x1_i, x2_j)/d x1_if = -h_ij (x1_if - x2_jf): h = -(2/ell^2) dk/d(d2) of the matrix
    kernels above (d2 carries the 1e-12 jitter of squared_distances; the Matern clamp is inactive)."""
    d2 = squared_distances(x1 / ell, x2 / ell)
    if kind == "se":
        return (2.0 / ell**2) * np.exp(-d2)
    r = np.sqrt(np.maximum(d2, 1e-36))
    if kind == "m12":
        return np.exp(-r) / (ell**2 * r)
    if kind == "m32":
        return (3.0 / ell**2) * np.exp(-np.sqrt(3.0) * r)
    if kind == "m52":
        d = np.sqrt(5.0) * r
        return (5.0 / (3.0 * ell**2)) * (1.0 + d) * np.exp(-d)
    raise ValueError(kind)


def sgpr_lml_grad_locs_nxn(kind, x_locs, x, y, ell, sigma, mean_function=data_mean):
    """d(lml/n)/d x_locs of _sgpr.py:659-697 (what value_and_grad yields for a trainable `x_locs` Parameter,
    sgpr.py:690-732) from the literal N x N form: C = K_nm A^-1 K_mn + s2 I, S = beta beta^T - C^-1,
    d(lml n) = tr(S K_nm A^-1 dK_mn) - 1/2 tr(A^-1 K_mn S K_nm A^-1 dA)."""
    m, n = x_locs.shape[0], y.shape[0]
    r = y - mean_function(y)
    A = kernel(kind, x_locs, x_locs, ell) + 1e-10 * np.eye(m)
    K_mn = kernel(kind, x_locs, x, ell)
    P = np.linalg.solve(A, K_mn)
    C = K_mn.T @ P + (sigma**2 + 1e-10) * np.eye(n)
    Cinv = np.linalg.inv(C)
    beta = Cinv @ r
    S = beta @ beta.T - Cinv
    T = P @ S                      # (m, n)
    U = T @ P.T                    # (m, m)
    h_mn = kernel_dx_profile(kind, x_locs, x, ell)
    h_mm = kernel_dx_profile(kind, x_locs, x_locs, ell)
    dz_mn = x_locs[:, None, :] - x[None, :, :]
    dz_mm = x_locs[:, None, :] - x_locs[None, :, :]
    g = -np.einsum("mn,mn,mnf->mf", T, h_mn, dz_mn) + np.einsum("ab,ab,abf->af", U, h_mm, dz_mm)
    return g / n


def sgpr_rpchol_lml_dense(kind, x, y, ell, sigma, key, n_locs, mean_function=data_mean,
                          partitionable=True, nxn=False):
    """_sgpr_rpchol.py:178-208."""
    _, pivots = rpcholesky(kind, x, n_locs, ell, key, partitionable)
    x_locs = x[pivots].copy()
    f = sgpr_lml_dense_nxn if nxn else sgpr_lml_dense
    return f(kind, x_locs, x, y, ell, sigma, mean_function), pivots


def sgpr_rpchol_fit_dense(kind, x, y, ell, sigma, key, n_locs, mean_function=data_mean, partitionable=True):
    """_sgpr_rpchol.py:36-67."""
    _, pivots = rpcholesky(kind, x, n_locs, ell, key, partitionable)
    x_locs = x[pivots].copy()
    c, mu = sgpr_fit_dense(kind, x_locs, x, y, ell, sigma, mean_function)
    return c, mu, x_locs
