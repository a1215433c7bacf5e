"""CPU stand-ins for the device calls of gpx_b200.ops, for HOST-LOGIC tests only (tests/test_host_glue_cpu.py).

The facade modules that are compositions of device calls (gpx_b200/models/_composite.py, ...) contain index
bookkeeping, adjoint algebra and parameter plumbing that can be wrong independently of any kernel.  `patched()`
swaps the handful of ops those modules use for NumPy restatements with the SAME contracts (in-place updates,
lower-triangle-only reads and writes, argument meaning) so that glue can be exercised by `-m "not gpu"` tests.
This is test infrastructure: nothing under gpx_b200/ imports it, and the GPU tests run the same glue on the
real kernels."""
from __future__ import annotations

import contextlib

import numpy as np
import torch

from oracle import ref_numpy as R

def _np(t):
    return t.detach().cpu().numpy()


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))


def to_device(a):
    if isinstance(a, torch.Tensor):
        return a.to(dtype=torch.float64).contiguous()
    return _t(a)


def gram(kind, x1, x2, ell, diag_add=0.0, lower_only=False, out=None):
    sym = x2 is x1 or x2 is None
    a = _np(x1)
    b = a if sym else _np(x2)
    K = R.kernel(kind, a, b, float(ell))
    if sym:
        K = K.copy()
        K[np.diag_indices_from(K)] = R.kernel(kind, a[:1], a[:1], float(ell))[0, 0]  # exact diagonal
    if diag_add:
        K = K + diag_add * np.eye(K.shape[0], K.shape[1])
    if out is None:
        out = torch.zeros(K.shape, dtype=torch.float64)
    if lower_only:
        assert sym
        o = _np(out)
        il = np.tril_indices(K.shape[0])
        o[il] = K[il]
        out.copy_(_t(o))
    else:
        out.copy_(_t(K))
    return out


def gram_composite_fits(leaf_dims):
    return True


def gram_composite(kinds, ells, x1_leaves, x2_leaves, prog, diag_add=0.0, lower_only=False, out=None):
    """NumPy stand-in of the fused composite Gram kernel: leaf matrices by the stand-in `gram` (Linear: a dot product),
    combined by the postfix program."""
    sym = x2_leaves is None
    stack = []
    for t in prog:
        if t >= 0:
            a = x1_leaves[t]
            b = a if sym else x2_leaves[t]
            if kinds[t] == "linear":
                stack.append(_np(a) @ _np(b).T)
            else:
                stack.append(_np(gram(kinds[t], a, b, ells[t])))
        else:
            y, x = stack.pop(), stack.pop()
            stack.append(x + y if t == -1 else x * y)
    K = stack[0]
    if diag_add:
        K = K + diag_add * np.eye(K.shape[0], K.shape[1])
    if lower_only:
        K = np.tril(K)
    res = _t(K)
    if out is not None:
        out.copy_(res)
        return out
    return res


def gram_dl_contract(kind, x1, x2, ell, Y, sym_lower=False):
    dK = R.kernel_dl(kind, _np(x1), _np(x2), float(ell))[1]
    Yn = _np(Y)
    if sym_lower:
        Yl = np.tril(Yn)
        Yn = Yl + np.tril(Yn, -1).T
    return _t(np.array([np.sum(Yn * dK)]))


def elementwise(op, X, Y, a=1.0, b=1.0, out=None):
    assert X.is_contiguous() and Y.is_contiguous() and X.shape == Y.shape
    res = a * X + b * Y if op == "axpby" else a * X * Y
    if out is None:
        return res.clone()
    out.copy_(res)
    return out


def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None, lower_only=False):
    a = _np(A).T if transa else _np(A)
    b = _np(B).T if transb else _np(B)
    prod = a @ b
    if C is None:
        C = torch.zeros(prod.shape, dtype=torch.float64)
    c = _np(C)
    new = alpha * prod + beta * c
    if lower_only:
        il = np.tril_indices(new.shape[0], 0, new.shape[1])
        c = c.copy()
        c[il] = new[il]
        new = c
    C.copy_(_t(new))
    return C


def potrf(A):
    a = _np(A)
    low = np.tril(a)
    sym = low + np.tril(a, -1).T
    L = np.linalg.cholesky(sym)
    out = np.triu(a, 1) + L  # the strict upper triangle is never touched
    A.copy_(_t(out))
    return A, torch.zeros(1, dtype=torch.float64)


def trsm(L, inv, B, trans=True):
    import scipy.linalg as sla

    Ln = np.tril(_np(L))
    b = _np(B)
    # B <- B L^-T  <=>  B^T <- L^-1 B^T ;   B <- B L^-1  <=>  B^T <- L^-T B^T
    x = sla.solve_triangular(Ln, b.T, lower=True, trans="N" if trans else "T")
    B.copy_(_t(x.T))
    return B


def logdet(L):
    return _t(np.array([np.sum(np.log(np.diag(_np(L))))]))


def potri(L, inv, out=None):
    Ln = np.tril(_np(L))
    Li = np.linalg.inv(Ln)
    Ainv = np.tril(Li.T @ Li)
    if out is None:
        out = torch.zeros(Ainv.shape, dtype=torch.float64)
    o = _np(out).copy()
    il = np.tril_indices(Ainv.shape[0])
    o[il] = Ainv[il]
    out.copy_(_t(o))
    return out


# ---- fused entry points (csrc/gpr.cu, sgpr.cu, rpchol.cu, matvec.cu): the oracle's restatement of each ----
def _mean_fn(mean_mode):
    return R.data_mean if int(mean_mode) == 1 else R.zero_mean


def _col(y):
    y = _np(y)
    return y.reshape(-1, 1) if y.ndim == 1 else y


def gpr_lml_grad(kind, x, y, ell, sigma, mean_mode=1, want_grad=True):
    x, y = _np(x), _col(y)
    if not want_grad:
        return _t(np.array([R.lml_dense(kind, x, y, float(ell), float(sigma), _mean_fn(mean_mode)), 0.0, 0.0]))
    return _t(np.array(R.lml_grad_dense(kind, x, y, float(ell), float(sigma), _mean_fn(mean_mode))))


def gpr_fit(kind, x, y, ell, sigma, mean_mode=1):
    c, mu = R.fit_dense(kind, _np(x), _col(y), float(ell), float(sigma), _mean_fn(mean_mode))
    return _t(c), _t(np.array([float(mu)]))


def gpr_predict(kind, x_train, x, c, mu, ell, sigma, full_covariance=False):
    out = R.predict_dense(kind, _np(x_train), _np(x), _col(c), float(_np(mu)[0]), float(ell), float(sigma),
                          full_covariance=full_covariance)
    return (_t(out[0]), _t(out[1])) if full_covariance else _t(out)


def kmatvec(kind, x_rows, x_all, V, ell, diag_add=0.0, row_offset=None):
    K = R.kernel(kind, _np(x_rows), _np(x_all), float(ell))
    if diag_add:
        off = int(row_offset or 0)
        for i in range(K.shape[0]):
            K[i, off + i] += diag_add
    v = _np(V)
    return _t(K @ (v.reshape(-1, 1) if v.ndim == 1 else v))


def rpcholesky(kind, x, n_pivots, ell, key, partitionable=True, want_factor=True):
    F, piv = R.rpcholesky(kind, _np(x), int(n_pivots), float(ell), np.asarray(key, dtype=np.uint32), bool(partitionable))
    return (_t(F) if want_factor else None), torch.as_tensor(np.asarray(piv, dtype=np.int64))


def gather_rows(x, idx):
    return x[idx.to(torch.long)].contiguous()


def sgpr_lml(kind, x, y, x_locs, ell, sigma, mean_mode=1):
    return _t(np.array([R.sgpr_lml_dense(kind, _np(x_locs), _np(x), _col(y), float(ell), float(sigma), _mean_fn(mean_mode))]))


def sgpr_lml_grad(kind, x, y, x_locs, ell, sigma, mean_mode=1, want_grad=True):
    xl, xn, yn, mf = _np(x_locs), _np(x), _col(y), _mean_fn(mean_mode)
    f = lambda e, s: R.sgpr_lml_dense(kind, xl, xn, yn, e, s, mf)  # noqa: E731
    ell, sigma = float(ell), float(sigma)
    if not want_grad:
        return _t(np.array([f(ell, sigma), 0.0, 0.0]))

    def rich(g, h):  # Richardson-extrapolated central difference, O(h^4)
        fd = lambda hh: (g(hh) - g(-hh)) / (2 * hh)  # noqa: E731
        return (4 * fd(h / 2) - fd(h)) / 3

    return _t(np.array([f(ell, sigma), rich(lambda h: f(ell + h, sigma), 1e-3 * ell),
                        rich(lambda h: f(ell, sigma + h), 1e-3 * sigma)]))


def sgpr_fit(kind, x, y, x_locs, ell, sigma, mean_mode=1):
    c, mu = R.sgpr_fit_dense(kind, _np(x_locs), _np(x), _col(y), float(ell), float(sigma), _mean_fn(mean_mode))
    return _t(c), _t(np.array([float(mu)]))


def sgpr_predict(kind, x_locs, x, c, mu, ell, sigma, full_covariance=False):
    out = R.sgpr_predict_dense(kind, _np(x_locs), _np(x), _col(c), float(_np(mu)[0]), float(ell), float(sigma),
                               full_covariance=full_covariance)
    return (_t(out[0]), _t(out[1])) if full_covariance else _t(out)


# ---- GPR on targets and derivatives (csrc/td.cu).  Library row order: [targets; (sample, variable)] ----
def td_gram(kind, x1, J1, x2, J2, ell, add_targets=0.0, add_derivs=0.0, lower_only=False):
    assert not lower_only
    A = R.td_A_lhs(kind, _np(x1), _np(J1), _np(x2), _np(J2), float(ell), 0.0, 0.0, noise=False)
    n1 = x1.shape[0]
    if add_targets or add_derivs:
        A[np.arange(n1), np.arange(n1)] += add_targets
        i = np.arange(n1, A.shape[0])
        A[i, i] += add_derivs
    return _t(A)


def _rich(g, h):
    fd = lambda hh: (g(hh) - g(-hh)) / (2 * hh)  # noqa: E731
    return (4 * fd(h / 2) - fd(h)) / 3


def gpr_td_lml_grad(kind, x, J, y, y_derivs, ell, sigma_targets, sigma_derivs, mean_mode=1, want_grad=True):
    xn, Jn, yn, ydn, mf = _np(x), _np(J), _col(y), _np(y_derivs), _mean_fn(mean_mode)
    f = lambda e, a, b: R.td_lml_dense(kind, xn, yn, Jn, ydn, e, a, b, mf)  # noqa: E731
    e, a, b = float(ell), float(sigma_targets), float(sigma_derivs)
    if not want_grad:
        return _t(np.array([f(e, a, b), 0.0, 0.0, 0.0]))
    return _t(np.array([f(e, a, b), _rich(lambda h: f(e + h, a, b), 1e-3 * e), _rich(lambda h: f(e, a + h, b), 1e-3 * a),
                        _rich(lambda h: f(e, a, b + h), 1e-3 * b)]))


def gpr_td_fit(kind, x, J, y, y_derivs, ell, sigma_targets, sigma_derivs, mean_mode=1):
    c, mu = R.td_fit_dense(kind, _np(x), _col(y), _np(J), _np(y_derivs), float(ell), float(sigma_targets),
                           float(sigma_derivs), _mean_fn(mean_mode))
    return _t(c), _t(np.array([float(mu)]))


def _td2_split(J, y_derivs, nv_first):
    Jn, ydn = _np(J), _np(y_derivs).reshape(J.shape[0], J.shape[2])
    return Jn[:, :, :nv_first], Jn[:, :, nv_first:], ydn[:, :nv_first], ydn[:, nv_first:]


def gpr_td2_lml_grad(kind, x, J, y, y_derivs, nv_first, ell, sigma_targets, sigma_derivs, sigma_derivs2, mean_mode=1,
                     want_grad=True):
    Ja, Jb, ya, yb = _td2_split(J, y_derivs, nv_first)
    xn, yn, mf = _np(x), _col(y), _mean_fn(mean_mode)
    f = lambda e, a, b, c: R.td2_lml_dense(kind, xn, yn, Ja, Jb, ya, yb, e, a, b, c, mf)  # noqa: E731
    e, a, b, c = float(ell), float(sigma_targets), float(sigma_derivs), float(sigma_derivs2)
    if not want_grad:
        return _t(np.array([f(e, a, b, c), 0.0, 0.0, 0.0, 0.0]))
    return _t(np.array([f(e, a, b, c), _rich(lambda h: f(e + h, a, b, c), 1e-3 * e),
                        _rich(lambda h: f(e, a + h, b, c), 1e-3 * a), _rich(lambda h: f(e, a, b + h, c), 1e-3 * b),
                        _rich(lambda h: f(e, a, b, c + h), 1e-3 * c)]))


def gpr_td2_fit(kind, x, J, y, y_derivs, nv_first, ell, sigma_targets, sigma_derivs, sigma_derivs2, mean_mode=1):
    Ja, Jb, ya, yb = _td2_split(J, y_derivs, nv_first)
    c, mu = R.td2_fit_dense(kind, _np(x), _col(y), Ja, Jb, ya, yb, float(ell), float(sigma_targets), float(sigma_derivs),
                            float(sigma_derivs2), _mean_fn(mean_mode))
    n, n1, n2 = x.shape[0], Ja.shape[2], Jb.shape[2]   # reference order [targets; block 1; block 2] -> library order
    a, b = c[n:n + n * n1].reshape(n, n1), c[n + n * n1:].reshape(n, n2)
    lib = np.concatenate([c[:n, 0], np.concatenate([a, b], axis=1).reshape(-1)]).reshape(-1, 1)
    return _t(lib), _t(np.array([float(mu)]))


# ---- derivative-trained GPR (csrc/hess.cu) ----
def hess_gram(kind, x1, J1, x2, J2, ell, diag_add=0.0, lower_only=False):
    sym = (x2 is x1 or x2 is None) and (J2 is J1 or J2 is None)
    a, Ja = _np(x1), _np(J1)
    b, Jb = (a, Ja) if sym else (_np(x2), _np(J2))
    K = R.d01kj(kind, a, b, float(ell), Ja, Jb)
    if diag_add:
        K = K + diag_add * np.eye(K.shape[0], K.shape[1])
    return _t(np.tril(K) if lower_only else K)


def gpr_lml_derivs(kind, x, J, y, ell, sigma):
    return _t(np.array([R.lml_derivs_dense(kind, _np(x), _np(y), _np(J), float(ell), float(sigma))]))


def gpr_lml_grad_derivs(kind, x, J, y, ell, sigma, want_grad=True):
    out = np.array(R.lml_derivs_grad_dense(kind, _np(x), _np(y), _np(J), float(ell), float(sigma)))
    if not want_grad:
        out[1:] = 0.0
    return _t(out)


def gpr_fit_derivs(kind, x, J, y, ell, sigma):
    c, _, jc = R.fit_derivs_dense(kind, _np(x), _np(y), _np(J), float(ell), float(sigma))
    return _t(c), _t(jc)


def gpr_predict_y_derivs(kind, x_train, jaccoef, x, ell, mu=0.0):
    return _t(np.asarray(R.predict_y_derivs_dense(kind, _np(x_train), _np(x), _np(jaccoef), float(mu), float(ell))).reshape(-1))


def gpr_predict_derivs_full(kind, x_train, J_train, x, J, c, ell, sigma, mu=0.0, full_covariance=False):
    mean, cov = R.predict_derivs_dense_full(kind, _np(x_train), _np(J_train), _np(x), _np(J), _np(c).reshape(-1, 1),
                                            float(mu), float(ell), float(sigma))
    return (_t(mean.reshape(-1)), _t(cov)) if full_covariance else _t(mean.reshape(-1))


def gpr_predict_y_derivs_full(kind, x_train, J_train, x, c, ell, sigma, mu=0.0, full_covariance=False):
    mean, cov = R.predict_y_derivs_dense_full(kind, _np(x_train), _np(J_train), _np(x), _np(c).reshape(-1, 1), float(mu),
                                              float(ell), float(sigma))
    return (_t(mean.reshape(-1)), _t(cov)) if full_covariance else _t(mean.reshape(-1))


def hess_matvec(kind, x1, J1, x2, J2, V, ell, diag_add=0.0):
    K = R.d01kj(kind, _np(x1), _np(x2), float(ell), _np(J1), _np(J2))
    if diag_add:
        K = K + diag_add * np.eye(K.shape[0], K.shape[1])
    v = _np(V)
    return _t(K @ v)


FAKES = dict(gram=gram, gram_composite=gram_composite, gram_composite_fits=gram_composite_fits,
             gram_dl_contract=gram_dl_contract, elementwise=elementwise, gemm=gemm, potrf=potrf, trsm=trsm,
             logdet=logdet, potri=potri, gpr_lml_grad=gpr_lml_grad, gpr_fit=gpr_fit, gpr_predict=gpr_predict,
             kmatvec=kmatvec, rpcholesky=rpcholesky, gather_rows=gather_rows, sgpr_lml=sgpr_lml,
             sgpr_lml_grad=sgpr_lml_grad, sgpr_fit=sgpr_fit, sgpr_predict=sgpr_predict, td_gram=td_gram,
             gpr_td_lml_grad=gpr_td_lml_grad, gpr_td_fit=gpr_td_fit, gpr_td2_lml_grad=gpr_td2_lml_grad,
             gpr_td2_fit=gpr_td2_fit, hess_gram=hess_gram, gpr_lml_derivs=gpr_lml_derivs,
             gpr_lml_grad_derivs=gpr_lml_grad_derivs, gpr_fit_derivs=gpr_fit_derivs,
             gpr_predict_y_derivs=gpr_predict_y_derivs, gpr_predict_derivs_full=gpr_predict_derivs_full,
             gpr_predict_y_derivs_full=gpr_predict_y_derivs_full, hess_matvec=hess_matvec)


@contextlib.contextmanager
def patched():
    """Swap the listed gpx_b200.ops entry points and every module-level `_to_device` for the CPU stand-ins."""
    import sys

    from gpx_b200 import ops

    saved_ops = {k: getattr(ops, k) for k in FAKES}
    saved_td = []
    for name, mod in list(sys.modules.items()):
        if name.startswith("gpx_b200") and mod is not None and hasattr(mod, "_to_device"):
            saved_td.append((mod, mod._to_device))
    try:
        for k, f in FAKES.items():
            setattr(ops, k, f)
        for mod, _ in saved_td:
            mod._to_device = to_device
        yield
    finally:
        for k, f in saved_ops.items():
            setattr(ops, k, f)
        for mod, f in saved_td:
            mod._to_device = f
