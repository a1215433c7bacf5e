"""GPU parity: Hessian-Jacobian kernel (Family 1b) and GPR on derivative values (config 3 shape:
x = coordinates, J = identity + dense perturbation) against the oracle's literal einsum form."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as R  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def _xj(N, nf, nv, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, nf))
    J = 0.1 * rng.standard_normal((N, nf, nv))
    for k in range(min(nf, nv)):
        J[:, k, k] += 1.0
    y = rng.standard_normal((N, nv))
    return x, J, y


@pytest.mark.parametrize("kind", ["se", "m52"])
@pytest.mark.parametrize("N1,N2,nf,nv1,nv2", [(7, 5, 6, 6, 4), (20, 31, 9, 9, 9), (5, 5, 2, 1, 1), (12, 3, 90, 90, 90)])
def test_d01kj_entries(kind, N1, N2, nf, nv1, nv2):
    from gpx_b200 import ops

    x1, J1, _ = _xj(N1, nf, nv1, 1)
    x2, J2, _ = _xj(N2, nf, nv2, 2)
    ell = 0.9 * np.sqrt(nf)
    ref = R.d01kj(kind, x1, x2, ell, J1, J2)
    out = host(ops.hess_gram(kind, dev(x1), dev(J1), dev(x2), dev(J2), ell))
    assert np.max(np.abs(out - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("kind", ["se", "m52"])
def test_d01kj_symmetric_lower_with_noise(kind):
    from gpx_b200 import ops

    x, J, _ = _xj(40, 12, 12, 3)
    ell = 3.0
    ref = R.d01kj(kind, x, x, ell, J, J) + 0.04 * np.eye(480)
    xd, Jd = dev(x), dev(J)
    out = host(ops.hess_gram(kind, xd, Jd, xd, Jd, ell, diag_add=0.04, lower_only=True))
    assert np.max(np.abs(np.tril(out) - np.tril(ref))) < 1e-12
    assert np.all(np.triu(out, 1) == 0.0)


@pytest.mark.parametrize("kind,N,nf,nv", [("se", 64, 90, 90), ("m52", 50, 30, 30), ("se", 128, 12, 9)])
def test_lml_and_fit_derivs(kind, N, nf, nv):
    from gpx_b200 import ops

    x, J, y = _xj(N, nf, nv, 2)
    ell, sigma = np.sqrt(nf), 0.05
    ref = R.lml_derivs_dense(kind, x, y, J, ell, sigma)
    out = float(host(ops.gpr_lml_derivs(kind, dev(x), dev(J), dev(y), ell, sigma))[0])
    assert abs(out - ref) < 1e-8 * abs(ref)
    c_ref, _, jc_ref = R.fit_derivs_dense(kind, x, y, J, ell, sigma)
    c, jc = ops.gpr_fit_derivs(kind, dev(x), dev(J), dev(y), ell, sigma)
    assert np.max(np.abs(host(c) - c_ref)) < 1e-8 * np.max(np.abs(c_ref))
    assert np.max(np.abs(host(jc) - jc_ref)) < 1e-8 * np.max(np.abs(jc_ref))


@pytest.mark.parametrize("kind,N,nf,nv", [("se", 20, 30, 30), ("se", 64, 32, 32), ("m52", 100, 30, 30), ("se", 37, 90, 90)])
def test_packed_lower_storage_matches_square(kind, N, nf, nv, monkeypatch):
    """Block-column packed lower storage (used when the square Hessian system does not fit in HBM:
    config 3 at N = 2000 is 259 GB square) against the square-storage path and the oracle; the sizes
    cover one panel, an exact multiple of the 1024 panel width, and ragged last panels."""
    from gpx_b200 import ops

    x, J, y = _xj(N, nf, nv, 11)
    ell, sigma = np.sqrt(nf), 0.05
    xd, Jd, yd = dev(x), dev(J), dev(y)
    monkeypatch.setenv("GPXB_DERIVS_LAYOUT", "square")
    lml_sq = float(host(ops.gpr_lml_derivs(kind, xd, Jd, yd, ell, sigma))[0])
    c_sq, jc_sq = [host(t) for t in ops.gpr_fit_derivs(kind, xd, Jd, yd, ell, sigma)]
    monkeypatch.setenv("GPXB_DERIVS_LAYOUT", "packed")
    lml_pk = float(host(ops.gpr_lml_derivs(kind, xd, Jd, yd, ell, sigma))[0])
    c_pk, jc_pk = [host(t) for t in ops.gpr_fit_derivs(kind, xd, Jd, yd, ell, sigma)]
    assert abs(lml_pk - lml_sq) <= 1e-12 * abs(lml_sq)
    assert np.max(np.abs(c_pk - c_sq)) <= 1e-10 * np.max(np.abs(c_sq))
    assert np.max(np.abs(jc_pk - jc_sq)) <= 1e-10 * np.max(np.abs(jc_sq))
    if N * nv <= 2100:
        ref = R.lml_derivs_dense(kind, x, y, J, ell, sigma)
        assert abs(lml_pk - ref) < 1e-8 * abs(ref)
        c_ref, _, _ = R.fit_derivs_dense(kind, x, y, J, ell, sigma)
        assert np.max(np.abs(c_pk - c_ref)) < 1e-8 * np.max(np.abs(c_ref))


@pytest.mark.parametrize("kind,N,nf,nv", [("se", 40, 12, 12), ("m52", 40, 12, 9), ("se", 24, 90, 90), ("m52", 30, 30, 30),
                                           ("se", 7, 3, 1)])
def test_lml_gradient_derivs(kind, N, nf, nv):
    """(lml, d/d ell, d/d sigma) of the derivative-trained GPR: Hessian kernel in contraction mode against
    the oracle's term-by-term derivative of the literal einsums."""
    from gpx_b200 import ops

    x, J, y = _xj(N, nf, nv, 4)
    ell, sigma = 0.8 * np.sqrt(nf), 0.07
    ref = np.array(R.lml_derivs_grad_dense(kind, x, y, J, ell, sigma))
    out = host(ops.gpr_lml_grad_derivs(kind, dev(x), dev(J), dev(y), ell, sigma))
    assert np.all(np.abs(out - ref) < 1e-8 * np.abs(ref)), (out, ref)
    only = host(ops.gpr_lml_grad_derivs(kind, dev(x), dev(J), dev(y), ell, sigma, want_grad=False))
    assert abs(only[0] - ref[0]) < 1e-8 * abs(ref[0]) and only[1] == 0.0 and only[2] == 0.0


def test_facade_fit_derivs_minimize_follows_oracle_optimiser():
    """BaseGP.fit_derivs(minimize=True) (gpx/models/base.py:487-589): L-BFGS-B on the device (value, grad)
    must land where the same optimiser lands with the oracle's (value, grad)."""
    from scipy.optimize import minimize

    from gpx_b200.kernels import SquaredExponential
    from gpx_b200.models import GPR
    from gpx_b200.models.gpr import neg_log_marginal_likelihood_derivs

    rng = np.random.default_rng(8)
    N, nf = 60, 3
    x = rng.uniform(-2, 2, (N, nf))
    J = np.tile(np.eye(nf)[None], (N, 1, 1))
    y = np.stack([np.cos(x[:, 0]) * np.sin(x[:, 1]), np.sin(x[:, 0]) * np.cos(x[:, 1]), 0.5 * np.ones(N) * x[:, 2]], 1)
    y = y + 0.01 * rng.standard_normal(y.shape)
    m = GPR(SquaredExponential(), loss_fn=neg_log_marginal_likelihood_derivs).fit_derivs(x, y, J)
    assert m.optimize_results_.success

    def fun(t):
        ell, sig = R.softplus(t[0]), R.softplus(t[1])
        lml, ge, gs = R.lml_derivs_grad_dense("se", x, y, J, ell, sig)
        return -lml, np.array([-ge * R.sigmoid(t[0]), -gs * R.sigmoid(t[1])])

    res = minimize(fun, R.inverse_softplus(np.array([1.0, 1.0])), jac=True, method="L-BFGS-B")
    ell, sig = float(m.state.params["kernel_params"]["lengthscale"].value), float(m.state.params["sigma"].value)
    assert abs(m.optimize_results_.fun - res.fun) < 1e-6 * max(1.0, abs(res.fun))
    assert abs(ell - R.softplus(res.x[0])) < 1e-3 * ell and abs(sig - R.softplus(res.x[1])) < 1e-3
    c_ref, _, jc_ref = R.fit_derivs_dense("se", x, y, J, ell, sig)
    assert np.max(np.abs(m.state.jaccoef - jc_ref)) < 1e-7 * np.max(np.abs(jc_ref))
    with pytest.raises(ValueError):  # the iterative fit needs the preconditioner size
        GPR(SquaredExponential()).fit_derivs(x, y, J, minimize=False, iterative=True)


@pytest.mark.parametrize("kind", ["se", "m52"])
@pytest.mark.parametrize("N,nf,ns", [(50, 6, 11), (300, 90, 257), (1000, 31, 64), (3, 1, 1)])
def test_predict_y_derivs(kind, N, nf, ns):
    """Target values from a derivative-trained model: one matrix-free block mat-vec with nf + 1 columns
    against the oracle's literal d0kjc."""
    from gpx_b200 import ops

    rng = np.random.default_rng(12)
    xt, xs, jc = rng.standard_normal((N, nf)), rng.standard_normal((ns, nf)), rng.standard_normal((N, nf))
    ell, mu = 0.9 * np.sqrt(nf), 0.3
    ref = R.predict_y_derivs_dense(kind, xt, xs, jc, mu, ell)[:, 0]
    out = host(ops.gpr_predict_y_derivs(kind, dev(xt), dev(jc), dev(xs), ell, mu))
    assert np.max(np.abs(out - ref)) < 1e-8 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("kind,N,nf,nv,ns", [("se", 30, 6, 6, 7), ("m52", 25, 9, 4, 5), ("se", 12, 90, 90, 3)])
def test_predict_derivs_and_targets_with_full_covariance(kind, N, nf, nv, ns):
    """_predict_derivs_dense / _predict_y_derivs_dense with full_covariance=True (uncontracted coefficients)."""
    from gpx_b200 import ops

    x, J, y = _xj(N, nf, nv, 21)
    xs, Js, _ = _xj(ns, nf, nv, 22)
    ell, sigma = 0.9 * np.sqrt(nf), 0.1
    c_ref, _, jc_ref = R.fit_derivs_dense(kind, x, y, J, ell, sigma)
    m_ref, C_ref = R.predict_derivs_dense_full(kind, x, J, xs, Js, c_ref, 0.0, ell, sigma)
    mean, cov = ops.gpr_predict_derivs_full(kind, dev(x), dev(J), dev(xs), dev(Js), dev(c_ref[:, 0]), ell, sigma,
                                            full_covariance=True)
    assert np.max(np.abs(host(mean) - m_ref[:, 0])) < 1e-8 * np.max(np.abs(m_ref))
    assert np.max(np.abs(host(cov) - C_ref)) < 1e-8 * np.max(np.abs(C_ref))
    only = ops.gpr_predict_derivs_full(kind, dev(x), dev(J), dev(xs), dev(Js), dev(c_ref[:, 0]), ell, sigma)
    assert np.array_equal(host(only), host(mean))
    my_ref, Cy_ref = R.predict_y_derivs_dense_full(kind, x, J, xs, c_ref, 0.25, ell, sigma)
    ymean, ycov = ops.gpr_predict_y_derivs_full(kind, dev(x), dev(J), dev(xs), dev(c_ref[:, 0]), ell, sigma, mu=0.25,
                                                full_covariance=True)
    assert np.max(np.abs(host(ymean) - my_ref[:, 0])) < 1e-8 * np.max(np.abs(my_ref))
    assert np.max(np.abs(host(ycov) - Cy_ref)) < 1e-8 * max(1.0, np.max(np.abs(Cy_ref)))
    # the contracted fast path gives the same mean
    fast = host(ops.gpr_predict_y_derivs(kind, dev(x), dev(jc_ref), dev(xs), ell, 0.25))
    assert np.max(np.abs(fast - my_ref[:, 0])) < 1e-8 * np.max(np.abs(my_ref))


@pytest.mark.parametrize("kind", ["se", "m52"])
@pytest.mark.parametrize("N1,N2,nf,nv1,nv2,P", [(30, 30, 6, 6, 6, 3), (17, 40, 9, 4, 9, 1), (8, 5, 90, 90, 90, 2), (5, 5, 1, 1, 1, 1)])
def test_matrix_free_hessian_operator(kind, N1, N2, nf, nv1, nv2, P):
    """_Ax_derivs_lhs_fun collapsed onto the configuration level against the dense Hessian kernel."""
    from gpx_b200 import ops

    x1, J1, _ = _xj(N1, nf, nv1, 31)
    x2, J2, _ = _xj(N2, nf, nv2, 32)
    V = np.random.default_rng(5).standard_normal((N2 * nv2, P))
    ell = 0.9 * np.sqrt(nf)
    ref = R.d01kj(kind, x1, x2, ell, J1, J2) @ V
    out = host(ops.hess_matvec(kind, dev(x1), dev(J1), dev(x2), dev(J2), dev(V), ell))
    assert np.max(np.abs(out - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))
    if N1 == N2 and nv1 == nv2:  # the symmetric operator with the noise shift
        xd, Jd = dev(x1), dev(J1)
        refs = (R.d01kj(kind, x1, x1, ell, J1, J1) + 0.3 * np.eye(N1 * nv1)) @ V
        outs = host(ops.hess_matvec(kind, xd, Jd, xd, Jd, dev(V), ell, diag_add=0.3))
        assert np.max(np.abs(outs - refs)) < 1e-10 * max(1.0, np.max(np.abs(refs)))


@pytest.mark.parametrize("kind,N,nf,nv,npiv", [("se", 40, 6, 6, 20), ("m52", 30, 9, 4, 15), ("se", 12, 90, 90, 50)])
def test_iterative_fit_and_lml_on_derivatives(kind, N, nf, nv, npiv):
    """fit_derivs / log_marginal_likelihood_derivs with iterative=True: rpcholesky_derivs preconditioner, PCG and
    SLQ on the matrix-free Hessian operator, against the oracle (same iteration count, same LML)."""
    from gpx_b200 import ops
    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.models import GPR
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    x, J, y = _xj(N, nf, nv, 41)
    ell, sigma = 0.9 * np.sqrt(nf), 0.3
    key = R.PRNGKey(2023)
    c_ref, it_ref = R.fit_derivs_iter(kind, x, y, J, ell, sigma, npiv, key)
    c, jc, iters = ops.gpr_fit_derivs_iter(kind, dev(x), dev(J), dev(y), ell, sigma, npiv, key)
    assert iters == it_ref
    assert np.max(np.abs(host(c) - c_ref[:, 0])) < 1e-6 * np.max(np.abs(c_ref))
    yf = y.reshape(-1)
    assert abs(yf @ host(c) - yf @ c_ref[:, 0]) < 1e-8 * abs(yf @ c_ref[:, 0])
    # without the preconditioner the same solution is reached (more iterations)
    c0, _, it0 = ops.gpr_fit_derivs_iter(kind, dev(x), dev(J), dev(y), ell, sigma, 0, key)
    c_dense, _, _ = R.fit_derivs_dense(kind, x, y, J, ell, sigma)
    assert np.max(np.abs(host(c0) - c_dense[:, 0])) < 1e-3 * np.max(np.abs(c_dense))
    kern = SquaredExponential() if kind == "se" else Matern52()
    m = GPR(kern, kernel_params=dict(lengthscale=Parameter(ell, True, Softplus(), GammaPrior())),
            sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    ref = R.lml_derivs_iter(kind, x, y, J, ell, sigma, 5, 10, key, npiv, key)
    out = m.log_marginal_likelihood_derivs(x, y, J, iterative=True, num_evals=5, num_lanczos=10, key_lanczos=key,
                                           n_pivots=npiv, key_precond=key)
    assert abs(out - ref) < 1e-7 * abs(ref)  # PCG at tol 1e-5 inside (see test_iterative_derivative_lml_gradient)
    m.fit_derivs(x, y, J, minimize=False, iterative=True, n_pivots=npiv)
    assert m.state.c.shape == (N * nv, 1) and m.state.jaccoef.shape == (N, nf)


@pytest.mark.parametrize("kind,N,nf,nv", [("se", 40, 6, 6), ("m52", 30, 9, 4)])
def test_iterative_derivative_lml_gradient(kind, N, nf, nv):
    """value_and_grad of _lml_derivs_iter: d/d ell operator of the matrix-free Hessian product and the tangent
    Lanczos on it against the oracle; fit_derivs(minimize=True, iterative=True) then trains on it."""
    from gpx_b200 import ops
    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.models import GPR
    from gpx_b200.models.gpr import neg_log_marginal_likelihood_derivs
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    x, J, y = _xj(N, nf, nv, 47)
    ell, sigma = 0.9 * np.sqrt(nf), 0.3
    key = R.PRNGKey(2023)
    V = np.random.default_rng(3).standard_normal((N * nv, 2))
    ref_dl = R.d01kj_dl(kind, x, x, ell, J, J) @ V
    out_dl = host(ops.hess_matvec_dl(kind, dev(x), dev(J), dev(V), ell))
    assert np.max(np.abs(out_dl - ref_dl)) < 1e-10 * max(1.0, np.max(np.abs(ref_dl)))
    ref = R.lml_derivs_iter_grad(kind, x, y, J, ell, sigma, 5, 10, key, 20, key)
    kern = SquaredExponential() if kind == "se" else Matern52()
    mk = lambda: GPR(kern, kernel_params=dict(lengthscale=Parameter(ell, True, Softplus(), GammaPrior())),  # noqa: E731
                     sigma=Parameter(sigma, True, Softplus(), GammaPrior()), loss_fn=neg_log_marginal_likelihood_derivs)
    loss, grads = mk()._loss_and_grad_derivs_fun(mk().state, x, y, J, iterative=True, num_evals=5, num_lanczos=10,
                                                 key_lanczos=key, n_pivots=20, key_precond=key)
    # the value goes through PCG stopped at a 1e-5 relative residual: rounding-level differences in the
    # preconditioner move the iterates, and with them y.c, at the 1e-8 level
    assert abs(-loss - ref[0]) < 1e-7 * abs(ref[0])
    assert abs(-grads[("kernel_params", "lengthscale")] - ref[1]) < 1e-6 * max(abs(ref[1]), abs(ref[0]))
    assert abs(-grads[("sigma",)] - ref[2]) < 1e-6 * max(abs(ref[2]), abs(ref[0]))
    m = mk().fit_derivs(x, y, J, iterative=True, n_pivots=20, loss_kwargs=dict(num_evals=5, num_lanczos=10, key_lanczos=key),
                        opt_kwargs=dict(options=dict(maxiter=4)))
    assert m.optimize_results_.fun < loss


@pytest.mark.parametrize("kind,N,nf,nv,M,part", [("se", 25, 5, 5, 40, True), ("m52", 60, 9, 4, 64, True),
                                                  ("se", 30, 6, 6, 33, False), ("se", 14, 90, 90, 100, True)])
def test_rpcholesky_derivs_pivots_bit_exact(kind, N, nf, nv, M, part):
    """rpcholesky_derivs: pivots bit-exact against the oracle (both jax_threefry_partitionable modes), factor to
    1e-9, and F F^T approaching the Hessian kernel."""
    from gpx_b200 import ops

    x, J, _ = _xj(N, nf, nv, 43)
    ell = 0.9 * np.sqrt(nf)
    key = R.PRNGKey(7)
    F_ref, piv_ref = R.rpcholesky_derivs(kind, x, J, M, ell, key, part)
    F, piv = ops.rpcholesky_derivs(kind, dev(x), dev(J), M, ell, key, part)
    assert np.array_equal(piv.cpu().numpy(), piv_ref)
    assert np.max(np.abs(host(F) - F_ref)) < 1e-9 * np.max(np.abs(F_ref))
    _, piv2 = ops.rpcholesky_derivs(kind, dev(x), dev(J), M, ell, key, part, want_factor=False)
    assert np.array_equal(piv2.cpu().numpy(), piv_ref)


@pytest.mark.parametrize("kind", ["se", "m52"])
def test_sgpr_fit_and_predict_on_derivatives(kind):
    """SGPR.fit_derivs / predict_derivs / predict_y_derivs (projected processes on the Hessian kernel)."""
    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.models import SGPR
    from gpx_b200.models.sgpr import locs_parameter
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Identity, Softplus
    from gpx_b200.priors import GammaPrior, NormalPrior

    N, nf, nv, Ml = 60, 5, 4, 12
    x, J, y = _xj(N, nf, nv, 51)
    xl, Jl, _ = _xj(Ml, nf, nv, 52)
    xs, Js, _ = _xj(9, nf, nv, 53)
    ell, sigma = 2.0, 0.2
    kern = SquaredExponential() if kind == "se" else Matern52()
    m = SGPR(kern, x_locs=locs_parameter(xl), jacobian_locs=Parameter(Jl, False, Identity(), NormalPrior(shape=Jl.shape)),
             kernel_params=dict(lengthscale=Parameter(ell, True, Softplus(), GammaPrior())),
             sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    m.fit_derivs(x, y, J, minimize=False)
    c_ref, jc_ref = R.sgpr_fit_derivs_dense(kind, xl, Jl, x, y, J, ell, sigma)
    assert np.max(np.abs(m.state.c - c_ref)) < 1e-7 * np.max(np.abs(c_ref))
    assert np.max(np.abs(m.state.jaccoef - jc_ref)) < 1e-7 * np.max(np.abs(jc_ref))
    pred = m.predict_derivs(xs, Js)
    ref = (R.d01kj(kind, xs, xl, ell, Js, Jl) @ c_ref).reshape(9, nv)
    assert np.max(np.abs(pred - ref)) < 1e-7 * np.max(np.abs(ref))
    yp = m.predict_y_derivs(xs)
    ref_y = R.predict_y_derivs_dense(kind, xl, xs, jc_ref, 0.0, ell)
    assert np.max(np.abs(yp - ref_y)) < 1e-7 * np.max(np.abs(ref_y))


@pytest.mark.parametrize("kind", ["se", "m52"])
def test_sgpr_cg_fit_on_derivatives(kind):
    """_fit_derivs_iter (gpx/models/_sgpr.py:399-428): CG on sigma^2 H_mm + H_mn H_nm + 1e-10 I through the
    matrix-free Hessian operators, against the oracle's CG on the dense matrices (same iteration count with the
    reference's stopping rule; 1e-8 on the coefficients' effect with both solves converged) and via the facade."""
    from gpx_b200 import ops
    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.models import SGPR
    from gpx_b200.models.sgpr import locs_parameter
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Identity, Softplus
    from gpx_b200.priors import GammaPrior, NormalPrior

    N, nf, nv, Ml = 60, 5, 4, 8
    x, J, y = _xj(N, nf, nv, 51)
    xl, Jl, _ = _xj(Ml, nf, nv, 52)
    xs, Js, _ = _xj(9, nf, nv, 53)
    ell, sigma = 2.0, 0.5
    yf = y.reshape(-1, 1)
    K_mn = R.d01kj(kind, xl, x, ell, Jl, J)
    C = sigma**2 * R.d01kj(kind, xl, xl, ell, Jl, Jl) + K_mn @ K_mn.T + 1e-10 * np.eye(K_mn.shape[0])
    c_cg, it_ref = R.cg(lambda v: C @ v, K_mn @ yf)
    c, it = ops.sgpr_fit_derivs_iter(kind, dev(x), dev(J), dev(y.reshape(-1)), dev(xl), dev(Jl), ell, sigma)
    assert abs(it - it_ref) <= 1
    pred_ref = R.d01kj(kind, xs, xl, ell, Js, Jl) @ c_cg
    pred = R.d01kj(kind, xs, xl, ell, Js, Jl) @ host(c)
    assert np.max(np.abs(pred - pred_ref)) < 1e-4 * np.max(np.abs(pred_ref))  # both stopped at tol 1e-5
    # converged on both sides: the solution itself
    c_t, _ = ops.sgpr_fit_derivs_iter(kind, dev(x), dev(J), dev(y.reshape(-1)), dev(xl), dev(Jl), ell, sigma, tol=1e-14,
                                      maxiter=20000)
    c_d = np.linalg.solve(C, K_mn @ yf)
    pred_d = R.d01kj(kind, xs, xl, ell, Js, Jl) @ c_d
    pred_t = R.d01kj(kind, xs, xl, ell, Js, Jl) @ host(c_t)
    assert np.max(np.abs(pred_t - pred_d)) < 1e-8 * np.max(np.abs(pred_d))
    kern = SquaredExponential() if kind == "se" else Matern52()
    m = SGPR(kern, x_locs=locs_parameter(xl), jacobian_locs=Parameter(Jl, False, Identity(), NormalPrior(shape=Jl.shape)),
             kernel_params=dict(lengthscale=Parameter(ell, True, Softplus(), GammaPrior())),
             sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    m.fit_derivs(x, y, J, minimize=False, iterative=True)
    assert np.max(np.abs(m.predict_derivs(xs, Js).reshape(-1, 1) - pred_ref)) < 1e-4 * np.max(np.abs(pred_ref))


def test_facade_fit_and_predict_derivs():
    from gpx_b200.kernels import SquaredExponential
    from gpx_b200.models import GPR
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    N, nf, nv = 40, 6, 6
    x, J, y = _xj(N, nf, nv, 5)
    xs, Js, _ = _xj(11, nf, nv, 6)
    ell, sigma = 2.0, 0.1
    m = GPR(SquaredExponential(), kernel_params=dict(lengthscale=Parameter(ell, True, Softplus(), GammaPrior())),
            sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    ref = R.lml_derivs_dense("se", x, y, J, ell, sigma)
    assert abs(m.log_marginal_likelihood_derivs(x, y, J) - ref) < 1e-8 * abs(ref)
    m.fit_derivs(x, y, J, minimize=False)
    c_ref, _, jc_ref = R.fit_derivs_dense("se", x, y, J, ell, sigma)
    assert np.max(np.abs(m.state.jaccoef - jc_ref)) < 1e-8 * np.max(np.abs(jc_ref))
    pred = m.predict_derivs(xs, Js)
    ref_pred = (R.d01kj("se", x, xs, ell, J, Js).T @ c_ref).reshape(11, nv)
    assert np.max(np.abs(pred - ref_pred)) < 1e-8 * np.max(np.abs(ref_pred))
    # the contracted route of the reference (d01kjc) gives the same numbers
    ref_c = R.d01kjc("se", x, xs, ell, jc_ref, Js).sum(0).reshape(11, nv)
    assert np.max(np.abs(pred - ref_c)) < 1e-8 * np.max(np.abs(ref_c))
    pred_it = m.predict_derivs(xs, Js, iterative=True)
    assert np.max(np.abs(pred_it - ref_pred)) < 1e-8 * np.max(np.abs(ref_pred))
    mu_c, cov_c = m.predict_derivs(xs, Js, full_covariance=True)
    m_ref, C_ref = R.predict_derivs_dense_full("se", x, J, xs, Js, c_ref, 0.0, ell, sigma)
    assert mu_c.shape == (11 * nv, 1) and np.max(np.abs(cov_c - C_ref)) < 1e-8 * np.max(np.abs(C_ref))
    assert np.max(np.abs(mu_c - m_ref)) < 1e-8 * np.max(np.abs(m_ref))
    y_c, ycov_c = m.predict_y_derivs(xs, full_covariance=True)
    my_ref, Cy_ref = R.predict_y_derivs_dense_full("se", x, J, xs, c_ref, 0.0, ell, sigma)
    assert np.max(np.abs(ycov_c - Cy_ref)) < 1e-8 and np.max(np.abs(y_c - my_ref)) < 1e-8 * np.max(np.abs(my_ref))
    yp = m.predict_y_derivs(xs)
    ref_y = R.predict_y_derivs_dense("se", x, xs, jc_ref, 0.0, ell)
    assert yp.shape == (11, 1) and np.max(np.abs(yp - ref_y)) < 1e-8 * np.max(np.abs(ref_y))
    with pytest.raises(RuntimeError):
        m.predict(xs)
