"""GPU parity: GPR on targets and derivatives (GPR_TD) against the oracle's literal block
assembly (oracle/ref_numpy.py: td_A_lhs / td_lml_dense / td_fit_dense / td_predict_dense)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as R  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def _data(N, nf, nv, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, nf))
    J = 0.1 * rng.standard_normal((N, nf, nv))
    for k in range(min(nf, nv)):
        J[:, k, k] += 1.0
    y = np.sin(x.sum(1, keepdims=True)) + 0.05 * rng.standard_normal((N, 1))
    yd = rng.standard_normal((N, nv))
    return x, J, y, yd


@pytest.mark.parametrize("kind", ["se", "m52"])
@pytest.mark.parametrize("N1,N2,nf,nv1,nv2", [(20, 20, 5, 5, 5), (13, 31, 6, 3, 6), (7, 5, 90, 90, 90)])
def test_td_block_matrix(kind, N1, N2, nf, nv1, nv2):
    from gpx_b200 import ops

    x1, J1, _, _ = _data(N1, nf, nv1, 1)
    x2, J2, _, _ = _data(N2, nf, nv2, 2)
    ell = 0.9 * np.sqrt(nf)
    ref = R.td_A_lhs(kind, x1, J1, x2, J2, ell, 0.0, 0.0, noise=False)
    out = host(ops.td_gram(kind, dev(x1), dev(J1), dev(x2), dev(J2), ell))
    assert np.max(np.abs(out - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
    if N1 == N2 and nv1 == nv2:
        xd, Jd = dev(x1), dev(J1)
        refs = R.td_A_lhs(kind, x1, J1, x1, J1, ell, 0.3, 0.2)
        outs = host(ops.td_gram(kind, xd, Jd, xd, Jd, ell, add_targets=0.3**2 + 1e-10, add_derivs=0.2**2 + 1e-10))
        assert np.max(np.abs(outs - refs)) < 1e-11 * max(1.0, np.max(np.abs(refs)))
        low = host(ops.td_gram(kind, xd, Jd, xd, Jd, ell, add_targets=0.3**2 + 1e-10, add_derivs=0.2**2 + 1e-10,
                               lower_only=True))
        assert np.max(np.abs(np.tril(low) - np.tril(refs))) < 1e-11 * max(1.0, np.max(np.abs(refs)))
        assert np.all(np.triu(low, 1) == 0.0)


@pytest.mark.parametrize("kind,N,nf,nv,mean", [("se", 40, 6, 6, "zero"), ("m52", 30, 9, 4, "data"), ("se", 10, 90, 90, "zero")])
def test_td_lml_gradient_fit_predict(kind, N, nf, nv, mean):
    from gpx_b200 import ops
    from gpx_b200.mean_functions import data_mean, zero_mean

    x, J, y, yd = _data(N, nf, nv, 3)
    xs, Js, _, _ = _data(9, nf, nv, 4)
    ell, st, sd = 0.9 * np.sqrt(nf), 0.3, 0.2
    mf_o = R.data_mean if mean == "data" else R.zero_mean
    mode = 1 if mean == "data" else 0
    f = lambda l, a, b: R.td_lml_dense(kind, x, y, J, yd, l, a, b, mf_o)  # noqa: E731
    out = host(ops.gpr_td_lml_grad(kind, dev(x), dev(J), dev(y), dev(yd), ell, st, sd, mode))
    ref = f(ell, st, sd)
    assert abs(out[0] - ref) < 1e-8 * abs(ref)

    def rich(g, h=1e-3):
        fd = lambda hh: (g(hh) - g(-hh)) / (2 * hh)  # noqa: E731
        return (4 * fd(h / 2) - fd(h)) / 3

    g_ref = [rich(lambda h: f(ell + h, st, sd)), rich(lambda h: f(ell, st + h, sd)), rich(lambda h: f(ell, st, sd + h))]
    for k in range(3):
        assert abs(out[1 + k] - g_ref[k]) < 1e-7 * max(abs(g_ref[k]), abs(ref)), (k, out, g_ref)
    only = host(ops.gpr_td_lml_grad(kind, dev(x), dev(J), dev(y), dev(yd), ell, st, sd, mode, want_grad=False))
    assert only[0] == out[0] and np.all(only[1:] == 0.0)
    c_ref, mu_ref = R.td_fit_dense(kind, x, y, J, yd, ell, st, sd, mf_o)
    c, mu = ops.gpr_td_fit(kind, dev(x), dev(J), dev(y), dev(yd), ell, st, sd, mode)
    assert abs(float(host(mu)[0]) - mu_ref) < 1e-14
    assert np.max(np.abs(host(c) - c_ref)) < 1e-8 * np.max(np.abs(c_ref))

    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.models import GPR_TD
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    par = lambda v: Parameter(v, True, Softplus(), GammaPrior())  # noqa: E731
    kern = SquaredExponential() if kind == "se" else Matern52()
    m = GPR_TD(kern, mean_function=data_mean if mean == "data" else zero_mean,
               kernel_params=dict(lengthscale=par(ell)), sigma_targets=par(st), sigma_derivs=par(sd))
    assert abs(m.log_marginal_likelihood(x, y, J, yd) - ref) < 1e-8 * abs(ref)
    m.fit(x, y, J, yd, minimize=False)
    yp, ydp = m.predict(xs, Js)
    yr, ydr = R.td_predict_dense(kind, x, J, xs, Js, c_ref, mu_ref, ell)
    assert np.max(np.abs(yp - yr)) < 1e-8 * max(1.0, np.max(np.abs(yr)))
    assert np.max(np.abs(ydp - ydr)) < 1e-8 * max(1.0, np.max(np.abs(ydr)))


def test_td_fit_minimize_lowers_the_loss():
    from gpx_b200.kernels import SquaredExponential
    from gpx_b200.models import GPR_TD

    rng = np.random.default_rng(6)
    N = 50
    x = rng.uniform(-2, 2, (N, 2))
    J = np.tile(np.eye(2)[None], (N, 1, 1))
    y = (np.sin(x[:, 0]) * np.cos(x[:, 1])).reshape(-1, 1) + 0.01 * rng.standard_normal((N, 1))
    yd = np.stack([np.cos(x[:, 0]) * np.cos(x[:, 1]), -np.sin(x[:, 0]) * np.sin(x[:, 1])], 1)
    yd = yd + 0.01 * rng.standard_normal(yd.shape)
    m = GPR_TD(SquaredExponential())
    l0, _ = m.loss_and_grad(x, y, J, yd)
    m.fit(x, y, J, yd)
    assert m.optimize_results_.fun < l0 - 0.1
    xs = rng.uniform(-1.5, 1.5, (20, 2))
    yp, ydp = m.predict(xs, np.tile(np.eye(2)[None], (20, 1, 1)))
    truth = (np.sin(xs[:, 0]) * np.cos(xs[:, 1])).reshape(-1, 1)
    assert np.max(np.abs(yp - truth)) < 0.05  # targets + gradients pin the surface


def test_reference_notebook_derivative_block_kernel_on_the_device():
    """examples/derivatives_solak.ipynb cell 4 prints [[k, d1k], [d0k, d0d1k]](x0 = 0, x = linspace(-3, 3, 1000)) for the
    SE kernel at lengthscale 1 (tests/golden/notebook_anchors.json).  The 1-feature problem is embedded in 5 features /
    3 derivative variables (zero padding), the regime the other tests of gpxb_td_gram cover."""
    import json
    import os

    from gpx_b200 import ops

    a = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_anchors.json")))["derivatives_solak"]
    N1, N2, nf, nv = 7, 1000, 5, 3
    x1, x2 = np.zeros((N1, nf)), np.zeros((N2, nf))
    x1[:, 0] = np.linspace(0.0, 1.5, N1)  # row 0 is the notebook's x0 = 0
    x2[:, 0] = np.linspace(-3.0, 3.0, N2)
    J1, J2 = np.zeros((N1, nf, nv)), np.zeros((N2, nf, nv))
    J1[:, 0, 0] = 1.0
    J2[:, 0, 0] = 1.0
    out = host(ops.td_gram("se", dev(x1), dev(J1), dev(x2), dev(J2), 1.0))
    assert out.shape == (N1 + N1 * nv, N2 + N2 * nv)
    ref = R.td_A_lhs("se", x1, J1, x2, J2, 1.0, 0.0, 0.0, noise=False)
    assert np.max(np.abs(out - ref)) < 1e-12
    k, d1k = out[0, :N2], out[0, N2::nv]            # targets row of x0: [k | d1k]
    d0k, d01k = out[N1, :N2], out[N1, N2::nv]       # first derivative row of x0: [d0k | d0d1k]
    np.testing.assert_allclose(k[:3], a["row0_first3"], rtol=0, atol=6e-9)
    np.testing.assert_allclose(d1k[-3:], a["row0_last3"], rtol=0, atol=6e-9)
    np.testing.assert_allclose(d0k[:3], a["row1_first3"], rtol=0, atol=6e-9)
    np.testing.assert_allclose(d01k[-3:], a["row1_last3"], rtol=0, atol=6e-9)


@pytest.mark.parametrize("kind,N,nf,nv1,nv2,mean", [("se", 30, 6, 4, 3, "data"), ("m52", 24, 8, 2, 5, "zero"), ("se", 12, 7, 7, 1, "zero")])
def test_td_second_derivative_block(kind, N, nf, nv1, nv2, mean):
    """GPR_TD with jacobian_2 / y_derivs_2 / sigma_derivs2 (gpr_td.py:114-158) through the facade against the oracle's
    literal 3 x 3 block assembly: LML, gradient (Richardson central differences of the oracle), fit coefficients in
    the reference's order, contracted jacobians, predictions of all three outputs, and L-BFGS-B training."""
    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.mean_functions import data_mean, zero_mean
    from gpx_b200.models import GPR_TD
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    x, Jc, y, ydc = _data(N, nf, nv1 + nv2, 21)
    xs, Jsc, _, _ = _data(8, nf, nv1 + nv2, 22)
    Ja, Jb, ya, yb = Jc[:, :, :nv1].copy(), Jc[:, :, nv1:].copy(), ydc[:, :nv1].copy(), ydc[:, nv1:].copy()
    ell, st, sd, sd2 = 0.9 * np.sqrt(nf), 0.3, 0.2, 0.45
    mf_o = R.data_mean if mean == "data" else R.zero_mean
    par = lambda v: Parameter(v, True, Softplus(), GammaPrior())  # noqa: E731
    kern = SquaredExponential() if kind == "se" else Matern52()
    mk = lambda: GPR_TD(kern, mean_function=data_mean if mean == "data" else zero_mean,  # noqa: E731
                        kernel_params=dict(lengthscale=par(ell)), sigma_targets=par(st), sigma_derivs=par(sd),
                        sigma_derivs2=par(sd2))
    m = mk()
    f = lambda l, a, b, c2: R.td2_lml_dense(kind, x, y, Ja, Jb, ya, yb, l, a, b, c2, mf_o)  # noqa: E731
    ref = f(ell, st, sd, sd2)
    assert abs(m.log_marginal_likelihood(x, y, Ja, ya, jacobian_2=Jb, y_derivs_2=yb) - ref) < 1e-8 * abs(ref)
    loss, grads = m.loss_and_grad(x, y, Ja, ya, jacobian_2=Jb, y_derivs_2=yb)
    assert abs(-loss - ref) < 1e-8 * abs(ref)

    def rich(g, h=1e-3):
        fd = lambda hh: (g(hh) - g(-hh)) / (2 * hh)  # noqa: E731
        return (4 * fd(h / 2) - fd(h)) / 3

    g_ref = {("kernel_params", "lengthscale"): rich(lambda h: f(ell + h, st, sd, sd2)),
             ("sigma_targets",): rich(lambda h: f(ell, st + h, sd, sd2)),
             ("sigma_derivs",): rich(lambda h: f(ell, st, sd + h, sd2)),
             ("sigma_derivs2",): rich(lambda h: f(ell, st, sd, sd2 + h))}
    assert set(grads) == set(g_ref)
    for k, g in g_ref.items():
        assert abs(-grads[k] - g) < 1e-7 * max(abs(g), abs(ref)), (k, grads[k], g)
    m.fit(x, y, Ja, ya, jacobian_2=Jb, y_derivs_2=yb, minimize=False)
    c_ref, mu_ref = R.td2_fit_dense(kind, x, y, Ja, Jb, ya, yb, ell, st, sd, sd2, mf_o)
    assert abs(float(m.state.mu) - mu_ref) < 1e-14
    assert np.max(np.abs(m.state.c - c_ref)) < 1e-8 * np.max(np.abs(c_ref))  # the reference's row order
    assert np.max(np.abs(m.state.jaccoef_2 - np.einsum("sv,sfv->sf", c_ref[N + N * nv1:].reshape(N, nv2), Jb))) < 1e-8 * np.max(np.abs(c_ref))
    yp, d1, d2 = m.predict(xs, Jsc[:, :, :nv1].copy(), jacobian_2=Jsc[:, :, nv1:].copy())
    yr, d1r, d2r = R.td2_predict_dense(kind, x, Ja, Jb, xs, Jsc[:, :, :nv1], Jsc[:, :, nv1:], c_ref, mu_ref, ell)
    for got, want in ((yp, yr), (d1, d1r), (d2, d2r)):
        assert got.shape == want.shape and np.max(np.abs(got - want)) < 1e-8 * max(1.0, np.max(np.abs(want)))
    with pytest.raises(ValueError):
        m.predict(xs, Jsc[:, :, :nv1].copy())  # fitted with two blocks: jacobian_2 is required
    if kind == "se" and mean == "data":
        m2 = mk()
        l0, _ = m2.loss_and_grad(x, y, Ja, ya, jacobian_2=Jb, y_derivs_2=yb)
        m2.fit(x, y, Ja, ya, jacobian_2=Jb, y_derivs_2=yb, opt_kwargs=dict(options=dict(maxiter=15)))
        assert m2.optimize_results_.fun < l0


# ---- iterative paths (gpx/models/gpr_td.py:162-505, 562-747) ---------------------------------------------------------
@pytest.mark.parametrize("kind,N,nf,nv", [("se", 60, 6, 4), ("m52", 45, 9, 9), ("se", 24, 30, 30)])
@pytest.mark.parametrize("partitionable", [True, False])
def test_td_rpcholesky_pivots_bit_exact(kind, N, nf, nv, partitionable):
    from gpx_b200 import ops

    x, J, _, _ = _data(N, nf, nv, 3)
    ell, M = 0.9 * np.sqrt(nf), 25
    key = R.PRNGKey(2023)
    Fref, pref = R.rpcholesky_td(kind, x, J, M, ell, key, partitionable)
    F, piv = ops.rpcholesky_td(kind, dev(x), dev(J), M, ell, key, partitionable)
    assert np.array_equal(host(piv), pref)
    assert np.max(np.abs(host(F) - Fref)) < 1e-9 * max(1.0, np.max(np.abs(Fref)))


@pytest.mark.parametrize("kind,N,nf,nv,mean", [("se", 60, 6, 4, "data"), ("m52", 45, 9, 9, "zero")])
def test_td_iterative_operator_fit_and_lml(kind, N, nf, nv, mean):
    """The matrix-free block operator against the dense block matrix (through a Lanczos step on unit vectors), the PCG
    fit against the dense solve with both CG runs converged, and the iterative LML against the oracle's."""
    from gpx_b200 import ops
    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.mean_functions import data_mean, zero_mean
    from gpx_b200.models import GPR_TD
    from gpx_b200.models import gpr_td as T

    x, J, y, yd = _data(N, nf, nv, 4)
    ell, st, sd = 0.9 * np.sqrt(nf), 0.3, 0.2
    key = R.PRNGKey(2023)
    mfun = R.data_mean if mean == "data" else R.zero_mean
    n = N + N * nv
    A = R.td_A_lhs(kind, x, J, x, J, ell, st, sd)
    # operator: one Lanczos step gives alpha = v^T A v, beta = |A v - alpha v| for each start vector
    rng = np.random.default_rng(0)
    U = rng.standard_normal((n, 7))
    U /= np.linalg.norm(U, axis=0)
    alpha, beta, flag = ops.lanczos_td(kind, dev(x), dev(J), ell, st, sd, dev(U), 3)
    for p in range(U.shape[1]):
        Tref, _ = R.lanczos_tridiagonal(lambda v: A @ v, 3, U[:, p])
        assert np.max(np.abs(host(alpha)[p] - np.diag(Tref))) < 1e-9 * np.max(np.abs(np.diag(Tref)))
        assert np.max(np.abs(host(beta)[p, :2] - np.diag(Tref, 1))) < 1e-9 * np.max(np.abs(np.diag(Tref)))
    # fit
    c_ref, mu_ref, _ = R.td_fit_iter(kind, x, y, J, yd, ell, st, sd, 20, key, mfun, cg_tol=1e-13, cg_atol=0.0)
    c, mu, iters = ops.gpr_td_fit_iter(kind, dev(x), dev(J), dev(y), dev(yd), ell, st, sd, 20, key,
                                       1 if mean == "data" else 0, tol=1e-13, atol=0.0)
    assert abs(float(host(mu)[0]) - float(mu_ref)) < 1e-14
    assert np.max(np.abs(host(c) - c_ref)) < 1e-8 * np.max(np.abs(c_ref))
    c_d, _ = R.td_fit_dense(kind, x, y, J, yd, ell, st, sd, mfun)
    assert np.max(np.abs(host(c) - c_d)) < 1e-8 * np.max(np.abs(c_d))
    # with the reference's stopping rule: same iteration count as the oracle's CG
    _, _, it_ref = R.td_fit_iter(kind, x, y, J, yd, ell, st, sd, 20, key, mfun)
    _, _, it_dev = ops.gpr_td_fit_iter(kind, dev(x), dev(J), dev(y), dev(yd), ell, st, sd, 20, key,
                                       1 if mean == "data" else 0)
    assert abs(it_dev - it_ref) <= 1
    # iterative LML through the facade against the oracle (converged solves) and, loosely, the dense LML
    K = SquaredExponential() if kind == "se" else Matern52()
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    P = lambda v: Parameter(v, True, Softplus(), GammaPrior())  # noqa: E731
    model = GPR_TD(K, mean_function=data_mean if mean == "data" else zero_mean, kernel_params=dict(lengthscale=P(ell)),
                   sigma_targets=P(st), sigma_derivs=P(sd))
    ref = R.td_lml_iter(kind, x, y, J, yd, ell, st, sd, 6, 12, key, 20, key, mfun, cg_tol=1e-13, cg_atol=0.0)
    out = T._lml_iter(model.state, x, y, J, yd, 6, 12, key, 20, key, tol=1e-13, atol=0.0)
    assert abs(out - ref) < 1e-8 * abs(ref)
    lml_facade = model.log_marginal_likelihood(x, y, J, yd, iterative=True, num_evals=6, num_lanczos=12, key_lanczos=key,
                                               n_pivots=20, key_precond=key)
    assert abs(lml_facade - ref) < 1e-6 * abs(ref)
    # fit(iterative=True, minimize=False) + predict(iterative=True) reproduce the dense model's predictions
    model.fit(x, y, J, yd, iterative=True, minimize=False, n_pivots=20, key_precond=key)
    xs, Js, _, _ = _data(11, nf, nv, 9)
    yp, dp = model.predict(xs, Js, iterative=True)
    y_ref, d_ref = R.td_predict_dense(kind, x, J, xs, Js, c_d, mfun(y), ell)
    assert np.max(np.abs(yp - y_ref)) < 1e-4 * max(1.0, np.max(np.abs(y_ref)))  # CG stopped at tol 1e-5
    assert np.max(np.abs(dp - d_ref)) < 1e-4 * max(1.0, np.max(np.abs(d_ref)))


@pytest.mark.parametrize("kind,N,nf,nv1,nv2", [("se", 30, 6, 4, 3), ("m52", 24, 8, 2, 5)])
def test_td_iterative_second_derivative_block(kind, N, nf, nv1, nv2):
    """The iterative GPR_TD paths with jacobian_2 / y_derivs_2: rpcholesky_TD pivots drawn over the reference's row
    order [targets; block 1; block 2] (bit-exact), PCG fit and iterative LML against the oracle with converged solves."""
    from gpx_b200 import ops
    from gpx_b200.kernels import Matern52, SquaredExponential
    from gpx_b200.models import GPR_TD
    from gpx_b200.models import gpr_td as T
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    x, Ja, y, ya = _data(N, nf, nv1, 5)
    _, Jb, _, yb = _data(N, nf, nv2, 6)
    ell, st, sd, sd2 = 0.9 * np.sqrt(nf), 0.3, 0.2, 0.25
    key = R.PRNGKey(2023)
    Jc = np.concatenate([Ja, Jb], axis=2)
    for part in (True, False):
        Fref, pref = R.rpcholesky_td2(kind, x, Ja, Jb, 20, ell, key, part)
        F, piv = ops.rpcholesky_td(kind, dev(x), dev(Jc), 20, ell, key, part, nv_first=nv1)
        assert np.array_equal(host(piv), pref)
        assert np.max(np.abs(host(F) - Fref)) < 1e-9 * max(1.0, np.max(np.abs(Fref)))
    P = lambda v: Parameter(v, True, Softplus(), GammaPrior())  # noqa: E731
    K = SquaredExponential() if kind == "se" else Matern52()
    model = GPR_TD(K, kernel_params=dict(lengthscale=P(ell)), sigma_targets=P(st), sigma_derivs=P(sd), sigma_derivs2=P(sd2))
    c_ref, mu_ref, _ = R.td2_fit_iter(kind, x, y, Ja, Jb, ya, yb, ell, st, sd, sd2, 20, key, R.zero_mean, cg_tol=1e-13,
                                      cg_atol=0.0)
    c, mu, _ = T._fit_iter(model.state, x, y, Ja, ya, 20, key, Jb, yb, tol=1e-13, atol=0.0)
    c_np = T._to_ref_order(host(c), N, nv1, nv2)
    assert np.max(np.abs(c_np - c_ref)) < 1e-8 * np.max(np.abs(c_ref))
    ref = R.td2_lml_iter(kind, x, y, Ja, Jb, ya, yb, ell, st, sd, sd2, 6, 12, key, 20, key, R.zero_mean, cg_tol=1e-13,
                         cg_atol=0.0)
    out = T._lml_iter(model.state, x, y, Ja, ya, 6, 12, key, 20, key, Jb, yb, tol=1e-13, atol=0.0)
    assert abs(out - ref) < 1e-8 * abs(ref)
    # facade: fit(iterative=True, minimize=False) stores c in the reference's order; predictions match the dense model
    model.fit(x, y, Ja, ya, jacobian_2=Jb, y_derivs_2=yb, iterative=True, minimize=False, n_pivots=20)
    c_d, mu_d = R.td2_fit_dense(kind, x, y, Ja, Jb, ya, yb, ell, st, sd, sd2, R.zero_mean)
    assert np.max(np.abs(np.asarray(model.state.c) - c_d)) < 1e-3 * np.max(np.abs(c_d))  # CG stopped at tol 1e-5
    xs, Jsa, _, _ = _data(7, nf, nv1, 8)
    _, Jsb, _, _ = _data(7, nf, nv2, 9)
    yp, d1, d2 = model.predict(xs, Jsa, jacobian_2=Jsb, iterative=True)
    y_ref, d1_ref, d2_ref = R.td2_predict_dense(kind, x, Ja, Jb, xs, Jsa, Jsb, c_d, mu_d, ell)
    assert np.max(np.abs(yp - y_ref)) < 1e-3 * max(1.0, np.max(np.abs(y_ref)))
    assert np.max(np.abs(d2 - d2_ref)) < 1e-3 * max(1.0, np.max(np.abs(d2_ref)))
