"""GPU parity: Sum / Prod composite kernels in the dense exact-GPR paths against the oracle's
literal composition; gradients against Richardson-extrapolated central differences of the oracle's LML."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as R  # noqa: E402


def _data(N, D, seed, T=1):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, D))
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, T))
    return x, y


def _ell_leaves(kernel, params):
    """Leaves that carry a lengthscale (Linear has no parameters)."""
    return [lf for lf in kernel.leaves(params) if "lengthscale" in lf[2]]


def _model(kernel, ells, sigma):
    from gpx_b200.models import GPR
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    m = GPR(kernel, sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    st = m.state
    for (path, _, _), ell in zip(_ell_leaves(kernel, st.params["kernel_params"]), ells):
        st = st.with_param_value(("kernel_params",) + path + ("lengthscale",), np.float64(ell))
    m.state = st
    return m


CASES = {
    "sum": (lambda K: K.SquaredExponential() + K.Matern52(),
            lambda e: ("sum", ("leaf", "se", e[0], None), ("leaf", "m52", e[1], None)), [1.7, 0.9]),
    "prod": (lambda K: K.Matern32() * K.SquaredExponential(),
             lambda e: ("prod", ("leaf", "m32", e[0], None), ("leaf", "se", e[1], None)), [2.1, 1.4]),
    "nested": (lambda K: K.SquaredExponential(active_dims=[0, 1]) + K.Matern52() * K.Matern12(active_dims=[2]),
               lambda e: ("sum", ("leaf", "se", e[0], [0, 1]),
                          ("prod", ("leaf", "m52", e[1], None), ("leaf", "m12", e[2], [2]))), [1.2, 2.5, 1.9]),
    "linear": (lambda K: K.Linear(active_dims=[0, 1]) + K.Linear() * K.Matern52(active_dims=[1, 2]) + K.SquaredExponential(),
               lambda e: ("sum", ("sum", ("leaf", "linear", None, [0, 1]),
                                  ("prod", ("leaf", "linear", None, None), ("leaf", "m52", e[0], [1, 2]))),
                          ("leaf", "se", e[1], None)), [1.6, 2.2]),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("T", [1, 2])
def test_composite_kernel_lml_gradient_fit_predict(name, T):
    import gpx_b200.kernels as K

    mk, spec_of, ells = CASES[name]
    kernel = mk(K)
    N, D, sigma = 300, 3, 0.25
    x, y = _data(N, D, 7, T)
    xs, _ = _data(40, D, 8)
    m = _model(kernel, ells, sigma)
    Kd = kernel.k(x, xs, m.state.params["kernel_params"])
    assert np.max(np.abs(Kd - R.composite_kernel(spec_of(ells), x, xs))) < 1e-13
    f = lambda e, s: R.composite_lml_dense(spec_of(e), x, y, s)  # noqa: E731
    ref = f(ells, sigma)
    assert abs(m.log_marginal_likelihood(x, y) - ref) < 1e-8 * abs(ref)
    loss, grads = m.loss_and_grad(x, y)
    assert abs(-loss - ref) < 1e-8 * abs(ref)

    def rich(g, h=1e-3):
        fd = lambda hh: (g(hh) - g(-hh)) / (2 * hh)  # noqa: E731
        return (4 * fd(h / 2) - fd(h)) / 3

    leaves = _ell_leaves(kernel, m.state.params["kernel_params"])
    for i, (path, _, _) in enumerate(leaves):
        def g(h, i=i):
            e = list(ells)
            e[i] += h
            return f(e, sigma)
        gi = rich(g)
        assert abs(-grads[("kernel_params",) + path + ("lengthscale",)] - gi) < 1e-7 * max(abs(gi), abs(ref)), (name, i)
    gs = rich(lambda h: f(ells, sigma + h))
    assert abs(-grads[("sigma",)] - gs) < 1e-7 * max(abs(gs), abs(ref))
    m.fit(x, y, minimize=False)
    A = R.composite_kernel(spec_of(ells), x, x) + (sigma**2 + 1e-10) * np.eye(N)
    mu = np.mean(y)
    c_ref = np.linalg.solve(A, y - mu)
    assert np.max(np.abs(m.c_ - c_ref)) < 1e-8 * np.max(np.abs(c_ref))
    mean, cov = m.predict(xs, full_covariance=True)
    Ks = R.composite_kernel(spec_of(ells), x, xs)
    assert np.max(np.abs(mean - (mu + Ks.T @ c_ref))) < 1e-8 * max(1.0, np.max(np.abs(mean)))
    cov_ref = R.composite_kernel(spec_of(ells), xs, xs) - Ks.T @ np.linalg.solve(A, Ks)
    assert np.max(np.abs(cov - cov_ref)) < 1e-8


def test_composite_kernel_fit_minimize():
    import gpx_b200.kernels as K
    from gpx_b200.models import GPR

    x, y = _data(200, 2, 9)
    m = GPR(K.SquaredExponential() + K.Matern32())
    l0, _ = m.loss_and_grad(x, y)
    m.fit(x, y)
    assert m.optimize_results_.fun < l0 - 0.05
    with pytest.raises(NotImplementedError):  # the iterative LML of a composite kernel exists, its gradient does not
        m.loss_and_grad(x, y, iterative=True)


def test_reference_notebook_composite_kernel_matrices_on_the_device():
    """examples/kernel_operations.ipynb cells 6, 10, 13, 16: the printed corners of Linear / Matern52 Sum / Prod kernel
    matrices (tests/golden/notebook_anchors.json) from the device path (DMMA GEMM + fused Gram + gpxb_elementwise)."""
    import json
    import os

    import gpx_b200.kernels as K
    from test_oracle_cpu import (KERNEL_OPERATIONS_SPECS, check_kernel_operations_corners,
                                 notebook_kernel_operations_data)

    anchors = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_anchors.json")))["kernel_operations"]
    X = notebook_kernel_operations_data()
    a4, b6 = [0, 1, 2, 3], [4, 5, 6, 7, 8, 9]
    kernels = {
        "linear_plus_linear_times_matern52": K.Linear(active_dims=a4) + K.Linear(active_dims=a4) * K.Matern52(active_dims=b6),
        "linear_plus_linear_times_matern52_same_dims": K.Linear(active_dims=b6) + K.Linear(active_dims=a4) * K.Matern52(active_dims=a4),
    }
    for name, kernel in kernels.items():
        Kd = kernel.k(X, X, kernel.default_params())  # every lengthscale 1.0, as in the notebook
        check_kernel_operations_corners(Kd, anchors[name])
        assert np.max(np.abs(Kd - R.composite_kernel(KERNEL_OPERATIONS_SPECS[name], X, X))) < 1e-13 * np.max(np.abs(Kd))


def test_linear_kernel_alone_in_the_dense_gpr_path():
    import gpx_b200.kernels as K
    from gpx_b200.models import GPR
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    x, y = _data(150, 4, 11)
    sigma = 0.4
    m = GPR(K.Linear(), sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    spec = ("leaf", "linear", None, None)
    ref = R.composite_lml_dense(spec, x, y, sigma)
    assert abs(m.log_marginal_likelihood(x, y) - ref) < 1e-8 * abs(ref)
    loss, grads = m.loss_and_grad(x, y)
    h = 1e-4
    gs = (R.composite_lml_dense(spec, x, y, sigma + h) - R.composite_lml_dense(spec, x, y, sigma - h)) / (2 * h)
    assert set(grads) == {("sigma",)} and abs(-grads[("sigma",)] - gs) < 1e-6 * abs(gs)
    m.fit(x, y, minimize=False)
    A = x @ x.T + (sigma**2 + 1e-10) * np.eye(150)
    c_ref = np.linalg.solve(A, y - np.mean(y))
    assert np.max(np.abs(m.c_ - c_ref)) < 1e-8 * np.max(np.abs(c_ref))
    xs, _ = _data(10, 4, 12)
    assert np.max(np.abs(m.predict(xs) - (np.mean(y) + xs @ x.T @ c_ref))) < 1e-8


def _sgpr_model(kernel, ells, sigma, x_locs):
    from gpx_b200.models import SGPR
    from gpx_b200.models.sgpr import locs_parameter
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    m = SGPR(kernel, x_locs=locs_parameter(x_locs), sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    st = m.state
    for (path, _, _), ell in zip(_ell_leaves(kernel, st.params["kernel_params"]), ells):
        st = st.with_param_value(("kernel_params",) + path + ("lengthscale",), np.float64(ell))
    m.state = st
    return m


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("T", [1, 2])
def test_composite_kernel_sgpr_lml_gradient_fit_predict(name, T):
    """SGPR (Projected Processes) with a Sum / Prod / Linear kernel against the oracle's literal N x N form
    (oracle/ref_numpy.py: sgpr_composite_*); gradients against Richardson-extrapolated central differences."""
    import gpx_b200.kernels as K

    mk, spec_of, ells = CASES[name]
    kernel = mk(K)
    N, M, D, sigma = 400, 24, 3, 0.3
    x, y = _data(N, D, 17, T)
    xs, _ = _data(30, D, 18)
    xl = x[:: N // M][:M].copy()
    m = _sgpr_model(kernel, ells, sigma, xl)
    f = lambda e, s: R.sgpr_composite_lml_dense_nxn(spec_of(e), xl, x, y, s)  # noqa: E731
    ref = f(ells, sigma)
    assert abs(m.log_marginal_likelihood(x, y) - ref) < 1e-8 * abs(ref)
    loss, grads = m.loss_and_grad(x, y)
    assert abs(-loss - ref) < 1e-8 * abs(ref)
    assert ("x_locs",) not in grads

    def rich(g, h=1e-3):
        fd = lambda hh: (g(hh) - g(-hh)) / (2 * hh)  # noqa: E731
        return (4 * fd(h / 2) - fd(h)) / 3

    for i, (path, _, _) in enumerate(_ell_leaves(kernel, m.state.params["kernel_params"])):
        def g(h, i=i):
            e = list(ells)
            e[i] += h
            return f(e, sigma)
        gi = rich(g)
        assert abs(-grads[("kernel_params",) + path + ("lengthscale",)] - gi) < 1e-6 * max(abs(gi), abs(ref)), (name, i)
    gs = rich(lambda h: f(ells, sigma + h))
    assert abs(-grads[("sigma",)] - gs) < 1e-6 * max(abs(gs), abs(ref))
    m.fit(x, y, minimize=False)
    c_ref, mu_ref = R.sgpr_composite_fit_dense(spec_of(ells), xl, x, y, sigma)
    assert np.max(np.abs(m.c_ - c_ref)) < 1e-6 * np.max(np.abs(c_ref))  # normal equations: cond^2
    mean, cov = m.predict(xs, full_covariance=True)
    mean_ref, cov_ref = R.sgpr_composite_predict_dense(spec_of(ells), xl, xs, c_ref, mu_ref, sigma)
    assert np.max(np.abs(mean - mean_ref)) < 1e-7 * max(1.0, np.max(np.abs(mean_ref)))
    assert np.max(np.abs(cov - cov_ref)) < 1e-7 * max(1.0, np.max(np.abs(cov_ref)))
    assert np.max(np.abs(m.predict(xs) - mean_ref)) < 1e-7 * max(1.0, np.max(np.abs(mean_ref)))


def test_composite_kernel_sgpr_fit_minimize_and_limits():
    import gpx_b200.kernels as K
    from gpx_b200.models import SGPR
    from gpx_b200.models.sgpr import locs_parameter

    x, y = _data(300, 2, 19)
    xl = x[::20].copy()
    m = SGPR(K.SquaredExponential() + K.Matern32(), x_locs=locs_parameter(xl))
    l0, _ = m.loss_and_grad(x, y)
    m.fit(x, y)
    assert m.optimize_results_.fun < l0 - 0.05
    with pytest.raises(NotImplementedError):
        m.log_marginal_likelihood(x, y, iterative=True)
    mt = SGPR(K.SquaredExponential() + K.Matern32(), x_locs=locs_parameter(xl, trainable=True))
    with pytest.raises(NotImplementedError):
        mt.loss_and_grad(x, y)


@pytest.mark.parametrize("name", list(CASES))
def test_fused_composite_gram_kernel_entries(name):
    """gpxb_gram_composite (every leaf and operator of the tree in one tile-level pass) against the oracle's literal
    composition: symmetric with the exact diagonal, rectangular with ragged tiles, lower-only with a diagonal shift;
    and against the component-matrix path it replaces (same numbers to rounding)."""
    import gpx_b200.kernels as K
    from gpx_b200 import ops

    mk, spec_of, ells = CASES[name]
    kernel = mk(K)
    m = _model(kernel, ells, 0.3)
    kp = m.state.params["kernel_params"]
    spec = spec_of(ells)
    rng = np.random.default_rng(3)
    x1, x2 = rng.standard_normal((333, 3)), rng.standard_normal((150, 3))
    d1, d2 = torch.as_tensor(x1).cuda(), torch.as_tensor(x2).cuda()
    Ks = kernel.k_device(d1, d1, kp).cpu().numpy()
    assert np.max(np.abs(Ks - R.composite_kernel(spec, x1, x1))) < 1e-12 * max(1.0, np.max(np.abs(Ks)))
    Kr = kernel.k_device(d1, d2, kp).cpu().numpy()
    assert np.max(np.abs(Kr - R.composite_kernel(spec, x1, x2))) < 1e-12 * max(1.0, np.max(np.abs(Kr)))
    # the fused kernel is what ran
    leaves, prog = kernel.program(kp)
    assert ops.gram_composite_fits([3 if k.active_dims is None else len(k.active_dims) for k, _ in leaves])
    x1l = [d1 if k.active_dims is None else d1[:, k.active_dims].contiguous() for k, _ in leaves]
    kinds = [k.kind for k, _ in leaves]
    el = [1.0 if k.kind == "linear" else k.lengthscale(p) for k, p in leaves]
    Kl = ops.gram_composite(kinds, el, x1l, None, prog, diag_add=0.5, lower_only=True).cpu().numpy()
    ref = np.tril(R.composite_kernel(spec, x1, x1) + 0.5 * np.eye(333))
    assert np.max(np.abs(Kl - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref)))
    assert np.all(np.triu(Kl, 1) == 0.0)
    # component-matrix path (what round 1 did) gives the same numbers
    mats = [k.k_device(d1, d1, p).cpu().numpy() for k, p in leaves]
    st = []
    for t in prog:
        if t >= 0:
            st.append(mats[t])
        else:
            b, a = st.pop(), st.pop()
            st.append(a + b if t == ops.OP_SUM else a * b)
    assert np.max(np.abs(Ks - st[0])) < 1e-12 * max(1.0, np.max(np.abs(Ks)))


def test_fused_composite_gram_eight_leaves_and_fallback():
    import gpx_b200.kernels as K
    from gpx_b200 import ops

    rng = np.random.default_rng(5)
    x = rng.standard_normal((200, 6))
    d = torch.as_tensor(x).cuda()
    # balanced tree of 8 leaves: stack depth 4
    mk = lambda i: (K.SquaredExponential if i % 2 else K.Matern52)(active_dims=[i % 6, (i + 1) % 6])  # noqa: E731
    kernel = ((mk(0) + mk(1)) * (mk(2) + mk(3))) + ((mk(4) * mk(5)) + (mk(6) * mk(7)))
    kp = kernel.default_params()
    leaves, prog = kernel.program(kp)
    assert len(leaves) == 8
    spec_leaf = lambda i: ("leaf", "se" if i % 2 else "m52", 1.0, [i % 6, (i + 1) % 6])  # noqa: E731
    spec = ("sum", ("prod", ("sum", spec_leaf(0), spec_leaf(1)), ("sum", spec_leaf(2), spec_leaf(3))),
            ("sum", ("prod", spec_leaf(4), spec_leaf(5)), ("prod", spec_leaf(6), spec_leaf(7))))
    Kd = kernel.k_device(d, d, kp).cpu().numpy()
    assert np.max(np.abs(Kd - R.composite_kernel(spec, x, x))) < 1e-12 * np.max(np.abs(Kd))
    # leaves too wide for shared memory: the component-matrix path takes over, same numbers
    xw = rng.standard_normal((150, 120))
    dw = torch.as_tensor(xw).cuda()
    kw = K.SquaredExponential() + K.Matern52()
    assert not ops.gram_composite_fits([120, 120])
    kpw = kw.default_params()
    ref = R.composite_kernel(("sum", ("leaf", "se", 1.0, None), ("leaf", "m52", 1.0, None)), xw, xw)
    assert np.max(np.abs(kw.k_device(dw, dw, kpw).cpu().numpy() - ref)) < 1e-12


# ---- composite kernels outside the dense paths (SURVEY 8f N3 remainder) -----------------------------------------------
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("partitionable", [True, False])
def test_composite_rpcholesky_pivots_bit_exact(name, partitionable):
    import gpx_b200.kernels as K
    from gpx_b200 import ops
    from gpx_b200.models import _composite

    mk, spec_of, ells = CASES[name]
    kernel = mk(K)
    m = _model(kernel, ells, 0.3)
    spec = spec_of(ells)
    x, _ = _data(700, 3, 8)
    key = R.PRNGKey(2023)
    Fref, pref = R.rpcholesky_composite(spec, x, 40, key, partitionable)
    d = torch.as_tensor(x).cuda()
    kinds, el, xl, prog = _composite._leaf_desc(m.state, d)
    F, piv = ops.rpcholesky_composite(kinds, el, xl, prog, 40, key, partitionable)
    assert np.array_equal(piv.cpu().numpy(), pref)
    assert np.max(np.abs(F.cpu().numpy() - Fref)) < 1e-9 * max(1.0, np.max(np.abs(Fref)))


@pytest.mark.parametrize("name", ["sum", "nested", "linear"])
def test_composite_iterative_fit_and_lml(name):
    """GPR(Sum / Prod kernel).fit(iterative=True) and log_marginal_likelihood(iterative=True): PCG with the composite
    RPCholesky preconditioner and SLQ on the row-chunked composite operator against the oracle (converged solves: 1e-8)."""
    import gpx_b200.kernels as K
    from gpx_b200.models import _composite

    mk, spec_of, ells = CASES[name]
    kernel = mk(K)
    sigma = 0.4
    m = _model(kernel, ells, sigma)
    spec = spec_of(ells)
    x, y = _data(600, 3, 9)
    key = R.PRNGKey(2023)
    c_ref, mu_ref, it_ref = R.composite_fit_iter(spec, x, y, sigma, 25, key, cg_tol=1e-13, cg_atol=0.0)
    c, mu, it = _composite.fit_iter(m.state, x, y, 25, key, tol=1e-13, atol=0.0, return_iters=True)
    assert abs(float(mu) - float(mu_ref)) < 1e-14
    assert np.max(np.abs(c - c_ref)) < 1e-8 * np.max(np.abs(c_ref))
    ref = R.composite_lml_iter(spec, x, y, sigma, 6, 12, key, 25, key, cg_tol=1e-13, cg_atol=0.0)
    out = _composite.lml_iter(m.state, x, y, 6, 12, key, 25, key, tol=1e-13, atol=0.0)
    assert abs(out - ref) < 1e-8 * abs(ref)
    # the public entries with the reference's stopping rule
    _, _, it_ref5 = R.composite_fit_iter(spec, x, y, sigma, 25, key)
    _, _, it5 = _composite.fit_iter(m.state, x, y, 25, key, return_iters=True)
    assert abs(it5 - it_ref5) <= 1
    lml = m.log_marginal_likelihood(x, y, iterative=True, num_evals=6, num_lanczos=12, key_lanczos=key, n_pivots=25,
                                    key_precond=key)
    assert abs(lml - ref) < 1e-6 * abs(ref)
    m.fit(x, y, iterative=True, minimize=False, n_pivots=25)  # preconditioner key: gpxargs.key_precond, as in the reference
    xs, _ = _data(40, 3, 10)
    pred = m.predict(xs)
    pref = float(mu_ref) + R.composite_kernel(spec, xs, x) @ c_ref
    assert np.max(np.abs(pred - pref)) < 1e-4 * max(1.0, np.max(np.abs(pref)))


def test_sgpr_rpchol_with_a_composite_kernel():
    """SGPR_RPChol(Sum kernel): landmarks drawn by RPCholesky of the composite kernel, then the dense composite SGPR."""
    import gpx_b200.kernels as K
    from gpx_b200.models import SGPR_RPChol
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    mk, spec_of, ells = CASES["nested"]
    kernel = mk(K)
    spec = spec_of(ells)
    x, y = _data(800, 3, 11)
    key = R.PRNGKey(7)
    M, sigma = 30, 0.3
    model = SGPR_RPChol(kernel, n_locs=M, key=key, sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    st = model.state
    for (path, _, _), ell in zip(_ell_leaves(kernel, st.params["kernel_params"]), ells):
        st = st.with_param_value(("kernel_params",) + path + ("lengthscale",), np.float64(ell))
    model.state = st
    _, piv = R.rpcholesky_composite(spec, x, M, key)
    xl = x[piv]
    ref = R.sgpr_composite_lml_dense_nxn(spec, xl, x, y, sigma)
    out = model.log_marginal_likelihood(x, y)
    assert abs(out - ref) < 1e-8 * abs(ref)
    model.fit(x, y, minimize=False)
    assert np.array_equal(np.asarray(model.state.pivots), piv)
    c_ref, mu_ref = R.sgpr_composite_fit_dense(spec, xl, x, y, sigma)
    xs, _ = _data(50, 3, 12)
    pred = model.predict(xs)
    pref, _ = R.sgpr_composite_predict_dense(spec, xl, xs, c_ref, mu_ref, sigma)
    assert np.max(np.abs(pred - pref)) < 1e-6 * max(1.0, np.max(np.abs(pref)))


def test_composite_lanczos_breakdown_branch():
    """The breakdown restart (gpx/lanczos.py:92-108) lives in the shared Lanczos core: here through the composite
    operator, on a kernel matrix that is exactly block diagonal (two far-apart groups of points)."""
    import gpx_b200.kernels as K
    from gpx_b200 import ops
    from gpx_b200.models import _composite

    rng = np.random.default_rng(12)
    D, n1, n2 = 3, 5, 35
    x = np.vstack([rng.standard_normal((n1, D)), 100.0 + rng.standard_normal((n2, D))])
    N = n1 + n2
    U = np.zeros((N, 2))
    U[:n1, 0] = rng.standard_normal(n1)
    U[:, 1] = rng.choice([-1.0, 1.0], N)
    U /= np.linalg.norm(U, axis=0)
    kernel = K.SquaredExponential() + K.Matern52()
    ells, sigma, m = [1.0, 1.3], 0.3, 16
    model = _model(kernel, ells, sigma)
    spec = ("sum", ("leaf", "se", ells[0], None), ("leaf", "m52", ells[1], None))
    A = R.composite_kernel(spec, x, x) + (sigma**2 + 1e-10) * np.eye(N)
    d = torch.as_tensor(x).cuda()
    kinds, el, xl, prog = _composite._leaf_desc(model.state, d)
    key = R.PRNGKey(7)
    r0 = ops.gpr_iter_composite(kinds, el, xl, prog, sigma, U=torch.as_tensor(U).cuda(), m=m)
    assert int(r0["flag"].cpu()[0]) == 1
    r = ops.gpr_iter_composite(kinds, el, xl, prog, sigma, U=torch.as_tensor(U).cuda(), m=m, restart_key=key)
    assert int(r["flag"].cpu()[0]) == 0
    alpha, beta = r["alpha"].cpu().numpy(), r["beta"].cpu().numpy()
    hit = False
    for p in range(2):
        T, _ = R.lanczos_tridiagonal(lambda v: A @ v, m, U[:, p], key=key)
        b_ref = np.diag(T, 1)
        hit = hit or bool(np.any(b_ref <= 1e-12))
        assert np.max(np.abs(alpha[p] - np.diag(T))) < 1e-8 * np.max(np.abs(np.diag(T)))
        assert np.max(np.abs(beta[p, : m - 1] - b_ref)) < 1e-8 * np.max(b_ref)
    assert hit
