"""Host-logic tests (no GPU): the facade modules that are pure compositions of device calls are run with the CPU
stand-ins of tests/_fake_ops.py, through the very assertions the GPU parity tests make.  This checks the glue --
adjoint algebra, lower-triangle conventions, parameter paths -- not the kernels (the `-m gpu` tests do that)."""
import pytest

torch = pytest.importorskip("torch")

import test_gpu_composite as G  # noqa: E402  (its tests are plain functions; the gpu mark stays on that module)
import test_gpu_derivs as DV  # noqa: E402
import test_gpu_facade as F  # noqa: E402
import test_gpu_golden as GF  # noqa: E402
import test_gpu_td as TD  # noqa: E402
from _fake_ops import patched  # noqa: E402


@pytest.mark.parametrize("name", list(G.CASES))
@pytest.mark.parametrize("T", [1, 2])
def test_composite_dense_gpr_glue(name, T):
    with patched():
        G.test_composite_kernel_lml_gradient_fit_predict(name, T)


def test_composite_notebook_anchor_glue():
    with patched():
        G.test_reference_notebook_composite_kernel_matrices_on_the_device()


def test_linear_kernel_glue():
    with patched():
        G.test_linear_kernel_alone_in_the_dense_gpr_path()


@pytest.mark.parametrize("name", list(G.CASES))
@pytest.mark.parametrize("T", [1, 2])
def test_composite_sgpr_glue(name, T):
    with patched():
        G.test_composite_kernel_sgpr_lml_gradient_fit_predict(name, T)


def test_composite_sgpr_minimize_glue():
    with patched():
        G.test_composite_kernel_sgpr_fit_minimize_and_limits()


# The GPX-style facade (GPR / SGPR / SGPR_RPChol: parameters, bijector chain rule, L-BFGS-B driver, state updates,
# landmark bookkeeping) with the fused entry points replaced by the oracle's restatement of each.
@pytest.mark.parametrize("test", [
    F.test_lml_and_loss_grad_match_oracle,
    F.test_fit_minimize_follows_oracle_optimiser,
    F.test_kernel_call_and_active_dims,
    F.test_sgpr_and_sgpr_rpchol_facades,
    F.test_iterative_predict_matches_dense,
    F.test_reference_notebook_anchors_through_the_facade,
    F.test_reference_notebook_anchor_sgpr_rpchol_through_the_facade,
], ids=lambda f: f.__name__)
def test_facade_glue(test):
    with patched():
        test()


@pytest.mark.parametrize("kind,N,nf,nv1,nv2,mean", [("se", 30, 6, 4, 3, "data"), ("m52", 24, 8, 2, 5, "zero")])
def test_gpr_td_second_block_glue(kind, N, nf, nv1, nv2, mean):
    """Concatenation of the two jacobians, the permutation of `c` to the reference's row order and back, the
    contracted jacobians and the three-way split of the predictions."""
    with patched():
        TD.test_td_second_derivative_block(kind, N, nf, nv1, nv2, mean)


def test_gpr_td_single_block_glue():
    with patched():
        TD.test_td_fit_minimize_lowers_the_loss()


# The fixture-based device tests, dry-run: the facade calls and the bookkeeping of tests/test_gpu_golden.py itself.
def test_golden_fixture_tests_glue():
    with patched():
        GF.test_exact_gpr_against_the_fixture()
        GF.test_rpcholesky_pivots_against_the_fixture_bit_exact(True, "pivots_part")
        GF.test_rpcholesky_pivots_against_the_fixture_bit_exact(False, "pivots_legacy")
        GF.test_sgpr_composite_and_td_against_the_fixture()


# Derivative-trained models through the facade: (sample, variable) flattening, contracted jacobians, the fast
# d01kjc / d0kjc prediction routes, full covariances, the L-BFGS-B driver on the derivative loss.
@pytest.mark.parametrize("test", [
    DV.test_facade_fit_and_predict_derivs,
    DV.test_facade_fit_derivs_minimize_follows_oracle_optimiser,
], ids=lambda f: f.__name__)
def test_derivs_facade_glue(test):
    with patched():
        test()


@pytest.mark.parametrize("kind", ["se", "m52"])
def test_sgpr_derivs_facade_glue(kind):
    with patched():
        DV.test_sgpr_fit_and_predict_on_derivatives(kind)


@pytest.mark.parametrize("kind", ["se", "m52"])
def test_kernel_derivative_methods_glue(kind):
    """Kernel.d0k / d1k / d01k / d0kj / d1kj / d01kj / d0kjc / d01kjc (gpx/kernels/kernels.py:1433-1488): argument order,
    block extraction from the [[k, d1kj], [d0kj, d01kj]] device matrix, identity jacobians of the jacobian-free forms,
    active_dims restriction of the jacobian rows, and the shapes the reference's kernelizers return."""
    import numpy as np

    import gpx_b200.kernels as K
    from oracle import ref_numpy as R

    rng = np.random.default_rng(0)
    n1, n2, nf, nv1, nv2 = 7, 5, 4, 3, 2
    x1, x2 = rng.standard_normal((n1, nf)), rng.standard_normal((n2, nf))
    J1, J2 = rng.standard_normal((n1, nf, nv1)), rng.standard_normal((n2, nf, nv2))
    jc = rng.standard_normal((n1, nf))
    I1, I2 = np.tile(np.eye(nf)[None], (n1, 1, 1)), np.tile(np.eye(nf)[None], (n2, 1, 1))
    ell = 1.7
    cls = K.SquaredExponential if kind == "se" else K.Matern52
    k = cls()
    p = {"lengthscale": k.default_params()["lengthscale"].update(dict(value=ell))}
    with patched():
        checks = [
            (k.d01kj(x1, x2, p, J1, J2), R.d01kj(kind, x1, x2, ell, J1, J2), (n1 * nv1, n2 * nv2)),
            (k.d01kjc(x1, x2, p, jc, J2), R.d01kjc(kind, x1, x2, ell, jc, J2), (n1, n2 * nv2)),
            (k.d0kj(x1, x2, p, J1), R.d0kj(kind, x1, x2, ell, J1), (n1 * nv1, n2)),
            (k.d1kj(x1, x2, p, J2), R.d0kj(kind, x2, x1, ell, J2).T, (n1, n2 * nv2)),
            (k.d0kjc(x1, x2, p, jc), R.d0kjc(kind, x1, x2, ell, jc), (n1, n2)),
            (k.d0k(x1, x2, p), R.d0kj(kind, x1, x2, ell, I1), (n1 * nf, n2)),
            (k.d1k(x1, x2, p), R.d0kj(kind, x2, x1, ell, I2).T, (n1, n2 * nf)),
            (k.d01k(x1, x2, p), R.d01kj(kind, x1, x2, ell, I1, I2), (n1 * nf, n2 * nf)),
        ]
        for got, want, shape in checks:
            assert got.shape == shape and np.max(np.abs(got - want)) < 1e-12
        dims = [0, 2]
        ka = cls(active_dims=dims)
        got = ka.d01kj(x1, x2, p, J1, J2)
        want = R.d01kj(kind, x1[:, dims], x2[:, dims], ell, J1[:, dims, :], J2[:, dims, :])
        assert np.max(np.abs(got - want)) < 1e-12
        # jacobian-free forms with active_dims: exact zeros for the inactive features (test_kernels.py:551-580)
        a0 = ka.d0k(x1, x2, p).reshape(n1, nf, n2)
        assert np.all(a0[:, [1, 3], :] == 0.0) and np.max(np.abs(a0[:, dims, :].reshape(n1 * 2, n2)
                                                                  - cls().d0k(x1[:, dims], x2[:, dims], p))) < 1e-12
        a1 = ka.d1k(x1, x2, p).reshape(n1, n2, nf)
        assert np.all(a1[:, :, [1, 3]] == 0.0)
        a01 = ka.d01k(x1, x2, p).reshape(n1, nf, n2, nf)
        assert np.all(a01[:, [1, 3]] == 0.0) and np.all(a01[:, :, :, [1, 3]] == 0.0)
        assert np.max(np.abs(a01[:, dims][:, :, :, dims].reshape(n1 * 2, n2 * 2)
                             - cls().d01k(x1[:, dims], x2[:, dims], p))) < 1e-12
        with pytest.raises(NotImplementedError):
            K.Matern32().d01kj(x1, x2, p, J1, J2)
        with pytest.raises(NotImplementedError):
            (K.SquaredExponential() + K.Matern52()).d01kj(x1, x2, {"kernel1": p, "kernel2": p}, J1, J2)


def test_state_checkpointer_callback(tmp_path):
    """StateCheckpointer (gpx/optimizers/scipy_optimize.py:175-229) through GPR.fit: one loadable .npz per L-BFGS-B
    iteration, the last one holding the fitted hyper-parameters."""
    import numpy as np

    from gpx_b200.kernels import SquaredExponential
    from gpx_b200.models import GPR
    from gpx_b200.optimizers import StateCheckpointer

    rng = np.random.default_rng(0)
    x = rng.uniform(-2, 2, (80, 1))
    y = np.sin(2 * x) + 0.1 * rng.standard_normal((80, 1))
    with patched():
        m = GPR(SquaredExponential())
        chk = StateCheckpointer(m.state, str(tmp_path / "run.npz"))
        m.fit(x, y, opt_kwargs=dict(callback=chk))
    assert chk.counter == m.optimize_results_.nit and chk.counter >= 2
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files[0] == "run.000.npz" and len(files) == chk.counter
    last = GPR(SquaredExponential()).load(str(tmp_path / files[-1]))
    for path in (("kernel_params", "lengthscale"), ("sigma",)):
        a = dict(last.state.flat_params())[path].value
        b = dict(m.state.flat_params())[path].value
        assert abs(a - b) < 1e-12
    first = GPR(SquaredExponential()).load(str(tmp_path / files[0]))
    assert first.state.params["sigma"].value != m.state.params["sigma"].value


@pytest.mark.parametrize("partitionable", [True, False])
def test_sampling_glue(partitionable):
    """GPR.sample(kind="prior" / "posterior") (gpx/models/base.py:99-128 -> gpr.py:761-832 -> models/utils.py:52-75)
    against the oracle's literal restatement, including the reference's quirks (one key split per column of the mean,
    only the last column's draw returned).  The normal deviates come from the product's own host PRNG."""
    import numpy as np

    from gpx_b200 import defaults
    from gpx_b200.kernels import Matern52
    from gpx_b200.models import GPR
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior
    from gpx_b200.random import PRNGKey
    from oracle import ref_numpy as R

    rng = np.random.default_rng(1)
    x = rng.uniform(-2, 2, (40, 2))
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((40, 1))
    xs = rng.uniform(-2, 2, (15, 2))
    ell, sigma = 1.3, 0.2
    old = defaults.threefry_partitionable()
    defaults.set_threefry_partitionable(partitionable)
    try:
        with patched():
            m = GPR(Matern52(), kernel_params=dict(lengthscale=Parameter(ell, True, Softplus(), GammaPrior())),
                    sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
            prior = m.sample(PRNGKey(0), xs, n_samples=4, kind="prior")
            with pytest.raises(RuntimeError):
                m.sample(PRNGKey(0), xs, kind="posterior")
            m.fit(x, y, minimize=False)
            post = m.sample(PRNGKey(7), xs, n_samples=3, kind="posterior")
            with pytest.raises(ValueError):
                m.sample(PRNGKey(0), xs, kind="both")
    finally:
        defaults.set_threefry_partitionable(old)
    ref_prior = R.sample_prior("m52", R.PRNGKey(0), xs, ell, sigma, 4, partitionable)
    assert prior.shape == (4, 15) and np.max(np.abs(prior - ref_prior)) < 1e-9
    c, mu = R.fit_dense("m52", x, y, ell, sigma)
    ref_post = R.sample_posterior("m52", R.PRNGKey(7), x, xs, c, mu, ell, sigma, 3, partitionable)
    assert post.shape == (3, 15) and np.max(np.abs(post - ref_post)) < 1e-7


def test_map_estimate_notebook_anchor_through_the_facade():
    """examples/map_estimate.ipynb: GPR(loss_fn=neg_log_posterior).fit lands on the printed MAP hyper-parameters
    (lengthscale 1.99896, sigma 0.4502): the prior term and its gradient are added on the host to the device LML."""
    import json
    import os

    import numpy as np

    from gpx_b200.bijectors import Softplus
    from gpx_b200.kernels import SquaredExponential
    from gpx_b200.models import GPR
    from gpx_b200.models.gpr import log_prior, neg_log_posterior
    from gpx_b200.parameters import Parameter
    from gpx_b200.priors import NormalPrior
    from test_oracle_cpu import notebook_gpr_data

    a = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_anchors.json")))["map_estimate"]["fit"]
    x, y = notebook_gpr_data()
    with patched():
        m = GPR(SquaredExponential(),
                kernel_params=dict(lengthscale=Parameter(2.0, True, Softplus(), NormalPrior(2.0, 0.15))),
                sigma=Parameter(1.0, True, Softplus(), NormalPrior()), loss_fn=neg_log_posterior)
        l0, g0 = m.loss_and_grad(x, y)
        assert abs(l0 - (-m.log_marginal_likelihood(x, y) - log_prior(m.state))) < 1e-12
        h = 1e-6  # the gradient of the MAP loss w.r.t. the constrained lengthscale, by central differences
        st = m.state
        fp = neg_log_posterior(st.with_param_value(("kernel_params", "lengthscale"), np.float64(2.0 + h)), x, y)
        fm = neg_log_posterior(st.with_param_value(("kernel_params", "lengthscale"), np.float64(2.0 - h)), x, y)
        assert abs(g0[("kernel_params", "lengthscale")] - (fp - fm) / (2 * h)) < 1e-6
        m.fit(x, y)
    assert abs(float(m.state.params["kernel_params"]["lengthscale"].value) - a["lengthscale"]) < 2e-5
    assert abs(float(m.state.params["sigma"].value) - a["sigma"]) < 1e-4


def test_vector_lengthscale_in_kernel_k_only():
    """A per-feature lengthscale broadcasts in Kernel.k (as x / lengthscale does in the reference) and is refused by
    the models, which assume the reference's default scalar."""
    import numpy as np

    import gpx_b200.kernels as K
    from gpx_b200.bijectors import Softplus
    from gpx_b200.models import GPR
    from gpx_b200.parameters import Parameter
    from gpx_b200.priors import GammaPrior
    from oracle import ref_numpy as R

    rng = np.random.default_rng(0)
    x1, x2 = rng.standard_normal((9, 3)), rng.standard_normal((4, 3))
    ell = np.array([0.7, 1.9, 3.1])
    p = {"lengthscale": Parameter(ell, True, Softplus(), GammaPrior(shape=(3,)))}
    with patched():
        for kind, cls in (("se", K.SquaredExponential), ("m52", K.Matern52)):
            got = cls().k(x1, x2, p)
            assert np.max(np.abs(got - R.kernel(kind, x1 / ell, x2 / ell, 1.0))) < 1e-14
            sym = cls().k(x1, x1, p)
            assert np.max(np.abs(sym - R.kernel(kind, x1 / ell, x1 / ell, 1.0))) < 1e-12
        with pytest.raises(ValueError):
            K.SquaredExponential(active_dims=[0, 1]).k(x1, x2, p)
        with pytest.raises(NotImplementedError):
            GPR(K.SquaredExponential(), kernel_params=p).log_marginal_likelihood(x1, rng.standard_normal((9, 1)))
