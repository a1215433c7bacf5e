"""GPU parity against the COMMITTED golden fixture tests/golden/oracle_small.npz (frozen oracle outputs on a small
seeded problem; tools/make_golden.py wrote it, tests/test_oracle_cpu.py checks that the oracle still reproduces it).
Everything goes through the GPX-style facade, i.e. through the C ABI; nothing of the oracle is evaluated here."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_small.npz")


def _par(v, trainable=True):
    from gpx_b200.bijectors import Softplus
    from gpx_b200.parameters import Parameter
    from gpx_b200.priors import GammaPrior

    return Parameter(float(v), trainable, Softplus(), GammaPrior())


def _close(a, b, rtol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= rtol * max(np.max(np.abs(b)), 1e-300), float(np.max(np.abs(a - b)))


def test_exact_gpr_against_the_fixture():
    import gpx_b200.kernels as K
    from gpx_b200.models import GPR

    g = np.load(GOLDEN)
    x, y, ell, sigma = g["x"], g["y"], float(g["ell"]), float(g["sigma"])
    kernels = {"se": K.SquaredExponential, "m12": K.Matern12, "m32": K.Matern32, "m52": K.Matern52}
    for kind, cls in kernels.items():
        m = GPR(cls(), kernel_params=dict(lengthscale=_par(ell)), sigma=_par(sigma))
        loss, grads = m.loss_and_grad(x, y)
        got = np.array([-loss, -grads[("kernel_params", "lengthscale")], -grads[("sigma",)]])
        _close(got, g[f"lml_grad_{kind}"], 1e-8)
    m = GPR(K.SquaredExponential(), kernel_params=dict(lengthscale=_par(ell)), sigma=_par(sigma)).fit(x, y, minimize=False)
    _close(m.c_, g["c_se"], 1e-8)
    assert abs(float(m.mu_) - float(g["mu"])) < 1e-14
    mean, cov = m.predict(g["xs"], full_covariance=True)
    _close(mean, g["pred_mean_se"], 1e-8)
    assert np.max(np.abs(cov - g["pred_cov_se"])) < 1e-8  # relative to k(x, x) = 1


@pytest.mark.parametrize("partitionable,name", [(True, "pivots_part"), (False, "pivots_legacy")])
def test_rpcholesky_pivots_against_the_fixture_bit_exact(partitionable, name):
    import gpx_b200.kernels as K
    from gpx_b200 import defaults
    from gpx_b200.models import SGPR_RPChol
    from gpx_b200.random import PRNGKey

    g = np.load(GOLDEN)
    old = defaults.threefry_partitionable()
    defaults.set_threefry_partitionable(partitionable)
    try:
        m = SGPR_RPChol(K.SquaredExponential(), n_locs=16, key=PRNGKey(2023),
                        kernel_params=dict(lengthscale=_par(float(g["ell"]))), sigma=_par(float(g["sigma"])))
        m.fit(g["x"], g["y"], minimize=False)
        assert np.array_equal(np.asarray(m.state.pivots), g[name])
    finally:
        defaults.set_threefry_partitionable(old)


def test_sgpr_composite_and_td_against_the_fixture():
    import gpx_b200.kernels as K
    from gpx_b200.models import GPR, GPR_TD, SGPR
    from gpx_b200.models.sgpr import locs_parameter

    g = np.load(GOLDEN)
    x, y, ell, sigma, xl, xs = g["x"], g["y"], float(g["ell"]), float(g["sigma"]), g["x_locs"], g["xs"]
    s = SGPR(K.SquaredExponential(), x_locs=locs_parameter(xl), kernel_params=dict(lengthscale=_par(ell)), sigma=_par(sigma))
    lml = s.log_marginal_likelihood(x, y)
    assert abs(lml - float(g["sgpr_lml_se"])) < 1e-8 * abs(lml) and abs(lml - float(g["sgpr_lml_nxn_se"])) < 1e-8 * abs(lml)
    s.fit(x, y, minimize=False)
    _close(s.c_, g["sgpr_c_se"], 1e-6)  # normal equations: cond^2
    mean, cov = s.predict(xs, full_covariance=True)
    _close(mean, g["sgpr_pred_mean_se"], 1e-7)
    assert np.max(np.abs(cov - g["sgpr_pred_cov_se"])) < 1e-7

    kern = K.SquaredExponential(active_dims=[0, 1]) + K.Matern52() * K.Linear(active_dims=[2, 3])
    kp = {"kernel1": dict(lengthscale=_par(1.4)), "kernel2": {"kernel1": dict(lengthscale=_par(2.2)), "kernel2": {}}}
    c = GPR(kern, kernel_params=kp, sigma=_par(sigma)).log_marginal_likelihood(x, y)
    assert abs(c - float(g["composite_lml"])) < 1e-8 * abs(c)
    cs = SGPR(kern, x_locs=locs_parameter(xl), kernel_params=kp, sigma=_par(sigma)).log_marginal_likelihood(x, y)
    assert abs(cs - float(g["composite_sgpr_lml"])) < 1e-8 * abs(cs)

    J, yd = g["td_J"], g["td_yd"]
    Nt = J.shape[0]
    t1 = GPR_TD(K.SquaredExponential(), mean_function=_data_mean(), kernel_params=dict(lengthscale=_par(ell)),
                sigma_targets=_par(0.3), sigma_derivs=_par(0.2))
    v = t1.log_marginal_likelihood(x[:Nt], y[:Nt], J, yd)
    assert abs(v - float(g["td_lml_se"])) < 1e-8 * abs(v)
    t2 = GPR_TD(K.Matern52(), mean_function=_data_mean(), kernel_params=dict(lengthscale=_par(ell)),
                sigma_targets=_par(0.3), sigma_derivs=_par(0.2), sigma_derivs2=_par(0.4))
    v2 = t2.log_marginal_likelihood(x[:Nt], y[:Nt], J[:, :, :3].copy(), yd[:, :3].copy(), jacobian_2=J[:, :, 3:].copy(),
                                    y_derivs_2=yd[:, 3:].copy())
    assert abs(v2 - float(g["td2_lml_m52"])) < 1e-8 * abs(v2)
    t2.fit(x[:Nt], y[:Nt], J[:, :, :3].copy(), yd[:, :3].copy(), jacobian_2=J[:, :, 3:].copy(), y_derivs_2=yd[:, 3:].copy(),
           minimize=False)
    _close(t2.state.c, g["td2_c_m52"], 1e-8)


def _data_mean():
    from gpx_b200.mean_functions import data_mean

    return data_mean
