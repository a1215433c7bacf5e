"""Host-side facade logic that needs no GPU: construction, parameter table, bijectors,
checkpoint round trip, error behaviour (mirrors tests/gpx/models/test_models.py:139-155,
tests/gpx/parameters/*, tests/gpx/test_utils.py:56-100 of the reference)."""
import numpy as np
import pytest

from gpx_b200.bijectors import Identity, Softplus
from gpx_b200.kernels import Matern12, Matern32, Matern52, SquaredExponential
from gpx_b200.models import GPR
from gpx_b200.parameters import Parameter
from gpx_b200.priors import GammaPrior, NormalPrior


@pytest.mark.parametrize("K", [SquaredExponential, Matern12, Matern32, Matern52])
def test_gpr_construction(K):
    m = GPR(kernel=K())
    assert isinstance(m, GPR)
    assert m.state.params["kernel_params"]["lengthscale"].value == 1.0
    assert m.state.params["sigma"].value == 1.0 and m.state.params["sigma"].trainable
    assert m.state.is_fitted is False and m.c_ is None and m.mu_ is None


def test_predict_unfitted_raises():
    m = GPR(kernel=SquaredExponential())
    with pytest.raises(RuntimeError):
        m.predict(np.zeros((3, 2)))


def test_softplus_roundtrip_and_grad():
    b = Softplus()
    x = np.array([-5.0, -1.0, 0.0, 0.3, 4.0, 30.0])
    np.testing.assert_allclose(b.backward(b.forward(x)), x, rtol=1e-10, atol=1e-12)
    h = 1e-6
    np.testing.assert_allclose(b.grad_forward(x), (b.forward(x + h) - b.forward(x - h)) / (2 * h), rtol=1e-6)
    assert np.array_equal(Identity().forward(x), x)


def test_parameter_shape_check_and_update():
    with pytest.raises(ValueError):
        Parameter(np.ones(3), True, Softplus(), NormalPrior())
    p = Parameter(2.0, True, Softplus(), GammaPrior())
    q = p.update(dict(value=3.0, trainable=False))
    assert p.value == 2.0 and q.value == 3.0 and not q.trainable


def test_state_save_load_roundtrip(tmp_path):
    m = GPR(kernel=Matern52(), sigma=Parameter(0.3, False, Softplus(), GammaPrior()))
    f = str(tmp_path / "state.npz")
    d = m.save(f)
    assert "params:kernel_params:lengthscale" in d and "params:sigma" in d
    m2 = GPR(kernel=Matern52()).load(f)
    assert float(m2.state.params["sigma"].value) == 0.3


def test_trainable_flattening_and_chain_rule():
    from gpx_b200.optimizers import ravel_backward_trainables, unravel_forward

    m = GPR(kernel=SquaredExponential(), sigma=Parameter(0.5, False, Softplus(), GammaPrior()))
    x0, paths = ravel_backward_trainables(m.state)
    assert paths == [(("kernel_params", "lengthscale"), ())]
    st = unravel_forward(m.state, paths, x0 + 0.1)
    assert st.params["kernel_params"]["lengthscale"].value > 1.0
    assert st.params["sigma"].value == 0.5


def test_load_reference_style_checkpoint(tmp_path):
    """A checkpoint laid out as the reference's ModelState.save writes it (gpx/parameters/model_state.py:246-300):
    "params:kernel_params:lengthscale"-style keys plus every registered auxiliary entry, unset ones as pickled
    None placeholders."""
    f = str(tmp_path / "ref_state.npz")
    np.savez(f, **{"params:kernel_params:lengthscale": np.float64(0.7), "params:sigma": np.float64(0.2),
                   "x_train": np.arange(6.0).reshape(3, 2), "y_train": np.ones((3, 1)), "c": np.full((3, 1), 2.0),
                   "mu": np.float64(1.5), "is_fitted": np.bool_(True), "is_fitted_derivs": np.bool_(False),
                   "jacobian_train": np.array(None, dtype=object), "jaccoef": np.array(None, dtype=object)})
    m = GPR(kernel=Matern52()).load(f)
    assert float(m.state.params["kernel_params"]["lengthscale"].value) == 0.7
    assert float(m.state.params["sigma"].value) == 0.2
    assert m.state.is_fitted is True and m.state.c.shape == (3, 1) and float(m.state.mu) == 1.5
    m2 = GPR(kernel=Matern52())
    m2.state = m2.state.load(f, allow_pickle=True)
    assert m2.state.is_fitted is True


def test_host_prng_matches_the_oracle_bit_for_bit():
    """gpx_b200.random (the product's host-side PRNGKey / split, written independently of the oracle) against the
    oracle's restatement of jax.random, which is pinned by the Random123 KATs and the notebook anchors: every word
    identical, for both jax_threefry_partitionable modes, over seeded random keys and the 2023 key the reference uses."""
    import json
    import os

    from gpx_b200 import random as PR
    from oracle import ref_numpy as R

    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "threefry2x32_kat.json")))
    for case in kat["vectors"]:
        k, c, o = case["key"], case["ctr"], case["out"]
        assert list(PR._threefry2x32(int(k[0], 16), int(k[1], 16), int(c[0], 16), int(c[1], 16))) == [int(o[0], 16), int(o[1], 16)]
    assert np.array_equal(PR.PRNGKey(2023), R.PRNGKey(2023)) and np.array_equal(PR.PRNGKey(2**40 + 7), R.PRNGKey(2**40 + 7))
    rng = np.random.default_rng(0)
    keys = [PR.PRNGKey(2023), PR.PRNGKey(0)] + [rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32) for _ in range(20)]
    for key in keys:
        for part in (True, False):
            for num in (2, 3, 5):
                assert np.array_equal(PR.split(key, num, part), R.split(key, num, part)), (key, part, num)
    with pytest.raises(ValueError):
        PR.as_key(np.zeros(3, dtype=np.uint32))
    for key in keys[:8]:
        for part in (True, False):
            for shape in ((), (1,), (7,), (4, 5), (100,)):
                a, b = np.asarray(PR.uniform(key, shape, part)), np.asarray(R.uniform(key, shape, part))
                assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
                if shape:
                    assert np.array_equal(PR.normal(key, shape, part).view(np.uint64),
                                          R.normal(key, shape, part).view(np.uint64))


def test_slq_host_numerics_value_and_closed_form_derivative():
    """Host part of the SLQ log-determinant (gpx_b200/models/_iterative.py; gpx/lanczos.py:171-183): the value against
    the oracle on tridiagonals produced by the oracle's Lanczos, and the closed-form eigh JVP against central
    differences of the value along a tridiagonal perturbation."""
    from gpx_b200.models._iterative import slq_logdet, slq_logdet_grad
    from oracle import ref_numpy as R

    rng = np.random.default_rng(5)
    n, P, m = 60, 3, 7
    x = rng.standard_normal((n, 2))
    A = R.kernel("se", x, x, 1.3) + 0.3 * np.eye(n)
    us, key = R.rademacher_probes(R.PRNGKey(2023), n, P)
    alpha, beta = np.zeros((P, m)), np.zeros((P, m))
    ref = 0.0
    for p in range(P):
        T, _ = R.lanczos_tridiagonal(lambda v: A @ v, m, us[:, p], key)
        alpha[p], beta[p, : m - 1] = np.diag(T), np.diag(T, 1)
        ref += R.slq_from_tridiag(T)
    assert abs(slq_logdet(alpha, beta, n) - (n / P) * ref) < 1e-12 * abs(ref)
    dalpha, dbeta = rng.standard_normal((2, P, m)), rng.standard_normal((2, P, m))
    val, grad = slq_logdet_grad(alpha, beta, dalpha, dbeta, n)
    assert abs(val - (n / P) * ref) < 1e-12 * abs(ref)
    h = 1e-6
    for q in range(2):
        fd = (slq_logdet(alpha + h * dalpha[q], beta + h * dbeta[q], n)
              - slq_logdet(alpha - h * dalpha[q], beta - h * dbeta[q], n)) / (2 * h)
        assert abs(grad[q] - fd) < 1e-7 * max(1.0, abs(fd)), (q, grad[q], fd)


def test_array_valued_trainables_ravel_like_ravel_pytree():
    """Flat vector of unconstrained trainables (gpx/optimizers/utils.py:73-131): scalars and a trainable (M, D) landmark
    array are concatenated in parameter-tree order and restored shape by shape; the chain rule uses each bijector."""
    from gpx_b200.models import SGPR
    from gpx_b200.models.sgpr import locs_parameter
    from gpx_b200.optimizers import _split, ravel_backward_trainables, unravel_forward

    xl = np.arange(12.0).reshape(6, 2) / 7.0
    m = SGPR(kernel=SquaredExponential(), x_locs=locs_parameter(xl, trainable=True))
    x0, paths = ravel_backward_trainables(m.state)
    assert [p for p, _ in paths] == [("kernel_params", "lengthscale"), ("sigma",), ("x_locs",)]
    assert x0.shape == (2 + 12,) and np.array_equal(x0[2:], xl.reshape(-1))  # Identity bijector on the landmarks
    np.testing.assert_allclose(x0[:2], Softplus().backward(np.array([1.0, 1.0])))
    xt = x0 + np.linspace(0.01, 0.14, 14)
    st = unravel_forward(m.state, paths, xt)
    assert st.params["x_locs"].value.shape == (6, 2)
    np.testing.assert_allclose(st.params["x_locs"].value, xt[2:].reshape(6, 2))
    np.testing.assert_allclose(st.params["sigma"].value, Softplus().forward(xt[1]))
    assert [v.shape for _, v in _split(paths, xt)] == [(), (), (6, 2)]
    assert m.state.params["x_locs"].value[0, 0] == 0.0  # the original state is untouched


def test_ops_wrappers_reject_mismatched_shapes_before_reaching_the_c_abi(monkeypatch):
    """The C ABI sees only pointers: N, D, T are taken from one tensor, so every wrapper must check the OTHER
    operands itself (a wrong feature count or a short y would make a kernel read outside its buffers).  The
    reference raises a shape error from jnp; here it is a ValueError raised before any library call."""
    import torch

    from gpx_b200 import ops

    class _Lib:
        def __getattr__(self, name):
            if name == "gpxb_potrf_inv_size":
                return lambda n: 0
            raise AssertionError(f"{name} reached with mismatched shapes")

    class _Ctx:
        lib, handle = _Lib(), None

    monkeypatch.setattr(ops, "context", lambda: _Ctx())
    monkeypatch.setattr(ops, "_f64", lambda t: t.contiguous())
    z = lambda *s: torch.zeros(s, dtype=torch.float64)  # noqa: E731
    N, D, nv, M = 6, 3, 2, 4
    x, y, J = z(N, D), z(N, 1), z(N, D, nv)
    bad = [
        lambda: ops.gpr_lml_grad("se", x, z(N - 1, 1), 1.0, 0.1),
        lambda: ops.gpr_fit("se", x, z(N + 2, 1), 1.0, 0.1),
        lambda: ops.gpr_predict("se", x, z(5, D + 1), z(N, 1), z(1), 1.0, 0.1),
        lambda: ops.gpr_predict("se", x, z(5, D), z(N - 1, 1), z(1), 1.0, 0.1),
        lambda: ops.hess_gram("se", x, z(N, D + 1, nv), None, None, 1.0),
        lambda: ops.hess_gram("se", x, J, z(4, D + 1), z(4, D + 1, nv), 1.0),
        lambda: ops.gpr_lml_derivs("se", x, J, z(N, nv + 1), 1.0, 0.1),
        lambda: ops.gpr_lml_grad_derivs("se", x, z(N - 1, D, nv), z(N, nv), 1.0, 0.1),
        lambda: ops.gpr_fit_derivs("se", x, J, z(N - 1, nv), 1.0, 0.1),
        lambda: ops.gpr_predict_y_derivs("se", x, z(N, D + 1), z(5, D), 1.0),
        lambda: ops.gpr_predict_derivs_full("se", x, J, z(5, D), z(5, D, nv), z(N * nv + 1), 1.0, 0.1),
        lambda: ops.gpr_predict_y_derivs_full("se", x, J, z(5, D + 1), z(N * nv), 1.0, 0.1),
        lambda: ops.hess_matvec("se", x, J, x, J, z(N * nv + 1, 2), 1.0),
        lambda: ops.td_gram("se", x, J, z(4, D), z(4, D + 1, nv), 1.0),
        lambda: ops.gpr_td_lml_grad("se", x, J, z(N - 1), z(N, nv), 1.0, 0.1, 0.1),
        lambda: ops.kmatvec("se", z(5, D + 1), x, z(N, 2), 1.0),
        lambda: ops.lanczos("se", x, 1.0, 0.1, z(N + 1, 2), 3),
        lambda: ops.kmatvec_dl("se", x, z(N - 1, 2), 1.0),
        lambda: ops.gpr_fit_iter("se", x, z(N - 1, 1), 1.0, 0.1, 2, (0, 1)),
        lambda: ops.sgpr_lml("se", x, z(N - 1, 1), z(M, D), 1.0, 0.1),
        lambda: ops.sgpr_lml_grad("se", x, y, z(M, D + 1), 1.0, 0.1),
        lambda: ops.sgpr_fit("se", x, z(N + 1, 1), z(M, D), 1.0, 0.1),
        lambda: ops.sgpr_predict("se", z(M, D), z(5, D + 1), z(M, 1), z(1), 1.0, 0.1),
        lambda: ops.sgpr_predict("se", z(M, D), z(5, D), z(M + 1, 1), z(1), 1.0, 0.1),
        lambda: ops.gram_dl_contract("se", x, z(4, D), 1.0, z(N, 5)),
    ]
    for i, call in enumerate(bad):
        with pytest.raises(ValueError):
            call()
            raise AssertionError(f"case {i} was accepted")


def test_facade_rejects_x_y_row_mismatch_before_any_device_call():
    from gpx_b200.models import gpr as gpr_mod

    class _K:
        active_dims = None

        @staticmethod
        def _slice(x):
            return x

    class _S:
        kernel = _K()

    import _fake_ops as F

    saved = gpr_mod._to_device
    gpr_mod._to_device = F.to_device
    try:
        with pytest.raises(ValueError):
            gpr_mod._xy(_S(), np.zeros((5, 2)), np.zeros((4, 1)))
        with pytest.raises(ValueError):
            gpr_mod._xyj(_S(), np.zeros((5, 2)), np.zeros((5, 3)), np.zeros((5, 3, 3)))
        with pytest.raises(ValueError):
            gpr_mod._xyj(_S(), np.zeros((5, 2)), np.zeros((4, 3)), np.zeros((5, 2, 3)))
    finally:
        gpr_mod._to_device = saved


def test_load_rejects_a_checkpoint_with_other_parameters(tmp_path):
    """gpx/parameters/model_state.py:321-327: 'Wrong number of parameters in dumped file'."""
    m = GPR(SquaredExponential())
    f = str(tmp_path / "s.npz")
    d = m.state.save(f)
    np.savez(f, **{k: v for k, v in d.items() if k != "params:sigma"})
    with pytest.raises(ValueError, match="Wrong number of parameters"):
        m.state.load(f)
    np.savez(f, **dict(d, **{"params:kernel_params:period": np.array(1.0)}))
    with pytest.raises(ValueError, match="unknown"):
        m.state.load(f)
    np.savez(f, **dict(d, **{"params:sigma": np.ones(3)}))
    with pytest.raises(ValueError, match="shape"):
        m.state.load(f)
