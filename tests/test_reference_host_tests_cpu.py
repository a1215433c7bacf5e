"""The reference's own host-side unit tests, restated for this package's classes (same cases, same expectations):
tests/gpx/parameters/test_parameter.py:10-100, tests/gpx/parameters/test_model_state.py:14-140 and
tests/gpx/models/test_models.py:139-155 of GPX.  (The JAX pytree round trip of test_parameter.py:43-49 has no
counterpart here: Parameter is a plain Python object.)"""
import numpy as np
import pytest
from numpy.testing import assert_equal

from gpx_b200.bijectors import Identity, Softplus
from gpx_b200.kernels import SquaredExponential
from gpx_b200.mean_functions import data_mean
from gpx_b200.models import GPR, SGPR
from gpx_b200.parameters import ModelState, Parameter
from gpx_b200.priors import NormalPrior
from gpx_b200.random import PRNGKey


# ---- Parameter (tests/gpx/parameters/test_parameter.py) ---------------------------------------------------
def test_value_and_prior_consistency():
    assert isinstance(Parameter(1.0, True, Softplus(), NormalPrior(shape=(), dtype=np.float64)), Parameter)
    with pytest.raises(ValueError):  # same shape, different dtype (a Python int is int64)
        Parameter(1, True, Softplus(), NormalPrior(shape=(), dtype=np.float64))
    with pytest.raises(ValueError):  # different shape, same dtype
        Parameter(1.0, True, Softplus(), NormalPrior(shape=(2, 1), dtype=np.float64))
    with pytest.raises(ValueError):  # different shape, different dtype
        Parameter(1.0, True, Softplus(), NormalPrior(shape=(2, 1), dtype=np.int64))


@pytest.mark.parametrize("update_dict", [dict(value=2.0), dict(value=np.array([2.0, 1.0]), prior=NormalPrior(shape=(2,)))])
def test_update(update_dict):
    assert isinstance(Parameter(1.0, True, Softplus(), NormalPrior()).update(update_dict), Parameter)


@pytest.mark.parametrize("update_dict", [dict(value=2.0, prior=NormalPrior(shape=(2,))),
                                         dict(value=1.0, prior=NormalPrior(dtype=np.int64))])
def test_update_fails(update_dict):
    with pytest.raises(ValueError):
        Parameter(1.0, True, Softplus(), NormalPrior()).update(update_dict)


def test_copy():
    p1 = Parameter(1.0, True, Softplus(), NormalPrior())
    p2 = p1.copy()
    assert p1 is not p2 and p1.value == p2.value
    p1.value = 5.0  # modifying one parameter does not affect the other
    assert p1.value != p2.value


def test_sample_prior():
    p1 = Parameter(1.0, True, Softplus(), NormalPrior())
    p2 = p1.sample_prior(np.random.default_rng(2023))
    assert isinstance(p2, Parameter) and p2.value != p1.value
    assert p2.value > 0.0  # constrained by the bijector


# ---- ModelState (tests/gpx/parameters/test_model_state.py) ------------------------------------------------
def create_model_state():
    return ModelState(kernel=SquaredExponential(), mean_function=data_mean,
                      params={"kernel_params": {"lengthscale": Parameter(1.0, True, Softplus(), NormalPrior())},
                              "sigma": Parameter(0.1, False, Softplus(), NormalPrior())})


@pytest.mark.parametrize("aux", [1, 2.0, np.array(3.0), np.array([4.0]), np.array("e"), "f", np.array([1.0, 2.0]),
                                 np.array([[3, 4]]), True, False, None])
def test_save_and_reload_state(tmp_path, aux):
    state_file = str(tmp_path / "tmp_state_file.npz")
    state = create_model_state().update({"aux": aux})
    state.save(state_file)
    new_state = state.load(state_file)
    assert state.kernel is new_state.kernel
    old = [state.params["sigma"], state.params["kernel_params"]["lengthscale"]]
    new = [new_state.params["sigma"], new_state.params["kernel_params"]["lengthscale"]]
    for param, new_param in zip(old, new):
        assert isinstance(param, Parameter) and isinstance(new_param, Parameter)
        assert_equal(param.value, new_param.value)
        assert param.value is not new_param.value
        assert param.bijector is new_param.bijector and param.trainable is new_param.trainable
    assert hasattr(state, "aux") and hasattr(new_state, "aux")
    if isinstance(state.aux, np.ndarray):
        assert np.all(state.aux == new_state.aux)
    else:
        assert state.aux == new_state.aux


def test_shallow_copy():
    state = create_model_state()
    copied = state.copy()
    assert state is not copied
    assert state.kernel is copied.kernel          # the kernel is not copied
    assert state.params is not copied.params      # the parameter dictionary is
    state.kernel = Softplus()                     # re-assignment touches one state only
    assert state.kernel is not copied.kernel


def test_randomize_model_state():
    state = create_model_state()
    randomized = state.randomize(PRNGKey(2023))
    assert isinstance(randomized, ModelState) and randomized is not state
    assert state.kernel is randomized.kernel
    pairs = [(state.params["kernel_params"]["lengthscale"], randomized.params["kernel_params"]["lengthscale"]),
             (state.params["sigma"], randomized.params["sigma"])]
    for param, new_param in pairs:
        assert isinstance(param, Parameter) and isinstance(new_param, Parameter)
        assert param.bijector is new_param.bijector and param.trainable is new_param.trainable
    assert pairs[0][0].value != pairs[0][1].value   # trainable: re-drawn from its prior
    assert pairs[1][0].value == pairs[1][1].value   # not trainable: kept


# ---- model construction (tests/gpx/models/test_models.py) -------------------------------------------------
def _x_locs():
    return np.linspace(-1.0, 1.0, 10).reshape(-1, 1)


def test_init_models():
    ell = Parameter(1.0, True, Softplus(), NormalPrior())
    sig = Parameter(0.1, False, Softplus(), NormalPrior())
    xl = Parameter(_x_locs(), False, Identity(), NormalPrior(shape=_x_locs().shape))
    assert isinstance(GPR(kernel=SquaredExponential()), GPR)
    assert isinstance(GPR(kernel=SquaredExponential(), kernel_params={"lengthscale": ell}, sigma=sig), GPR)
    assert isinstance(SGPR(kernel=SquaredExponential(), x_locs=xl), SGPR)
    assert isinstance(SGPR(kernel=SquaredExponential(), x_locs=xl, kernel_params={"lengthscale": ell}, sigma=sig), SGPR)
    with pytest.raises(TypeError):
        SGPR(kernel=SquaredExponential(), x_locs=_x_locs())  # x_locs must be a Parameter (gpx/models/sgpr.py:706-709)


def test_models_from_state():
    state_gpr = create_model_state()
    state_sgpr = ModelState(kernel=SquaredExponential(), mean_function=data_mean,
                            params={"kernel_params": {"lengthscale": Parameter(1.0, True, Softplus(), NormalPrior())},
                                    "sigma": Parameter(0.1, False, Softplus(), NormalPrior()),
                                    "x_locs": Parameter(_x_locs(), False, Identity(), NormalPrior(shape=_x_locs().shape))})
    assert isinstance(GPR.from_state(state_gpr), GPR)
    assert isinstance(SGPR.from_state(state_sgpr), SGPR)


# ---- bijectors (tests/gpx/test_utils.py:44-100) ------------------------------------------------------------
@pytest.mark.parametrize("value", [-1e4, -1e2, -50.0, 0.0, 1e3])
def test_softplus(value):
    assert Softplus().forward(np.array([value])) >= 0


@pytest.mark.parametrize("value", [-600.0, -1e2, 0.0, 1e2, 1e5, 1e10])
def test_inverse_softplus(value):
    x = np.array([value])
    bij = Softplus()
    assert_equal(x, bij.backward(bij.forward(x)))  # exact recovery, as the reference asserts


@pytest.mark.parametrize("value", [-1e10, -1e3, -1e-5, -1e-300])
def test_inverse_softplus_negative(value):
    with np.errstate(invalid="ignore", over="ignore"):
        assert np.isnan(Softplus().backward(np.array([value])))


@pytest.mark.parametrize("value", [-1e10, -1e1, 0.0, 1e1, 1e10])
def test_softplus_plus_inversion_stability(value):
    bij = Softplus()
    with np.errstate(over="ignore"):
        assert np.isfinite(bij.backward(bij.forward(np.array([value]))))


def test_identity():
    x = np.array([1.0])
    assert_equal(x, Identity().forward(x))
