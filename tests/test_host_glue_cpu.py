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
