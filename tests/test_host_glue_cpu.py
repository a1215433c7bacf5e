"""Host-logic tests (no GPU): the facade modules that are pure compositions of device calls are run with the CPU
stand-ins of tests/_fake_ops.py, through the very assertions the GPU parity tests make.  This checks the glue --
adjoint algebra, lower-triangle conventions, parameter paths -- not the kernels (the `-m gpu` tests do that)."""
import pytest

torch = pytest.importorskip("torch")

import test_gpu_composite as G  # noqa: E402  (its tests are plain functions; the gpu mark stays on that module)
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
