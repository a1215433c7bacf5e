"""GPU edge cases through the C ABI: tiny and ragged sizes, single rows/columns, duplicates,
wide feature counts, argument errors (no CPU fallback: errors surface as GpxbError)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as R  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("N", [1, 2, 3, 127, 129, 1025])
def test_lml_grad_tiny_and_ragged(N):
    from gpx_b200 import ops

    rng = np.random.default_rng(N)
    x = rng.standard_normal((N, 3))
    y = rng.standard_normal((N, 1))
    ref = np.array(R.lml_grad_dense("se", x, y, 1.3, 0.4))
    out = host(ops.gpr_lml_grad("se", dev(x), dev(y), 1.3, 0.4))
    assert np.max(np.abs(out - ref)) <= 1e-8 * max(1.0, np.max(np.abs(ref)))


def test_duplicate_points_and_single_feature():
    from gpx_b200 import ops

    x = np.array([[0.0], [0.0], [1.0], [1.0], [2.5]])
    K = host(ops.gram("m12", dev(x), dev(x.copy()), 0.7))
    assert np.max(np.abs(K - R.kernel("m12", x, x, 0.7))) < 1e-9  # sqrt(1e-12 + eta) on coincident points
    xd = dev(x)
    Ks = host(ops.gram("m52", xd, xd, 0.7))
    assert np.max(np.abs(Ks - R.kernel("m52", x, x, 0.7))) < 1e-13


@pytest.mark.parametrize("D", [1, 4, 5, 33, 100, 104])
def test_gram_feature_counts(D):
    from gpx_b200 import ops

    rng = np.random.default_rng(D)
    x1, x2 = rng.standard_normal((70, D)), rng.standard_normal((45, D))
    ell = 1.1 * np.sqrt(D)
    K = host(ops.gram("se", dev(x1), dev(x2), ell))
    assert np.max(np.abs(K - R.kernel("se", x1, x2, ell))) < 1e-13


def test_too_many_features_is_an_error_not_a_fallback():
    from gpx_b200 import _lib, ops

    x = dev(np.zeros((10, 200)))
    with pytest.raises(_lib.GpxbError):
        ops.gram("se", x, x, 1.0)
    assert "argument check" in _lib.last_error()


def test_rpcholesky_more_pivots_than_rank_and_single_pivot():
    from gpx_b200 import ops

    rng = np.random.default_rng(0)
    x = rng.standard_normal((40, 2))
    key = R.PRNGKey(5)
    for M in (1, 40):
        Fref, pref = R.rpcholesky("se", x, M, 1.5, key)
        F, piv = ops.rpcholesky("se", dev(x), M, 1.5, key)
        assert np.array_equal(host(piv), pref)
        assert np.max(np.abs(host(F) - Fref)) < 1e-6  # late columns divide by sqrt(~0)+1e-6


def test_kmatvec_tiny():
    from gpx_b200 import ops

    rng = np.random.default_rng(1)
    for N in (1, 7, 130):
        x = rng.standard_normal((N, 2))
        V = rng.standard_normal((N, 3))
        xd = dev(x)
        W = host(ops.kmatvec("m32", xd, xd, dev(V), 1.2, diag_add=0.3))
        ref = (R.kernel("m32", x, x, 1.2) + 0.3 * np.eye(N)) @ V
        assert np.max(np.abs(W - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_predict_single_test_point_and_multi_output():
    from gpx_b200 import ops

    rng = np.random.default_rng(2)
    x = rng.standard_normal((200, 3))
    y = rng.standard_normal((200, 2))
    c, mu = ops.gpr_fit("se", dev(x), dev(y), 1.4, 0.3)
    c_ref, mu_ref = R.fit_dense("se", x, y, 1.4, 0.3)
    assert np.max(np.abs(host(c) - c_ref)) < 1e-8 * np.max(np.abs(c_ref))
    xs = rng.standard_normal((1, 3))
    mean, cov = ops.gpr_predict("se", dev(x), dev(xs), c, mu, 1.4, 0.3, full_covariance=True)
    mr, cr = R.predict_dense("se", x, xs, c_ref, mu_ref, 1.4, 0.3, full_covariance=True)
    assert np.max(np.abs(host(mean) - mr)) < 1e-8 and np.max(np.abs(host(cov) - cr)) < 1e-8
