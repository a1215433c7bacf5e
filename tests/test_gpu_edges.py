"""GPU edge cases through the C ABI: tiny and ragged sizes, single rows/columns, duplicates,
wide feature counts, argument errors (no CPU fallback: errors surface as GpxbError)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as R  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


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


@pytest.mark.parametrize("kind", ["se", "m52"])
@pytest.mark.parametrize("D", [105, 256, 1000])
def test_gram_streams_large_feature_counts(kind, D):
    """Beyond two whole operand tiles of shared memory (D > 104) the Gram kernel streams the features in 32-column
    slabs (K-tiled DMMA): kernelizers.ipynb's D = 100 timing case and SOAP-like descriptors (D ~ 1000)."""
    from gpx_b200 import ops

    rng = np.random.default_rng(D)
    x1, x2 = rng.standard_normal((300, D)), rng.standard_normal((131, D))
    ell = 1.1 * np.sqrt(D)
    K = host(ops.gram(kind, dev(x1), dev(x2), ell))
    assert np.max(np.abs(K - R.kernel(kind, x1, x2, ell))) < 1e-13
    x1d = dev(x1)
    Ks = host(ops.gram(kind, x1d, x1d, ell, diag_add=0.5, lower_only=True))
    ref = np.tril(R.kernel(kind, x1, x1, ell) + 0.5 * np.eye(300))
    assert np.max(np.abs(np.tril(Ks) - ref)) < 1e-13


@pytest.mark.parametrize("D", [100, 256, 1000])
def test_dense_gpr_and_matvec_with_large_feature_counts(D):
    """LML + gradient, fit / predict and the matrix-free operator (row-chunked Gram + GEMM beyond the fused kernel's
    tile) against the oracle at feature counts the round-1 kernels rejected."""
    from gpx_b200 import ops

    rng = np.random.default_rng(D + 1)
    N = 700
    x = rng.standard_normal((N, D))
    y = np.sin(x[:, :3].sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    ell, sigma = 1.3 * np.sqrt(D), 0.3
    out = host(ops.gpr_lml_grad("m52", dev(x), dev(y), ell, sigma))
    ref = np.array(R.lml_grad_dense("m52", x, y, ell, sigma))
    assert np.max(np.abs(out - ref)) < 1e-8 * np.max(np.abs(ref))
    c, mu = ops.gpr_fit("m52", dev(x), dev(y), ell, sigma)
    c_ref, mu_ref = R.fit_dense("m52", x, y, ell, sigma)
    xs = rng.standard_normal((90, D))
    mean = host(ops.gpr_predict("m52", dev(x), dev(xs), c, mu, ell, sigma))
    assert rel(mean, R.predict_dense("m52", x, xs, c_ref, mu_ref, ell, sigma)) < 1e-8
    V = rng.standard_normal((N, 5))
    xd = dev(x)
    W = host(ops.kmatvec("se", xd, xd, dev(V), ell, diag_add=0.25))
    Wref = (R.kernel("se", x, x, ell) + 0.25 * np.eye(N)) @ V
    assert rel(W, Wref) < 1e-12
    W1 = host(ops.kmatvec("se", dev(x[100:333]), xd, dev(V[:, :1]), ell, diag_add=0.25, row_offset=100))
    assert rel(W1, Wref[100:333, :1]) < 1e-12
    # iterative fit through the same operator
    key = R.PRNGKey(3)
    ci, mui, _ = ops.gpr_fit_iter("se", dev(x), dev(y), ell, sigma, 20, key, tol=1e-13, atol=0.0)
    cr, _ = R.fit_dense("se", x, y, ell, sigma)
    assert rel(host(ci), cr) < 1e-8


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
