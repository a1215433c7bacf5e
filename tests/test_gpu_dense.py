"""GPU parity: Family 1 (Gram) and Family 2 (GEMM / Cholesky / TRSM / inverse / exact GPR)
through the C ABI, against the NumPy oracle (oracle/ref_numpy.py) on the same seeded inputs.

Tolerances: kernel entries 1e-13 abs; LML, gradient, fit and predict 1e-8 relative (the
tolerance BASELINE.json's north_star states)."""
import numpy as np
import pytest
import scipy.linalg as sla

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as R  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (256, 384, 200), (100, 37, 53), (1, 300, 129), (513, 5, 1000)])
def test_gemm_variants(ta, tb, M, N, K):
    from gpx_b200 import ops

    rng = np.random.default_rng(0)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 1.3 * C0
    C = ops.gemm(dev(A), dev(B), transa=ta, transb=tb, alpha=0.7, beta=-1.3, C=dev(C0))
    torch.cuda.synchronize()
    assert rel(host(C), ref) < 1e-13


def test_gemm_lower_only():
    from gpx_b200 import ops

    rng = np.random.default_rng(1)
    A = rng.standard_normal((300, 70))
    C0 = rng.standard_normal((300, 300))
    C = ops.gemm(dev(A), dev(A), transb=True, alpha=-1.0, beta=1.0, C=dev(C0), lower_only=True)
    out = host(C)
    ref = C0 - A @ A.T
    assert rel(np.tril(out), np.tril(ref)) < 1e-13
    assert np.array_equal(np.triu(out, 1), np.triu(C0, 1))  # strictly-upper untouched


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, 8))
    return R.A_lhs("se", X, X, 2.0, 0.3)


@pytest.mark.parametrize("n", [5, 128, 129, 300, 1000, 2048])
def test_potrf_trsm_potri_logdet(n):
    from gpx_b200 import ops

    A = _spd(n, n)
    Lref = sla.cholesky(A, lower=True)
    Ad = dev(np.tril(A) + np.triu(np.full((n, n), np.nan), 1))  # upper triangle must never be read
    L, inv = ops.potrf(Ad)
    Lh = np.tril(host(L))
    assert rel(Lh, Lref) < 1e-11
    assert abs(float(host(ops.logdet(L))[0]) - np.sum(np.log(np.diag(Lref)))) < 1e-9 * n
    rng = np.random.default_rng(3)
    B = rng.standard_normal((7, n))
    X1 = host(ops.trsm(L, inv, dev(B), trans=True))
    assert rel(X1, sla.solve_triangular(Lref, B.T, lower=True).T) < 1e-9
    X2 = host(ops.trsm(L, inv, dev(B), trans=False))
    assert rel(X2, sla.solve_triangular(Lref, B.T, lower=True, trans="T").T) < 1e-9
    Ainv = np.tril(host(ops.potri(L, inv)))
    assert rel(Ainv, np.tril(np.linalg.inv(A))) < 1e-8


@pytest.mark.parametrize("n", [128, 300])
def test_trsm_in_place_many_rows(n):
    """The right solve multiplies B by the inverted diagonal blocks IN PLACE through the GEMM; with many rows and few
    columns the launch is small enough for the 64 x 64 tile rule, which must not apply to in-place products (two
    column tiles of one tile row would read and overwrite the same operand rows).  Repeated: the hazard is a race."""
    from gpx_b200 import ops

    A = _spd(n, 11 + n)
    Lref = sla.cholesky(A, lower=True)
    L, inv = ops.potrf(dev(A))
    rng = np.random.default_rng(4)
    B = rng.standard_normal((6000, n))
    ref_t = sla.solve_triangular(Lref, B.T, lower=True).T
    ref_n = sla.solve_triangular(Lref, B.T, lower=True, trans="T").T
    for _ in range(8):
        assert rel(host(ops.trsm(L, inv, dev(B), trans=True)), ref_t) < 1e-9
        assert rel(host(ops.trsm(L, inv, dev(B), trans=False)), ref_n) < 1e-9


@pytest.mark.parametrize("kind", R.KINDS)
@pytest.mark.parametrize("n1,n2,D", [(200, 333, 8), (129, 1, 3), (64, 64, 16), (300, 257, 32), (50, 70, 90)])
def test_gram_entries(kind, n1, n2, D):
    from gpx_b200 import ops

    rng = np.random.default_rng(5)
    x1, x2 = rng.standard_normal((n1, D)), rng.standard_normal((n2, D))
    ell = 1.7
    K = host(ops.gram(kind, dev(x1), dev(x2), ell))
    assert np.max(np.abs(K - R.kernel(kind, x1, x2, ell))) < 1e-13


@pytest.mark.parametrize("kind", R.KINDS)
def test_gram_symmetric_lower_diag(kind):
    from gpx_b200 import ops

    rng = np.random.default_rng(6)
    x = rng.standard_normal((300, 8))
    xd = dev(x)
    K = host(ops.gram(kind, xd, xd, 2.0, diag_add=0.25, lower_only=True))
    ref = R.kernel(kind, x, x, 2.0) + 0.25 * np.eye(300)
    tol = 1e-9 if kind == "m12" else 1e-13  # sqrt(1e-12 + eta) amplification on the diagonal
    assert np.max(np.abs(np.tril(K) - np.tril(ref))) < tol
    assert np.all(np.triu(K, 1) == 0.0)


CASES = [("se", 2000, 8, 2.0, 0.1, 0), ("m52", 1500, 16, 4.0, 0.1, 1), ("m32", 700, 5, 1.5, 0.3, 2),
         ("m12", 513, 3, 1.0, 0.5, 3)]


def _data(N, D, seed, T=1):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, D))
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, T))
    return x, y


@pytest.mark.parametrize("kind,N,D,ell,sigma,seed", CASES)
def test_gpr_lml_grad_parity(kind, N, D, ell, sigma, seed):
    from gpx_b200 import ops

    x, y = _data(N, D, seed)
    ref = np.array(R.lml_grad_dense(kind, x, y, ell, sigma))
    out = host(ops.gpr_lml_grad(kind, dev(x), dev(y), ell, sigma))
    assert abs(out[0] - ref[0]) <= 1e-8 * abs(ref[0])
    scale = max(np.max(np.abs(ref[1:])), abs(ref[0]))
    assert np.max(np.abs(out[1:] - ref[1:])) <= 1e-8 * scale
    lml_only = host(ops.gpr_lml_grad(kind, dev(x), dev(y), ell, sigma, want_grad=False))
    assert lml_only[0] == out[0]


def test_gpr_lml_multi_output_zero_mean():
    from gpx_b200 import ops

    x, y = _data(600, 4, 9, T=3)
    ref = np.array(R.lml_grad_dense("se", x, y, 1.5, 0.2, mean_function=R.zero_mean))
    out = host(ops.gpr_lml_grad("se", dev(x), dev(y), 1.5, 0.2, mean_mode=ops.MEAN_ZERO))
    assert np.max(np.abs(out - ref)) <= 1e-8 * np.max(np.abs(ref))


@pytest.mark.parametrize("kind,N,D,ell,sigma,seed", CASES[:2])
def test_gpr_fit_predict_parity(kind, N, D, ell, sigma, seed):
    from gpx_b200 import ops

    x, y = _data(N, D, seed)
    xs = np.random.default_rng(100 + seed).standard_normal((500, D))
    c_ref, mu_ref = R.fit_dense(kind, x, y, ell, sigma)
    mean_ref, cov_ref = R.predict_dense(kind, x, xs, c_ref, mu_ref, ell, sigma, full_covariance=True)
    c, mu = ops.gpr_fit(kind, dev(x), dev(y), ell, sigma)
    assert abs(float(host(mu)[0]) - mu_ref) < 1e-14
    assert rel(host(c), c_ref) < 1e-8
    mean, cov = ops.gpr_predict(kind, dev(x), dev(xs), c, mu, ell, sigma, full_covariance=True)
    assert rel(host(mean), mean_ref) < 1e-8
    var, var_ref = np.diag(host(cov)), np.diag(cov_ref)
    assert np.max(np.abs(var - var_ref) / np.maximum(var_ref, 1e-8)) < 1e-8
    assert np.max(np.abs(host(cov) - cov_ref)) < 1e-8
    mean_only = ops.gpr_predict(kind, dev(x), dev(xs), c, mu, ell, sigma)
    assert np.array_equal(host(mean_only), host(mean))


def test_nonpd_gives_nan_not_error():
    from gpx_b200 import ops

    A = -np.eye(64)
    L, _ = ops.potrf(dev(A))
    torch.cuda.synchronize()
    assert np.isnan(host(L)).any()


def test_random_shape_sweep_against_torch_fp64():
    """Seeded sweep over irregular sizes (panel / leaf / tile boundaries +-1, primes, tiny and multi-panel):
    Cholesky, both triangular solves (few and many right-hand sides), the SPD inverse and the GEMM with
    alpha / beta / lower_only against torch's FP64 linear algebra on the same device."""
    from gpx_b200 import ops

    rng = np.random.default_rng(2024)
    sizes = [1, 2, 127, 128, 129, 255, 257, 1023, 1024, 1025, 1151, 2047, 2049, 3071, 3201] + \
        [int(v) for v in rng.integers(3, 3400, 8)]
    g = torch.Generator(device="cuda").manual_seed(5)
    for n in sizes:
        G = torch.randn((n, n + 3), dtype=torch.float64, device="cuda", generator=g)
        A = G @ G.T / (n + 3) + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
        Lref = torch.linalg.cholesky(A)
        L, inv = ops.potrf(torch.tril(A).clone())
        assert float((torch.tril(L) - Lref).abs().max() / Lref.abs().max()) < 1e-11, n
        for m in (1, 3, 40):
            B = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
            X = ops.trsm(L, inv, B.clone(), trans=True)    # X L^T = B
            assert float((X @ Lref.T - B).abs().max()) < 1e-9 * max(1.0, float(B.abs().max())), (n, m)
            X = ops.trsm(L, inv, B.clone(), trans=False)   # X L = B
            assert float((X @ Lref - B).abs().max()) < 1e-9 * max(1.0, float(B.abs().max())), (n, m)
        Ainv = torch.tril(ops.potri(L, inv))
        ref = torch.tril(torch.cholesky_inverse(Lref))
        assert float((Ainv - ref).abs().max() / ref.abs().max()) < 1e-9, n
    for _ in range(12):
        M, N, K = (int(v) for v in rng.integers(1, 700, 3))
        ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device="cuda", generator=g)
        B = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device="cuda", generator=g)
        C0 = torch.randn((M, N), dtype=torch.float64, device="cuda", generator=g)
        alpha, beta = float(rng.standard_normal()), float(rng.standard_normal())
        ref = alpha * ((A.T if ta else A) @ (B.T if tb else B)) + beta * C0
        out = ops.gemm(A, B, transa=ta, transb=tb, alpha=alpha, beta=beta, C=C0.clone())
        assert float((out - ref).abs().max()) < 1e-11 * max(1.0, float(ref.abs().max())), (M, N, K, ta, tb)
        S = min(M, N)  # lower trapezoid of the leading S x S block
        out = ops.gemm(A[:, :S] if ta else A[:S], B[:S] if tb else B[:, :S], transa=ta, transb=tb, alpha=alpha,
                       beta=beta, C=C0[:S, :S].clone(), lower_only=True)
        assert float((torch.tril(out) - torch.tril(ref[:S, :S])).abs().max()) < 1e-11 * max(1.0, float(ref.abs().max()))
        assert torch.equal(torch.triu(out, 1), torch.triu(C0[:S, :S], 1))  # the strict upper triangle is untouched
