"""Generates tests/golden/oracle_small.npz from the oracle (the reference itself cannot run
here: JAX is absent).  The fixture freezes the oracle's outputs on a small seeded problem."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_numpy as R  # noqa: E402

rng = np.random.default_rng(42)
x = rng.standard_normal((160, 5))
y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((160, 1))
ell, sigma = 1.9, 0.15
out = {"x": x, "y": y, "ell": ell, "sigma": sigma}
for kind in R.KINDS:
    out[f"lml_grad_{kind}"] = np.array(R.lml_grad_dense(kind, x, y, ell, sigma))
out["pivots_part"] = R.rpcholesky("se", x, 16, ell, R.PRNGKey(2023), True)[1]
out["pivots_legacy"] = R.rpcholesky("se", x, 16, ell, R.PRNGKey(2023), False)[1]
c, mu = R.fit_dense("se", x, y, ell, sigma)
out["c_se"], out["mu"] = c, mu
xs = rng.standard_normal((20, 5))
m, C = R.predict_dense("se", x, xs, c, mu, ell, sigma, full_covariance=True)
out["xs"], out["pred_mean_se"], out["pred_cov_se"] = xs, m, C
np.savez(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "oracle_small.npz"), **out)
print("written")
