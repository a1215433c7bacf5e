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
# --- added after the first fixture was frozen: every draw below comes after the ones above, so the arrays above
# are unchanged.  SGPR (fixed landmarks), a nested composite kernel, GPR_TD with two derivative blocks.
xl = x[::10].copy()
out["x_locs"] = xl
out["sgpr_lml_se"] = R.sgpr_lml_dense("se", xl, x, y, ell, sigma)
out["sgpr_lml_nxn_se"] = R.sgpr_lml_dense_nxn("se", xl, x, y, ell, sigma)
cs, _ = R.sgpr_fit_dense("se", xl, x, y, ell, sigma)
out["sgpr_c_se"] = cs
ms, Cs = R.sgpr_predict_dense("se", xl, xs, cs, mu, ell, sigma, full_covariance=True)
out["sgpr_pred_mean_se"], out["sgpr_pred_cov_se"] = ms, Cs
spec = ("sum", ("leaf", "se", 1.4, [0, 1]), ("prod", ("leaf", "m52", 2.2, None), ("leaf", "linear", None, [2, 3])))
out["composite_lml"] = R.composite_lml_dense(spec, x, y, sigma)
out["composite_sgpr_lml"] = R.sgpr_composite_lml_dense_nxn(spec, xl, x, y, sigma)
Nt, nv1, nv2 = 24, 3, 2
xt = x[:Nt].copy()
Jt = 0.1 * rng.standard_normal((Nt, 5, nv1 + nv2))
for k in range(nv1 + nv2):
    Jt[:, k, k] += 1.0
ydt = rng.standard_normal((Nt, nv1 + nv2))
out["td_J"], out["td_yd"] = Jt, ydt
out["td_lml_se"] = R.td_lml_dense("se", xt, y[:Nt], Jt, ydt, ell, 0.3, 0.2)
out["td2_lml_m52"] = R.td2_lml_dense("m52", xt, y[:Nt], Jt[:, :, :nv1], Jt[:, :, nv1:], ydt[:, :nv1], ydt[:, nv1:], ell,
                                      0.3, 0.2, 0.4)
ct, _ = R.td2_fit_dense("m52", xt, y[:Nt], Jt[:, :, :nv1], Jt[:, :, nv1:], ydt[:, :nv1], ydt[:, nv1:], ell, 0.3, 0.2, 0.4)
out["td2_c_m52"] = ct
np.savez(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "oracle_small.npz"), **out)
print("written")
