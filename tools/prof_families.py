"""Short single-purpose programs for `ncu --set full -k regex:<kernel>` captures of every kernel family (round 2):

    python tools/prof_families.py gram      # gram_kernel store + contraction (LML+grad, Matern-5/2, N = 8192, D = 16)
    python tools/prof_families.py gram_kt   # K-tiled Gram (D = 256)
    python tools/prof_families.py hess      # hess_kernel (30 atoms: nf = nv = 90, 160 configurations)
    python tools/prof_families.py potrf     # potrf_leaf_dmma_kernel + panel GEMMs (n = 4096)
    python tools/prof_families.py fit       # trsv_* (gpr_fit, one right-hand side, N = 8192)
    python tools/prof_families.py rpchol    # rp_gemv_kernel / rp_finish_kernel (N = 262144, 600 pivots)
    python tools/prof_families.py matvec1   # single-vector mat-vec (P = 1 FMA path), N = 65536
    python tools/prof_families.py matvec    # block mat-vec (P = 30), N = 65536
    python tools/prof_families.py gemm64    # gemm_f64_kernel with 64 x 64 tiles (1024^3 and the skinny K = 128 product)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402

what = sys.argv[1]
key = np.array([0, 2023], dtype=np.uint32)
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(s, dtype=torch.float64, device="cuda", generator=g)  # noqa: E731
if what == "gram":
    x = rn(8192, 16)
    y = torch.sin(x.sum(1, keepdim=True))
    for _ in range(2):
        ops.gpr_lml_grad("m52", x, y, 4.0, 0.1)
elif what == "gram_kt":
    x = rn(8192, 256)
    for _ in range(2):
        ops.gram("m52", x, x, 16.0, diag_add=0.01, lower_only=True)
elif what == "hess":
    x, J = rn(160, 90), rn(160, 90, 90)
    for _ in range(2):
        ops.hess_gram("se", x, J, x, J, 9.0, diag_add=0.01, lower_only=True)
elif what == "potrf":
    x = rn(4096, 16)
    A = ops.gram("se", x, x, 4.0, diag_add=0.5)
    for _ in range(2):
        ops.potrf(A.clone())
elif what == "fit":
    x = rn(8192, 16)
    y = torch.sin(x.sum(1, keepdim=True))
    for _ in range(2):
        ops.gpr_fit("se", x, y, 4.0, 0.1)
elif what == "rpchol":
    x = rn(262144, 32)
    ops.rpcholesky("se", x, 600, 32**0.5, key, True, want_factor=False)
elif what in ("matvec", "matvec1"):
    N = 65536
    x = rn(N, 16)
    V = rn(N, 30 if what == "matvec" else 1)
    for _ in range(2):
        ops.kmatvec("se", x, x, V, 4.0, diag_add=0.25)
elif what == "gemm64":
    A, B = rn(1024, 1024), rn(1024, 1024)
    C = torch.zeros((1024, 1024), dtype=torch.float64, device="cuda")
    for _ in range(3):
        ops.gemm(A, B, transb=True, alpha=-1.0, beta=1.0, C=C)
else:
    raise SystemExit(f"unknown case {what}")
torch.cuda.synchronize()
print("done")
