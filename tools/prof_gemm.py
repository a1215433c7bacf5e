"""Short program for `ncu --set full`: a few DMMA GEMM launches of the trailing-update shape."""
import sys

import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
A = torch.randn((m, k), dtype=torch.float64, device="cuda")
C = torch.zeros((m, m), dtype=torch.float64, device="cuda")
for _ in range(3):
    ops.gemm(A, A, transb=True, alpha=-1.0, beta=1.0, C=C, lower_only=True)
torch.cuda.synchronize()
print("done")
