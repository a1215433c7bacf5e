"""Short programs for ncu: `matvec` = one block mat-vec; `rpchol` = a short RPCholesky loop."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402

what = sys.argv[1]
key = np.array([0, 2023], dtype=np.uint32)
g = torch.Generator(device="cuda").manual_seed(0)
if what == "matvec":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    x = torch.randn((N, 16), dtype=torch.float64, device="cuda", generator=g)
    V = torch.randn((N, 30), dtype=torch.float64, device="cuda", generator=g)
    W = ops.kmatvec("se", x, x, V, 4.0, diag_add=0.25)
else:
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    x = torch.randn((N, 32), dtype=torch.float64, device="cuda", generator=g)
    ops.rpcholesky("se", x, M, 32**0.5, key, True, want_factor=False)
torch.cuda.synchronize()
print("done")
