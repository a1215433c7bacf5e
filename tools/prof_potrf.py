"""Short program for the ncu launch list: one potrf + potri (+ optionally LML+grad) at size n."""
import sys

import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = sys.argv[2] if len(sys.argv) > 2 else "potrf"
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, 16), dtype=torch.float64, device="cuda", generator=g)
y = torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
if mode == "potrf":
    S = ops.gram("m52", x, x, 4.0, diag_add=0.01)
    L, inv = ops.potrf(S)
    ops.potri(L, inv)
else:
    out = ops.gpr_lml_grad("m52", x, y, 4.0, 0.1)
    print(out.cpu().numpy())
torch.cuda.synchronize()
print("done", ops.launches())
