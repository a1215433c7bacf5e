"""Three LML+grad evaluations of BASELINE configs[0] (exact GPR, SE, N = 2000, D = 8): the command behind the ncu
launch list profiles/r2_launches_cfg0_lml_grad_N2000.csv."""
import sys

import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((2000, 8), dtype=torch.float64, device="cuda", generator=g)
y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn((2000, 1), dtype=torch.float64, device="cuda", generator=g)
for _ in range(3):
    out = ops.gpr_lml_grad("se", x, y, 2.0, 0.1)
torch.cuda.synchronize()
print(out[0])
