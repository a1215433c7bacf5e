import torch, sys
sys.path.insert(0, ".")
from gpx_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for n in (128, 256, 1024):
    G = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
    S = G @ G.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    W = S.clone()
    for _ in range(3):
        W.copy_(S); ops.potrf(W)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        W.copy_(S)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.potrf(W); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(n, "potrf us", ts, "err", float((torch.tril(W) @ torch.tril(W).T - S).abs().max()))
