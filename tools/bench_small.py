"""Latency of the mid-size dense path (BASELINE configs[0]: exact GPR, SE, N = 2000, D = 8, and n = 4096 / 8192):
potrf against cuSOLVER on the same box, LML+grad, fit (two few-right-hand-side triangular solves).  One JSON line
per measurement; CUDA events, best of 5 after 2 warm-up calls."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for n, D, kind, ell in ((2000, 8, "se", 2.0), (4096, 16, "m52", 4.0), (8192, 16, "m52", 4.0)):
        x = torch.randn((n, D), dtype=torch.float64, device="cuda", generator=g)
        y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
        S = ops.gram(kind, x, x, ell, diag_add=0.01)
        W = torch.empty_like(S)
        t_copy, _ = timeit(lambda: W.copy_(S))
        t_cus, _ = timeit(lambda: torch.linalg.cholesky(S, out=W))

        def ours():
            W.copy_(S)
            ops.potrf(W)

        t_our, _ = timeit(ours)
        t_lml, _ = timeit(lambda: ops.gpr_lml_grad(kind, x, y, ell, 0.1))
        t_val, _ = timeit(lambda: ops.gpr_lml_grad(kind, x, y, ell, 0.1, want_grad=False))
        t_fit, _ = timeit(lambda: ops.gpr_fit(kind, x, y, ell, 0.1))
        fl = n**3 + 2.0 * n * n * D
        print(json.dumps(dict(what="mid_size_dense", n=n, D=D, kind=kind, cusolver_potrf_ms=t_cus,
                              gpxb_potrf_ms=t_our - t_copy, lml_grad_ms=t_lml, lml_only_ms=t_val, fit_ms=t_fit,
                              lml_grad_tflops=fl / t_lml / 1e9)), flush=True)


if __name__ == "__main__":
    main()
