"""Micro-benchmarks on one B200: measured cuBLAS/cuSOLVER FP64 denominators next to
libgpxb200's GEMM / Cholesky / LML.  Prints one JSON line per measurement."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402


def timeit(fn, reps=3, warm=1):
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
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


def out(**kw):
    print(json.dumps(kw), flush=True)


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [8192, 16384]
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in sizes:
        A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
        B = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
        C = torch.empty_like(A)
        fl = 2.0 * n**3
        t, tm = timeit(lambda: torch.matmul(A, B.T, out=C))
        out(what="cublas_dgemm_nt", n=n, s=t, tflops=fl / t / 1e12, tflops_median=fl / tm / 1e12)
        t, tm = timeit(lambda: ops.gemm(A, B, transb=True, C=C))
        out(what="gpxb_gemm_nt", n=n, s=t, tflops=fl / t / 1e12, tflops_median=fl / tm / 1e12)
        t, tm = timeit(lambda: ops.gemm(A, B, C=C))
        out(what="gpxb_gemm_nn", n=n, s=t, tflops=fl / t / 1e12)
        t, tm = timeit(lambda: ops.gemm(A, A, transb=True, C=C, lower_only=True))
        out(what="gpxb_syrk_lower", n=n, s=t, tflops=fl / 2 / t / 1e12)
        del B
        # SPD matrix
        x = torch.randn((n, 16), dtype=torch.float64, device="cuda", generator=g)
        S = ops.gram("m52", x, x, 4.0, diag_add=0.01)
        W = torch.empty_like(S)

        def cusolver():
            torch.linalg.cholesky(S, out=W)

        t, tm = timeit(cusolver, reps=2)
        out(what="cusolver_potrf", n=n, s=t, tflops=n**3 / 3 / t / 1e12)

        def ours():
            W.copy_(S)
            ops.potrf(W)

        t0, _ = timeit(lambda: W.copy_(S), reps=2)
        t, tm = timeit(ours, reps=2)
        out(what="gpxb_potrf", n=n, s=t - t0, tflops=n**3 / 3 / (t - t0) / 1e12)
        L, inv = ops.potrf(W.copy_(S))
        t, tm = timeit(lambda: ops.potri(L, inv, out=C), reps=2)
        out(what="gpxb_potri", n=n, s=t, tflops=2 * n**3 / 3 / t / 1e12)
        t, tm = timeit(lambda: ops.gram("m52", x, x, 4.0, diag_add=0.01, lower_only=True, out=C), reps=3)
        out(what="gpxb_gram_m52_lower", n=n, s=t, gbs_written=8.0 * n * n / 2 / t / 1e9)
        del A, C, S, W, L, inv
        y = torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
        t, tm = timeit(lambda: ops.gpr_lml_grad("m52", x, y, 4.0, 0.1), reps=2)
        fl = n**3 + 2.0 * n * n * 16
        out(what="gpxb_lml_grad_m52", n=n, s=t, evals_per_s=1 / t, tflops=fl / t / 1e12)
        t, tm = timeit(lambda: ops.gpr_lml_grad("m52", x, y, 4.0, 0.1, want_grad=False), reps=2)
        out(what="gpxb_lml_only_m52", n=n, s=t, tflops=(n**3 / 3) / t / 1e12)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
