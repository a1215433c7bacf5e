"""Trailing-update shape of the blocked Cholesky: C[n x n, lower] -= P P^T with P n x 1024 (and the square SYRK
n x n x n).  Prints TFLOP/s; GPXB_GEMM_TRI_GROUP selects the raster of the triangular tile map."""
import json
import os
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
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


g = torch.Generator(device="cuda").manual_seed(0)
for n, k in ((16384, 1024), (8192, 1024), (32768 - 1024, 1024), (8192, 8192)):
    P = torch.randn((n, k), dtype=torch.float64, device="cuda", generator=g)
    C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    t, tm = timeit(lambda: ops.gemm(P, P, transb=True, alpha=-1.0, beta=1.0, C=C, lower_only=True))
    fl = 2.0 * (n * (n + 1) / 2) * k
    print(json.dumps(dict(what="syrk_lower_update", n=n, k=k, tri_group=os.environ.get("GPXB_GEMM_TRI_GROUP", "8"),
                          s=t, tflops=fl / t / 1e12, tflops_median=fl / tm / 1e12)), flush=True)
    del P, C
