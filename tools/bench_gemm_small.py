"""Small-K GEMM shapes of the panel factorisation and of the triangular inverse at mid size: time per launch with
128 x 128 tiles (GPXB_GEMM_SMALL_MAX=0) and with 64 x 64 tiles.  One JSON line per shape."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


g = torch.Generator(device="cuda").manual_seed(0)
shapes = [  # (M, N, K, lower)
    (1872, 896, 128, True), (3968, 896, 128, True), (8064, 896, 128, True), (3968, 3968, 128, True),
    (1872, 128, 128, False), (3968, 128, 128, False), (7936, 128, 128, False),
    (2048, 2048, 1024, True), (1024, 1024, 1024, False), (2048, 2048, 2048, False), (4096, 4096, 1024, True),
]
for M, N, K, lower in shapes:
    A = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g)
    C = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    t = timeit(lambda: ops.gemm(A, B, transb=True, alpha=-1.0, beta=1.0, C=C, lower_only=lower))
    area = (N * (N + 1) / 2 + (M - N) * N) if lower else M * N
    fl = 2.0 * area * K
    print(json.dumps(dict(what="gemm_small", M=M, N=N, K=K, lower=lower,
                          small_max=os.environ.get("GPXB_GEMM_SMALL_MAX", "default"),
                          us=t * 1e6, tflops=fl / t / 1e12)), flush=True)
