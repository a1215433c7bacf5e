"""Throughput of the matrix-free / sparse paths at the BASELINE.json shapes (configs[3], [4]),
one JSON line per measurement, each with its roofline figure.

    python tools/bench_configs.py rpchol [N] [M]     # configs[3] part 1 (HBM bound)
    python tools/bench_configs.py sgpr   [N] [M]     # configs[3] part 2 (FP64 tensor bound)
    python tools/bench_configs.py matvec [N] [P]     # configs[4]: one block mat-vec (FP64 pipe bound)
    python tools/bench_configs.py slq    [N] [P] [m] # configs[4]: Lanczos/SLQ log-det end to end
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402

HBM = 6554.2  # GB/s, MEASURED_PEAKS.json


def ev_time(fn, reps=1, warm=0):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps, out


def data(N, D, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((N, D), dtype=torch.float64, device="cuda", generator=g)
    y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn((N, 1), dtype=torch.float64, device="cuda", generator=g)
    return x, y


def main():
    what = sys.argv[1]
    a = [int(v) for v in sys.argv[2:] if v.lstrip('-').isdigit()]
    key = np.array([0, 2023], dtype=np.uint32)
    if what == "rpchol":
        N, M = (a + [1_000_000, 4096])[:2] if len(a) < 2 else a[:2]
        x, _ = data(N, 32, 3)
        t, (F, piv) = ev_time(lambda: ops.rpcholesky("se", x, M, 32**0.5, key, True, want_factor=False))
        alg = 8.0 * N * M * (M - 1) / 2 + 5 * 8.0 * N * M
        print(json.dumps(dict(what="rpcholesky", N=N, M=M, D=32, s=t, alg_bytes=alg, gbs=alg / t / 1e9,
                              hbm_frac=alg / t / 1e9 / HBM, pivots_head=piv[:6].cpu().tolist())), flush=True)
    elif what == "sgpr":
        N, M = a[:2] if len(a) >= 2 else (1_000_000, 4096)
        x, y = data(N, 32, 3)
        _, piv = ops.rpcholesky("se", x[:65536].contiguous(), M, 32**0.5, key, True, want_factor=False)
        xl = ops.gather_rows(x, piv)
        t, out = ev_time(lambda: ops.sgpr_lml("se", x, y, xl, 32**0.5, 0.1), reps=1, warm=1)
        fl = 2.0 * N * M * 32 + 2.0 * M * M * N + 2.0 * M**3 / 3
        print(json.dumps(dict(what="sgpr_lml_woodbury", N=N, M=M, D=32, s=t, alg_flops=fl, tflops=fl / t / 1e12,
                              lml=float(out.cpu()[0]))), flush=True)
    elif what == "matvec":
        N, P = a[:2] if len(a) >= 2 else (2_000_000, 30)
        x, _ = data(N, 16, 4)
        V = torch.randn((N, P), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
        t, W = ev_time(lambda: ops.kmatvec("se", x, x, V, 4.0, diag_add=0.25 + 1e-10), reps=1,
                       warm=1 if N <= 600_000 else 0)
        fl = float(N) ** 2 * (2 * 16 + 4 + 2 * P)
        print(json.dumps(dict(what="kmatvec_block", N=N, D=16, P=P, s=t, alg_flops=fl, tflops=fl / t / 1e12,
                              pair_evals_per_s=float(N) ** 2 / t, checksum=float(W.sum().cpu()))), flush=True)
    elif what == "derivs":
        # configs[2]: GPR on derivative values, nf = nv = 90 (30 atoms), N = 2000 -> 180 000 rows: 259 GB as a
        # square matrix, 131 GB in the block-column packed lower storage the library switches to by itself.
        # `derivs N check` also runs the fit and verifies it through the residual |A c - y| with A rebuilt in
        # row blocks (an independent pass over the Hessian kernel; nothing of the factorisation is reused).
        N = a[0] if a else 2000
        check = "check" in sys.argv[2:]
        nf = nv = 90
        ell, sigma = 90**0.5, 0.05
        g = torch.Generator(device="cuda").manual_seed(2)
        x = torch.randn((N, nf), dtype=torch.float64, device="cuda", generator=g)
        J = 0.1 * torch.randn((N, nf, nv), dtype=torch.float64, device="cuda", generator=g)
        J += torch.eye(nf, dtype=torch.float64, device="cuda").unsqueeze(0)
        y = torch.randn((N, nv), dtype=torch.float64, device="cuda", generator=g)
        n = N * nv
        t, out = ev_time(lambda: ops.gpr_lml_derivs("se", x, J, y, ell, sigma), warm=1 if "warm" in sys.argv else 0)
        fl = n**3 / 3.0 + float(n) ** 2 * nf
        import os
        print(json.dumps(dict(what="gpr_lml_derivs", N=N, nf=nf, nv=nv, rows=n, s=t, alg_flops=fl,
                              tflops=fl / t / 1e12, square_gb=8.0 * n * n / 1e9,
                              layout=os.environ.get("GPXB_DERIVS_LAYOUT", "auto"),
                              peak_alloc_gb=(torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9,
                              lml=float(out.cpu()[0]))), flush=True)
        if check:
            t, (c, jc) = ev_time(lambda: ops.gpr_fit_derivs("se", x, J, y, ell, sigma))
            cf, yf = c.reshape(-1), y.reshape(-1)
            s2 = sigma * sigma + 1e-10
            blk = 50
            num = 0.0
            for s0 in range(0, N, blk):
                s1 = min(N, s0 + blk)
                K = ops.hess_gram("se", x[s0:s1], J[s0:s1], x, J, ell)
                r = K @ cf + s2 * cf[s0 * nv:s1 * nv] - yf[s0 * nv:s1 * nv]
                num += float((r * r).sum().cpu())
                del K
            print(json.dumps(dict(what="gpr_fit_derivs", N=N, rows=n, s=t,
                                  rel_residual=(num ** 0.5) / float(yf.norm().cpu()),
                                  c_norm=float(cf.norm().cpu()))), flush=True)
    elif what == "derivs_grad":
        # configs[2] shape, LML + analytic gradient (needs the inverse: square storage, three n x n buffers)
        N = a[0] if a else 400
        nf = nv = 90
        ell, sigma = 90**0.5, 0.05
        g = torch.Generator(device="cuda").manual_seed(2)
        x = torch.randn((N, nf), dtype=torch.float64, device="cuda", generator=g)
        J = 0.1 * torch.randn((N, nf, nv), dtype=torch.float64, device="cuda", generator=g)
        J += torch.eye(nf, dtype=torch.float64, device="cuda").unsqueeze(0)
        y = torch.randn((N, nv), dtype=torch.float64, device="cuda", generator=g)
        n = N * nv
        t, out = ev_time(lambda: ops.gpr_lml_grad_derivs("se", x, J, y, ell, sigma), warm=1)
        fl = float(n) ** 3 + 2.0 * float(n) ** 2 * nf * 2
        print(json.dumps(dict(what="gpr_lml_grad_derivs", N=N, nf=nf, nv=nv, rows=n, s=t, alg_flops=fl,
                              tflops=fl / t / 1e12, out=[float(v) for v in out.cpu()])), flush=True)
    elif what == "derivs_iter":
        # configs[2] at its named size through the matrix-free Hessian operator: iterative LML + gradient
        # (PCG with the rpcholesky_derivs preconditioner, 30-probe / 50-step SLQ with tangents)
        from gpx_b200.models import _iterative as it
        N = a[0] if a else 2000
        npiv = a[1] if len(a) > 1 else 200
        nf = nv = 90
        ell, sigma = 90**0.5, 0.05
        g = torch.Generator(device="cuda").manual_seed(2)
        x = torch.randn((N, nf), dtype=torch.float64, device="cuda", generator=g)
        J = 0.1 * torch.randn((N, nf, nv), dtype=torch.float64, device="cuda", generator=g)
        J += torch.eye(nf, dtype=torch.float64, device="cuda").unsqueeze(0)
        y = torch.randn((N, nv), dtype=torch.float64, device="cuda", generator=g)
        key = np.array([0, 2023], dtype=np.uint32)
        t0 = time.perf_counter()
        c, jc, iters = ops.gpr_fit_derivs_iter("se", x, J, y, ell, sigma, npiv, key)
        torch.cuda.synchronize()
        t_fit = time.perf_counter() - t0
        t0 = time.perf_counter()
        lml = it.lml_derivs_iter("se", x, J, y, ell, sigma, 30, 50, key, npiv, key)
        torch.cuda.synchronize()
        t_lml = time.perf_counter() - t0
        t0 = time.perf_counter()
        out = it.lml_derivs_iter_grad("se", x, J, y, ell, sigma, 30, 50, key, npiv, key)
        torch.cuda.synchronize()
        t_grad = time.perf_counter() - t0
        print(json.dumps(dict(what="derivs_iterative", N=N, rows=N * nv, n_pivots=npiv, pcg_iters=iters, s_fit=t_fit,
                              s_lml=t_lml, s_lml_grad=t_grad, lml=lml, lml_grad=list(out))), flush=True)
    elif what == "lml_iter":
        # configs[4] end to end at a reduced N (the named N = 2e6 costs (2e6/N)^2 more per mat-vec):
        # PCG fit (RPCholesky preconditioner, n_pivots = 100) + 30-probe, 50-step SLQ log-det
        N = a[0] if a else 262144
        x, y = data(N, 16, 4)
        import time as _t

        from gpx_b200.models._iterative import slq_logdet

        torch.cuda.synchronize()
        t0 = _t.perf_counter()
        c, mu, iters = ops.gpr_fit_iter("se", x, y, 4.0, 0.5, 100, key)
        torch.cuda.synchronize()
        t1 = _t.perf_counter()
        U = ops.rademacher_probes(key, N, 30, True)
        al, be, fl_ = ops.lanczos("se", x, 4.0, 0.25 + 1e-10, U, 50)
        ld = slq_logdet(al.cpu().numpy(), be.cpu().numpy(), N)
        r = (y - mu).cpu().numpy()
        lml = (-0.5 * float(r.T @ c.cpu().numpy()) - 0.5 * ld - N * 0.5 * np.log(2 * np.pi)) / N
        t2 = _t.perf_counter()
        print(json.dumps(dict(what="gpr_lml_iter", N=N, D=16, P=30, m=50, n_pivots=100, pcg_iters=iters,
                              s_pcg=t1 - t0, s_slq=t2 - t1, s_total=t2 - t0, lml=lml,
                              matvec_flops=(50 * (2 * 16 + 4 + 60) + iters * (2 * 16 + 4 + 2)) * float(N) ** 2,
                              breakdown=int(fl_.cpu()[0]))), flush=True)
    elif what == "slq":
        N, P, m = a[:3] if len(a) >= 3 else (262144, 30, 50)
        x, _ = data(N, 16, 4)
        U = ops.rademacher_probes(key, N, P, True)
        t, (al, be, fl_) = ev_time(lambda: ops.lanczos("se", x, 4.0, 0.25 + 1e-10, U, m))
        from gpx_b200.models._iterative import slq_logdet

        ld = slq_logdet(al.cpu().numpy(), be.cpu().numpy(), N)
        fl = m * float(N) ** 2 * (2 * 16 + 4 + 2 * P)
        print(json.dumps(dict(what="lanczos_slq", N=N, D=16, P=P, m=m, s=t, matvec_flops=fl, tflops=fl / t / 1e12,
                              logdet=ld, breakdown=int(fl_.cpu()[0]))), flush=True)


if __name__ == "__main__":
    main()
