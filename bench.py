#!/usr/bin/env python
"""bench.py -- headline benchmark of the GP hot path on B200.

Workload (BASELINE.json configs[1]): exact GPR, Matern-5/2, N=32768, D=16, FP64;
one "step" = one LML + gradient evaluation (Gram build, blocked Cholesky, triangular
solves, SPD inverse, fused dK/d-lengthscale contraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu] [--no-extra]

At N = 1 the JSON line also carries `other_configs`: RPCholesky, the SGPR Woodbury LML and one block mat-vec of the
Lanczos / SLQ log-det at N = 1e6 (BASELINE configs[3], [4]) with their rooflines, measured after the headline.

N > 1 (torchrun): the dense Cholesky path does not shard ("replicas only") --
every rank evaluates an independent replica (its own hyper-parameter point, as the
reference's random restarts do) and `value` is the whole-job evaluations per second.

`--impl reference` times the CPU restatement of the reference (oracle/ref_numpy.py; the
reference itself needs JAX, which this image does not have) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(kind="m52", N=32768, D=16, ell=4.0, sigma=0.1, seed=1)
CPU_SAMPLE_N = 4096


def synth(N, D, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, D))
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    return x, y


def alg_flops(N, D):
    return float(N) ** 3 + 2.0 * float(N) ** 2 * D


# ----------------------------------------------------------------------------- CPU arm
def cpu_sample_seconds(steps, warmup, n0):
    """Seconds per LML+grad of the CPU restatement with every host core the process may use
    (torchrun exports OMP_NUM_THREADS=1; the BLAS pool is widened explicitly)."""
    from threadpoolctl import threadpool_limits

    from oracle import ref_numpy as R  # the checker, timed as the CPU baseline

    x, y = synth(n0, CFG["D"], CFG["seed"])
    nthr = cpu_cores()
    with threadpool_limits(limits=nthr):
        for _ in range(warmup):
            R.lml_grad_dense(CFG["kind"], x, y, CFG["ell"], CFG["sigma"], threads=nthr)
        t0 = time.perf_counter()
        for _ in range(steps):
            R.lml_grad_dense(CFG["kind"], x, y, CFG["ell"], CFG["sigma"], threads=nthr)
        return (time.perf_counter() - t0) / steps


def cpu_extrapolated(steps, warmup, n0, N):
    """(seconds per LML+grad at N, description): the oracle timed at n0 and extrapolated term by term.  At n0 = 4096
    most of the oracle's time is O(N^2) work (kernel entries, exp, dK/d ell, the W o dK contraction); only the
    LAPACK factorisation and inverse are O(N^3).  The cubic part is therefore timed separately (dpotrf + dpotri on
    a matrix of the same size, same thread pool) and scaled by (N/n0)^3, the remainder by (N/n0)^2 -- a plain
    (N/n0)^3 on the whole would overstate the CPU time severalfold."""
    import scipy.linalg as sla
    from threadpoolctl import threadpool_limits

    t_total = cpu_sample_seconds(steps, warmup, n0)
    rng = np.random.default_rng(0)
    B = rng.standard_normal((n0, 64))
    A = B @ B.T + n0 * np.eye(n0)
    with threadpool_limits(limits=cpu_cores()):
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            L = sla.cholesky(A, lower=True, check_finite=False)
            sla.lapack.dpotri(L, lower=1)
            best = min(best, time.perf_counter() - t0)
    t_cubic = min(best, t_total)
    r = N / n0
    t_N = t_cubic * r**3 + (t_total - t_cubic) * r**2
    desc = (f"oracle/ref_numpy.py (NumPy/OpenBLAS restatement of GPX _lml_dense + analytic gradient; JAX absent) at "
            f"N={n0}, D={CFG['D']}: {t_total:.3f} s per LML+grad, of which dpotrf+dpotri {t_cubic:.3f} s; extrapolated to "
            f"N={N} as cubic part x(N/{n0})^3 + remainder x(N/{n0})^2 = {t_N:.1f} s")
    return t_N, desc


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = args.n or CFG["N"]
    n0 = min(CPU_SAMPLE_N, N)
    t_N, sample = cpu_extrapolated(max(args.steps, 1), min(args.warmup, 1), n0, N)
    value = 1.0 / t_N
    line = {
        "impl": "reference", "metric": "LML+grad evals/s (exact GPR, Matern-5/2, N=32768, D=16, FP64)",
        "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_N * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"exact GPR Matern-5/2 N={N} D={CFG['D']} LML+grad (BASELINE configs[1])",
                   "parallelism": "host cores (CPU restatement of the reference)", "cpu_sample_N": n0},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cpu_cores(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().strip().splitlines():
            c = [a.strip() for a in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])), mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling
    recipe's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6554.2, "fallback (B200_PROFILING.md)"


def other_configs(dgemm_tflops, scale=1.0):
    """The SGPR / Lanczos half of BASELINE.json's metric line (configs[3], configs[4]) at N = 1e6 on this GPU, one
    measurement each, device-timed with CUDA events on torch's current stream (the stream the library launches on);
    inputs are created on the device and are far larger than L2.  Reported next to the headline, never instead of
    it; every entry is isolated (a failure is reported as {"error": ...} and cannot affect the main line, whose
    numbers are already on the host).  `scale` shrinks the sizes for smoke tests (GPXB_BENCH_EXTRA_SCALE)."""
    import torch

    from gpx_b200 import ops

    key = np.array([0, 2023], dtype=np.uint32)
    N = max(4096, int(1_000_000 * scale))
    M = 4096 if scale >= 1.0 else max(64, int(4096 * scale * 4) // 64 * 64)
    hbm, hbm_src = hbm_peak_gbs()
    out = []

    def timed(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3, res

    def data(n, d, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        x = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
        y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
        return x, y

    def entry(name, fn):
        try:
            out.append(dict(config=name, **fn()))
        except Exception as e:  # noqa: BLE001 -- reported, never fatal for the headline
            out.append(dict(config=name, error=f"{type(e).__name__}: {e}"[:300]))

    x32, y32 = data(N, 32, 3)
    ell32 = 32**0.5
    state = {}

    def rpchol():
        t, (_, piv) = timed(lambda: ops.rpcholesky("se", x32, M, ell32, key, True, want_factor=False))
        state["piv"], state["t_rp"] = piv, t
        alg = 8.0 * N * M * (M - 1) / 2 + 5 * 8.0 * N * M
        return dict(what="RPCholesky landmark loop (pivots only)", N=N, D=32, M=M, s=t, bound="hbm",
                    alg_bytes=alg, achieved_gbs=alg / t / 1e9, peak_gbs=hbm, peak_source=hbm_src,
                    frac=alg / t / 1e9 / hbm, pivots_head=[int(v) for v in piv[:4].cpu().tolist()])

    def sgpr():
        piv = state.get("piv")
        if piv is None:
            _, piv = ops.rpcholesky("se", x32[:65536].contiguous(), M, ell32, key, True, want_factor=False)
        xl = ops.gather_rows(x32, piv)
        ops.sgpr_lml("se", x32, y32, xl, ell32, 0.1)  # warm-up (allocator pools)
        t, res = timed(lambda: ops.sgpr_lml("se", x32, y32, xl, ell32, 0.1))
        fl = 2.0 * N * M * 32 + 2.0 * M * M * N + 2.0 * M**3 / 3
        d = dict(what="SGPR Woodbury LML (fixed landmarks)", N=N, D=32, M=M, s=t, evals_per_s=1.0 / t,
                 bound="tensor", alg_flops=fl, achieved_tflops=fl / t / 1e12, peak_tflops=dgemm_tflops,
                 frac=fl / t / 1e12 / dgemm_tflops, lml=float(res.cpu()[0]))
        if "t_rp" in state:  # SGPR_RPChol re-draws the landmarks at every LML evaluation (_sgpr_rpchol.py:193-200)
            d["evals_per_s_with_landmark_selection"] = 1.0 / (t + state["t_rp"])
        return d

    def lanczos_step():
        P = 30
        x16, _ = data(N, 16, 4)
        V = torch.randn((N, P), dtype=torch.float64, device="cuda",
                        generator=torch.Generator(device="cuda").manual_seed(9))
        t, W = timed(lambda: ops.kmatvec("se", x16, x16, V, 4.0, diag_add=0.25 + 1e-10))
        fl = float(N) ** 2 * (2 * 16 + 4 + 2 * P)
        return dict(what="matrix-free block mat-vec (K + s2 I) V: one of the 50 steps of each 30-probe SLQ log-det",
                    N=N, D=16, P=P, s=t, bound="tensor", alg_flops=fl, achieved_tflops=fl / t / 1e12,
                    peak_tflops=dgemm_tflops, frac=fl / t / 1e12 / dgemm_tflops, checksum=float(W.sum().cpu()))

    entry(f"configs[3] SGPR + RPCholesky, N={N} D=32 M={M}: landmark selection", rpchol)
    entry(f"configs[3] SGPR + RPCholesky, N={N} D=32 M={M}: LML", sgpr)
    entry(f"configs[4] Lanczos/SLQ, SE kernel, D=16, 30 probes, at N={N}", lanczos_step)
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from gpx_b200 import ops
    from gpx_b200.kernels import Matern52
    from gpx_b200.models import GPR
    from gpx_b200.parameters import Parameter
    from gpx_b200.bijectors import Softplus
    from gpx_b200.priors import GammaPrior

    N, D = args.n or CFG["N"], CFG["D"]
    kind, sigma = CFG["kind"], CFG["sigma"]
    ell = CFG["ell"] * (1.0 + 0.05 * rank)  # replicas evaluate different hyper-parameter points
    x_np, y_np = synth(N, D, CFG["seed"])
    x = torch.as_tensor(x_np).cuda()
    y = torch.as_tensor(y_np).cuda()
    W, K = max(args.warmup, 3), max(args.steps, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = None
    for _ in range(W):
        out = ops.gpr_lml_grad(kind, x, y, ell, sigma)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ops.launches()
    ops.prof_enable(True)  # brackets every DMMA GEMM launch with CUDA events on its own stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = ops.gpr_lml_grad(kind, x, y, ell, sigma)
    e1.record()
    barrier()
    ops.prof_enable(False)
    reg_ms, reg_flops, reg_n = ops.prof_collect()
    launches = ops.launches() - l0
    # one more, untimed, step with the look-ahead serialised: no two launches overlap, so the summed event
    # durations are pure kernel time (in the timed region the panel factorisation runs concurrently with the
    # trailing update and overlapped time is counted twice)
    ops.prof_enable(2)
    ops.gpr_lml_grad(kind, x, y, ell, sigma)
    torch.cuda.synchronize()
    ops.prof_enable(False)
    gemm_ms, gemm_flops, gemm_n = ops.prof_collect()
    ms = max_over_ranks(e0.elapsed_time(e1)) / K
    clocks = sampler.stop() if sampler else None
    res = out.cpu().numpy()

    # ---- end to end through the public facade with (pinned) HOST buffers ----
    model = GPR(Matern52(),
                kernel_params=dict(lengthscale=Parameter(ell, True, Softplus(), GammaPrior())),
                sigma=Parameter(sigma, True, Softplus(), GammaPrior()))
    xh, yh = torch.as_tensor(x_np).pin_memory(), torch.as_tensor(y_np).pin_memory()
    model.loss_and_grad(xh, yh)
    barrier()
    Ke = max(1, min(K, 5))
    t0 = time.perf_counter()
    for _ in range(Ke):
        loss, grads = model.loss_and_grad(xh, yh)  # H2D of x, y; D2H of (lml, grad)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / Ke)
    assert abs(-loss - res[0]) <= 1e-12 * abs(res[0])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (DMMA GEMM, trailing-update shape), timed live ----
    m_t, k_t = min(N - 2048, 16384) if N > 4096 else N, 1024 if N > 4096 else min(N, 1024)
    A = torch.randn((m_t, k_t), dtype=torch.float64, device="cuda")
    C = torch.zeros((m_t, m_t), dtype=torch.float64, device="cuda")
    for _ in range(2):
        ops.gemm(A, A, transb=True, alpha=-1.0, beta=1.0, C=C, lower_only=True)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a0.record()
    for _ in range(reps):
        ops.gemm(A, A, transb=True, alpha=-1.0, beta=1.0, C=C, lower_only=True)
    a1.record()
    torch.cuda.synchronize()
    t_k = a0.elapsed_time(a1) * 1e-3 / reps
    tiles = (m_t + 127) // 128
    flops_k = tiles * (tiles + 1) / 2 * 128 * 128 * k_t * 2.0
    probe = flops_k / t_k / 1e12
    # dominant kernel over the timed region: algorithmic flops of all its launches / their summed durations
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    del A, C
    # FP64 tensor denominator: MEASURED_PEAKS.json has no FP64 entry -> cuBLAS Dgemm measured here
    nb = 8192 if N >= 8192 else 4096
    P, Q = torch.randn((nb, nb), dtype=torch.float64, device="cuda"), torch.randn((nb, nb), dtype=torch.float64, device="cuda")
    R_ = torch.empty_like(P)
    best = 1e9
    for i in range(6):
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        torch.matmul(P, Q.T, out=R_)
        b1.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, b0.elapsed_time(b1) * 1e-3)
    peak = 2.0 * nb**3 / best / 1e12
    del P, Q, R_

    traffic = None
    tf = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    cpu = None
    if world == 1 and not args.no_cpu:
        n0 = min(CPU_SAMPLE_N, N)
        t_N, sample = cpu_extrapolated(3, 1, n0, N)
        cpu = {"value": 1.0 / t_N, "unit": "evals/s", "cores": cpu_cores(), "kind": "port", "sample": sample}

    step_tflops = alg_flops(N, D) / (ms * 1e-3) / 1e12
    extra = None
    if world == 1 and not args.no_extra:
        del x, y, model
        torch.cuda.empty_cache()
        try:
            extra = other_configs(peak, float(os.environ.get("GPXB_BENCH_EXTRA_SCALE", "1.0")))
        except Exception as e:  # noqa: BLE001
            extra = [{"error": f"{type(e).__name__}: {e}"[:300]}]
    line = {
        "metric": "LML+grad evals/s (exact GPR, Matern-5/2, N=32768, D=16, FP64)",
        "value": world * 1e3 / ms, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"exact GPR Matern-5/2 N={N} D={D} LML+grad (BASELINE configs[1])",
                   "parallelism": "replicas only" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (A is %.1f GB)" % (8.0 * N * N / 1e9),
                   "lml": float(res[0]), "grad": [float(res[1]), float(res[2])]},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": "gemm_f64_kernel (DMMA GEMM / SYRK, all layouts): %d launches of one step with the "
                               "look-ahead serialised (no overlapping launches), summed launch durations %.1f ms; "
                               "algorithmic flops counted exactly per tile" % (gemm_n, gemm_ms),
                     "in_timed_region": {"launches": reg_n, "summed_launch_ms": reg_ms, "region_ms": ms * K,
                                         "tflops_over_summed_ms": reg_flops / (reg_ms * 1e-3) / 1e12,
                                         "note": "look-ahead launches overlap here: overlapped time is counted twice"},
                     "avg_launch_ms": gemm_ms / max(gemm_n, 1), "alg_flops_per_launch": gemm_flops / max(gemm_n, 1),
                     "probe": {"shape": f"lower SYRK M=N={m_t} K={k_t}", "tflops": probe},
                     "peak_source": f"cuBLAS Dgemm {nb}^3 measured in this run",
                     "step_tflops": step_tflops, "step_frac": step_tflops / peak},
        "cpu_baseline": cpu,
        "e2e": {"value": world / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(x_np.nbytes + y_np.nbytes),
                "d2h_bytes_per_step": 24},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "other_configs": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpxb", choices=["gpxb", "reference"])
    ap.add_argument("--rows", "--n", dest="n", type=int, default=0,
                    help="override N (debugging only; the bench line names it).  Under torchrun use --rows: "
                         "torchrun's own parser rejects the abbreviation-ambiguous --n")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the other_configs section (SGPR / RPCholesky / Lanczos at N=1e6, about 20 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
