#!/usr/bin/env python
"""bench.py -- headline benchmark of the GP hot path on B200.

Workload (BASELINE.json configs[1]): exact GPR, Matern-5/2, N=32768, D=16, FP64;
one "step" = one LML + gradient evaluation (Gram build, blocked Cholesky, triangular
solves, SPD inverse, fused dK/d-lengthscale contraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu] [--no-extra]

The JSON line also carries `other_configs`: the row-sharded matrix-free paths of BASELINE configs[3], [4] --
RPCholesky landmark selection and the SGPR Woodbury LML at N = 1e6 (D = 32, M = 4096), block Lanczos steps and PCG
iterations of the iterative LML at N = 2e6 (D = 16, 30 probes), and one complete iterative LML (PCG + SLQ) at a
size that finishes in seconds -- each with its roofline fraction and a per-phase split (mat-vec / all-gather /
dot products / all-reduce / replicated Cholesky), measured after the headline at EVERY --gpus N.

N > 1 (torchrun): the dense Cholesky path does not shard ("replicas only") --
every rank evaluates an independent replica (its own hyper-parameter point, as the
reference's random restarts do) and `value` is the whole-job evaluations per second.  The `other_configs`
entries ARE sharded: rows of X are split over the ranks (strong scaling: the problem size is fixed), NCCL
all-reduce / all-gather between them; times are device times, max over ranks.

`--impl reference` times the CPU restatement of the reference (oracle/ref_numpy.py; the
reference itself needs JAX, which this image does not have) at the FULL headline size: one measured
evaluation (about 2-4 minutes), never an extrapolation.
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
CPU_SAMPLE_N = 8192  # bounded sample of the GPU arm's cpu_baseline leg (a few seconds per evaluation)
GOLDEN = os.path.join(ROOT, "tests", "golden", "headline_N32768.json")


def synth(N, D, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, D))
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    return x, y


def alg_flops(N, D):
    return float(N) ** 3 + 2.0 * float(N) ** 2 * D


# ----------------------------------------------------------------------------- CPU arm
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_mem_available_gb():
    """Memory this process may still take: MemAvailable, capped by the cgroup limit when there is one."""
    avail = 1e9
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = float(ln.split()[1]) / 1e6
    except OSError:
        pass
    for f_max, f_cur in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                         ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(f_max).read().strip()
            if lim != "max" and float(lim) < 1e15:
                avail = min(avail, (float(lim) - float(open(f_cur).read().strip())) / 1e9)
        except (OSError, ValueError):
            pass
    return avail


def cpu_measure(n, steps, warmup):
    """(seconds per LML+grad, per-phase seconds of the last step, (lml, grads)) of the CPU restatement of the
    reference at n rows of the headline workload, on every host core the process may use (torchrun exports
    OMP_NUM_THREADS=1; the BLAS pool is widened explicitly).  Measured, never extrapolated."""
    from threadpoolctl import threadpool_limits

    from oracle import ref_numpy as R  # the checker, timed as the CPU baseline

    x, y = synth(n, CFG["D"], CFG["seed"])
    nthr = cpu_cores()
    tm, res = {}, None
    with threadpool_limits(limits=nthr):
        for _ in range(warmup):
            R.lml_grad_dense_lowmem(CFG["kind"], x, y, CFG["ell"], CFG["sigma"], threads=nthr)
        t0 = time.perf_counter()
        for _ in range(steps):
            res = R.lml_grad_dense_lowmem(CFG["kind"], x, y, CFG["ell"], CFG["sigma"], threads=nthr, timings=tm)
        t = (time.perf_counter() - t0) / steps
    return t, tm, res


def golden_headline():
    try:
        return json.load(open(GOLDEN))
    except Exception:
        return None


def run_reference(args):
    """The reference arm: the oracle's LML + gradient at the headline size, measured.  One evaluation of
    N = 32768 is 2-4 minutes of host time, so exactly ONE step is timed whatever --steps says (no warm-up: the
    time is dominated by LAPACK dpotrf / dpotri); "steps" and "warmup" report what was run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = args.n or CFG["N"]
    need_gb = 2.4 * 8.0 * N * N / 1e9  # two N x N matrices + row-block temporaries
    note = ""
    n_run = N
    if need_gb > host_mem_available_gb():
        n_run = N // 2
        note = (f" (host memory {host_mem_available_gb():.0f} GB < {need_gb:.0f} GB needed at N={N}: measured at "
                f"N={n_run} and scaled by (N/{n_run})^3 for dpotrf+dpotri, (N/{n_run})^2 for the rest)")
    steps = max(1, args.steps) if n_run <= 4096 else 1
    warm = min(args.warmup, 1) if n_run <= 4096 else 0
    t, tm, res = cpu_measure(n_run, steps, warm)
    if n_run != N:
        r = N / n_run
        cubic = tm.get("potrf", 0.0) + tm.get("potri", 0.0)
        t = cubic * r**3 + (t - cubic) * r**2
    value = 1.0 / t
    sample = (f"oracle/ref_numpy.py lml_grad_dense_lowmem (NumPy/OpenBLAS restatement of GPX _lml_dense + analytic "
              f"gradient; JAX absent) MEASURED at N={n_run}, D={CFG['D']}: {steps} timed evaluation(s), "
              + ", ".join(f"{k} {v:.1f} s" for k, v in tm.items()) + note)
    gold = golden_headline() if N == CFG["N"] else None
    parity = None
    if gold and n_run == N:
        ref = np.array([gold["lml"]] + list(gold["grad"]))
        parity = float(np.max(np.abs(np.array(res) - ref) / np.abs(ref)))
    line = {
        "impl": "reference", "metric": "LML+grad evals/s (exact GPR, Matern-5/2, N=32768, D=16, FP64)",
        "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"exact GPR Matern-5/2 N={N} D={CFG['D']} LML+grad (BASELINE configs[1])",
                   "parallelism": "host cores (CPU restatement of the reference)", "measured_N": n_run,
                   "extrapolated": n_run != N,
                   "lml": float(res[0]), "grad": [float(res[1]), float(res[2])],
                   "max_rel_diff_vs_committed_golden": parity},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cpu_cores(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    # LAPACK worked in place on multi-GB buffers; skip the interpreter's teardown of them
    sys.stdout.flush()
    os._exit(0)


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


N1_REFERENCE = os.path.join(ROOT, "profiles", "r2_other_configs_n1.json")


def other_configs(dgemm_tflops, world, rank, scale=1.0):
    """The SGPR / Lanczos half of BASELINE.json's metric line (configs[3], configs[4]), row-sharded over the
    `world` ranks this run was launched with (gpx_b200.ops shards rows inside each call once the library's NCCL
    communicator is up; world == 1 runs the same code unsharded).  STRONG scaling: the problem size is fixed.
    Called by every rank (the calls are collective); device-timed with CUDA events on torch's current stream (the
    stream the library launches on) between barriers, MAX over ranks; inputs are created on the device with the same
    seed on every rank and are far larger than L2.  Each entry carries the per-phase split the library records
    (ops.phase_collect: summed event-bracketed durations of this rank's phases, MAX over ranks per phase;
    collective phases include waiting for the slowest rank).  Reported next to the headline, never instead of it;
    every entry is isolated (a failure is reported as {"error": ...}).  `scale` shrinks the sizes for smoke
    tests (GPXB_BENCH_EXTRA_SCALE)."""
    import torch
    import torch.distributed as dist

    from gpx_b200 import ops
    from gpx_b200.models._iterative import slq_logdet

    key = np.array([0, 2023], dtype=np.uint32)
    N = max(4096, int(1_000_000 * scale))
    N2 = max(4096, int(2_000_000 * scale))
    M = 4096 if scale >= 1.0 else max(64, int(4096 * scale * 4) // 64 * 64)
    hbm, hbm_src = hbm_peak_gbs()
    out = []
    try:
        n1ref = json.load(open(N1_REFERENCE)) if scale >= 1.0 else {}
    except Exception:
        n1ref = {}

    def allmax(v):
        t = torch.tensor(v, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().tolist()

    def timed(fn, phases=True):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if phases:
            ops.phase_collect()
            ops.phase_enable(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = fn()
        b.record()
        torch.cuda.synchronize()
        ph = {}
        if phases:
            ops.phase_enable(False)
            mine = ops.phase_collect()
            vals = allmax([mine.get(k, (0.0, 0))[0] for k in ops.PHASES])
            vmin = [-v for v in allmax([-mine.get(k, (0.0, 0))[0] for k in ops.PHASES])]
            ph = {k: {"ms_max_over_ranks": round(v, 3), "ms_min_over_ranks": round(u, 3)}
                  for k, v, u in zip(ops.PHASES, vals, vmin) if v > 0.0}
        return allmax([a.elapsed_time(b) * 1e-3])[0], res, ph

    def data(n, d, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        x = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
        y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
        return x, y

    only = [v for v in os.environ.get("GPXB_BENCH_ONLY", "").split(",") if v]  # builder's aid: run these entries only

    def entry(name, fn):
        if only and fn.__name__ not in only:
            return
        try:
            d = dict(config=name, n_gpus=world, scaling="strong", **fn())
            ref = n1ref.get(d.get("id", ""))
            if isinstance(ref, dict):  # {"s": seconds, "N": rows}: only the same problem size is comparable
                ref = ref.get("s") if ref.get("N") == d.get("N") else None
            if ref and world > 1 and d.get("s"):
                d["n1_reference_s"] = ref
                d["speedup_vs_committed_n1"] = ref / d["s"]
                d["efficiency_vs_committed_n1"] = ref / d["s"] / world
                d["n1_reference_source"] = "profiles/r2_other_configs_n1.json (this bench at --gpus 1, another box)"
            out.append(d)
        except Exception as e:  # noqa: BLE001 -- reported, never fatal for the headline
            out.append(dict(config=name, n_gpus=world, error=f"{type(e).__name__}: {e}"[:300]))

    COLLECTIVE = ("allgather", "allreduce")

    def non_compute(t, ph):
        """Wall time not covered by this rank's compute brackets: collectives, waiting for the slowest rank and
        launch gaps (robust against what an event pair around a collective can and cannot see)."""
        return t - sum(v["ms_max_over_ranks"] for k, v in ph.items() if k not in COLLECTIVE) * 1e-3

    # NCCL connects lazily: the first all-reduce / all-gather / send-recv on a communicator sets up its channels
    # (hundreds of ms).  Run every collective pattern once at toy size, untimed, before anything is measured.
    xw, yw = data(16384, 16, 1)
    ops.rpcholesky("se", xw, 16, 4.0, key, True, want_factor=False)
    ops.lanczos("se", xw, 4.0, 0.25, ops.rademacher_probes(key, 16384, 4, True), 2)
    ops.gpr_fit_iter("se", xw, yw, 4.0, 0.5, 8, key, maxiter=2)
    torch.cuda.synchronize()
    del xw, yw

    x32, y32 = data(N, 32, 3)
    ell32 = 32**0.5
    state = {}

    def rpchol():
        call = lambda: ops.rpcholesky("se", x32, M, ell32, key, True, want_factor=False)  # noqa: E731
        t, (_, piv), ph = timed(call)
        extra = {}
        if non_compute(t, ph) > 0.03 * t:
            # The call's first act is a stream-ordered allocation of its 8 N M-byte factor workspace; on a box whose
            # driver maps those 32 GB slowly that is up to seconds inside the first timed call (seen: 0.16, 0.6 and 2.8 s;
            # the kernels' own 9.9 s unchanged).  The library's pool keeps the memory, so a second call measures the
            # landmark loop itself; both times are reported.
            t2, (_, piv2), ph2 = timed(call)
            extra = dict(first_call_s=t, first_call_non_compute_s=non_compute(t, ph),
                         retimed="the first call spent >3 % of its time outside the kernels (workspace allocation); "
                                 "s is the second call")
            if t2 < t:
                t, piv, ph = t2, piv2, ph2
        state["piv"], state["t_rp"] = piv, t
        alg = 8.0 * N * M * (M - 1) / 2 + 5 * 8.0 * N * M
        return dict(id="rpcholesky", what="RPCholesky landmark loop (pivots only), rows sharded", N=N, D=32, M=M, s=t,
                    **extra,
                    bound="hbm", alg_bytes=alg, achieved_gbs_total=alg / t / 1e9, achieved_gbs_per_gpu=alg / t / 1e9 / world,
                    peak_gbs=hbm, peak_source=hbm_src, frac=alg / t / 1e9 / world / hbm,
                    pivots_head=[int(v) for v in piv[:6].cpu().tolist()],
                    pivots_checksum=int(piv.sum().cpu()), phases=ph, non_compute_s=non_compute(t, ph),
                    phases_note="phase events are sampled on every 16th pivot and scaled by 16",
                    frac_note="denominator = measured COPY bandwidth (read + write streams); this pass is 99.9 % "
                              "reads, and a read-only stream can exceed the copy rate, so frac may pass 1")

    def sgpr():
        piv = state.get("piv")
        if piv is None:
            _, piv = ops.rpcholesky("se", x32[:65536].contiguous(), M, ell32, key, True, want_factor=False)
        xl = ops.gather_rows(x32, piv)
        ops.sgpr_lml("se", x32, y32, xl, ell32, 0.1)  # warm-up (allocator pools)
        t, res, ph = timed(lambda: ops.sgpr_lml("se", x32, y32, xl, ell32, 0.1))
        fl = 2.0 * N * M * 32 + 2.0 * M * M * N + 2.0 * M**3 / 3
        d = dict(id="sgpr_lml", what="SGPR Woodbury LML (fixed landmarks), rows sharded", N=N, D=32, M=M, s=t,
                 evals_per_s=1.0 / t, bound="tensor", alg_flops=fl, achieved_tflops_total=fl / t / 1e12,
                 achieved_tflops_per_gpu=fl / t / 1e12 / world, peak_tflops=dgemm_tflops,
                 frac=fl / t / 1e12 / world / dgemm_tflops, lml=float(res.cpu()[0]), phases=ph,
                 phases_note="sgpr_chunk brackets run on three concurrent streams: their sum exceeds the wall time")
        if "t_rp" in state:  # SGPR_RPChol re-draws the landmarks at every LML evaluation (_sgpr_rpchol.py:193-200)
            d["evals_per_s_with_landmark_selection"] = 1.0 / (t + state["t_rp"])
        return d

    P, D16 = 30, 16

    def lanczos_steps():
        m = 2
        x16, _ = data(N2, D16, 4)
        U = ops.rademacher_probes(key, N2, P, True)
        t, (al, be, fl_), ph = timed(lambda: ops.lanczos("se", x16, 4.0, 0.25 + 1e-10, U, m))
        fl = m * float(N2) ** 2 * (2 * D16 + 4 + 2 * P)
        mv = ph.get("matvec", {}).get("ms_max_over_ranks", 0.0) * 1e-3
        return dict(id="lanczos_step", what=f"block Lanczos, {m} steps (mat-vec + full MGS) of the 50 of each 30-probe SLQ "
                    "log-det, rows sharded", N=N2, D=D16, P=P, m=m, s=t / m, s_total=t, bound="tensor", alg_flops_per_step=fl / m,
                    achieved_tflops_total=fl / t / 1e12, achieved_tflops_per_gpu=fl / t / 1e12 / world,
                    peak_tflops=dgemm_tflops, frac=fl / t / 1e12 / world / dgemm_tflops,
                    matvec_only_tflops_per_gpu=(fl / mv / 1e12 / world) if mv > 0 else None,
                    alpha00=float(al[0, 0].cpu()), phases=ph, non_compute_s=non_compute(t, ph))

    def pcg_iters():
        it = 3
        x16, y16 = data(N2, D16, 4)
        t, (c, mu, iters), ph = timed(lambda: ops.gpr_fit_iter("se", x16, y16, 4.0, 0.5, 100, key, maxiter=it))
        mv = ph.get("matvec", {}).get("ms_max_over_ranks", 0.0) * 1e-3
        fl = float(N2) ** 2 * (2 * D16 + 4 + 2) * iters
        return dict(id="pcg_iter", what=f"PCG fit, {it} iterations (P = 1 mat-vec + RPCholesky preconditioner with 100 "
                    "pivots; the time includes building the preconditioner), rows sharded", N=N2, D=D16, iters=int(iters),
                    s=t, s_matvec_per_iter=mv / max(iters, 1), bound="tensor", alg_flops_per_iter=fl / max(iters, 1),
                    matvec_tflops_per_gpu=(fl / mv / 1e12 / world) if mv > 0 else None, peak_tflops=dgemm_tflops,
                    frac=(fl / mv / 1e12 / world / dgemm_tflops) if mv > 0 else None, phases=ph,
                    non_compute_s=non_compute(t, ph))

    def lml_iter_full():
        Nf = int(os.environ.get("GPXB_BENCH_ITER_N", "0")) or max(4096, int(262144 * scale))
        npiv = int(os.environ.get("GPXB_BENCH_ITER_PIVOTS", "1024" if scale >= 1.0 else "64"))
        x16, y16 = data(Nf, D16, 4)
        ell, sigma, m = 4.0, 0.5, 50

        def run(xx, yy, mm):
            U = ops.rademacher_probes(key, xx.shape[0], P, True)
            return ops.gpr_lml_iter("se", xx, yy, ell, sigma, npiv, key, U, mm)

        run(x16[:16384].contiguous(), y16[:16384].contiguous(), 4)  # warm-up at toy size (allocator pools, NCCL channels)
        t, (c, mu, iters, al, be, flag), ph = timed(lambda: run(x16, y16, m))
        logdet = slq_logdet(al.cpu().numpy(), be.cpu().numpy(), Nf)
        quad = float(((y16 - mu) * c).sum().cpu())
        lml = (-0.5 * quad - 0.5 * logdet - 0.5 * Nf * np.log(2 * np.pi)) / Nf
        fl = float(Nf) ** 2 * ((2 * D16 + 4 + 2 * P) * m + (2 * D16 + 4 + 2) * iters)
        return dict(id="lml_iter_full", what=f"COMPLETE iterative LML (_lml_iter, ONE gpxb_gpr_lml_iter call: probes, "
                    f"{npiv}-pivot RPCholesky preconditioner, PCG fit to tol 1e-5 whose first {m} iterations ride in the "
                    f"SLQ block mat-vecs, SLQ log-det with 30 probes x {m} Lanczos steps), rows sharded; host eigh of the "
                    "30 tridiagonals excluded", N=Nf, D=D16, P=P, m=m, n_pivots=npiv, pcg_iters=int(iters), s=t,
                    evals_per_s=1.0 / t, bound="tensor", alg_flops=fl, achieved_tflops_per_gpu=fl / t / 1e12 / world,
                    peak_tflops=dgemm_tflops, frac=fl / t / 1e12 / world / dgemm_tflops, lml=lml,
                    breakdown=int(flag.cpu()[0]), phases=ph, non_compute_s=non_compute(t, ph))

    def cfg0():
        """configs[0]: exact GPR, SquaredExponential, N = 2000, D = 8: LML+grad and predict (every rank runs its own
        replica: the dense path does not shard).  Latency bound: 16 sequential 128-column panel steps."""
        n0, d0 = 2000, 8
        xa, ya = data(n0, d0, 11)
        xs, _ = data(512, d0, 12)
        f = lambda: ops.gpr_lml_grad("se", xa, ya, 2.0, 0.1)  # noqa: E731
        for _ in range(3):
            f()
        best = min(timed(f, phases=False)[0] for _ in range(5))
        c, mu = ops.gpr_fit("se", xa, ya, 2.0, 0.1)
        gp = lambda: ops.gpr_predict("se", xa, xs, c, mu, 2.0, 0.1)  # noqa: E731
        gp()
        tp = min(timed(gp, phases=False)[0] for _ in range(5))
        fl = float(n0) ** 3 + 2.0 * n0 * n0 * d0
        return dict(id="cfg0_lml_grad", what="exact GPR SE N=2000 D=8: LML+grad (one fused call) and predict mean of 512 "
                    "points; replicas, best of 5", N=n0, D=d0, s=best, ms=best * 1e3, evals_per_s=1.0 / best,
                    predict_ms=tp * 1e3, bound="tensor", alg_flops=fl, achieved_tflops_per_gpu=fl / best / 1e12,
                    peak_tflops=dgemm_tflops, frac=fl / best / 1e12 / dgemm_tflops,
                    note="latency bound: the chain leaf -> head solve -> head update of 16 panel steps, not flops")

    def cfg2_iter():
        """configs[2]: GPR on derivative values, 30 atoms (nf = nv = 90), N = 2000 configurations = a 180 000-row
        Hessian system (259 GB as a dense square: over one GPU's HBM as such; the dense packed-lower path takes
        63 s, profiles/r1_configs.jsonl).  Here: the matrix-free path (PCG fit with the rpcholesky_derivs
        preconditioner + SLQ log-det on the Hessian operator), replicas."""
        from gpx_b200.models import _iterative as it

        nc, nf = max(16, int(2000 * min(scale, 1.0))), 90
        g = torch.Generator(device="cuda").manual_seed(21)
        xa = torch.randn((nc, nf), dtype=torch.float64, device="cuda", generator=g)
        Ja = torch.eye(nf, dtype=torch.float64, device="cuda").repeat(nc, 1, 1).contiguous()
        ya = torch.randn((nc * nf,), dtype=torch.float64, device="cuda", generator=g)
        npv = 200 if scale >= 1.0 else 16
        t, lml, _ = timed(lambda: it.lml_derivs_iter("se", xa, Ja, ya, 9.0, 0.3, 30, 50, key, npv, key), phases=False)
        return dict(id="cfg2_lml_derivs_iter", what="derivative-trained GPR, Hessian kernel, matrix-free LML (PCG + SLQ "
                    "30 probes x 50 steps on the Hessian operator; host eigh included); replicas", N=nc, nf=nf, nv=nf,
                    rows=nc * nf, n_pivots=npv, s=t, evals_per_s=1.0 / t, lml=float(lml))

    if world == 1:  # single-GPU configurations (nothing to shard): reported at --gpus 1 only
        entry("configs[0] exact GPR SE N=2000 D=8: LML+grad, predict", cfg0)
        entry("configs[2] GPR on derivatives, 30 atoms (D=90), N=2000 (180k-row Hessian system): iterative LML",
              cfg2_iter)
    entry(f"configs[3] SGPR + RPCholesky, N={N} D=32 M={M}: landmark selection", rpchol)
    entry(f"configs[3] SGPR + RPCholesky, N={N} D=32 M={M}: LML", sgpr)
    del x32, y32
    torch.cuda.empty_cache()
    entry(f"configs[4] Lanczos/SLQ, SE kernel, D=16, 30 probes, N={N2}: Lanczos steps", lanczos_steps)
    entry(f"configs[4] iterative LML, SE kernel, D=16, N={N2}: PCG iterations", pcg_iters)
    entry("configs[4] iterative LML end to end (LML evals/s)", lml_iter_full)
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # ONE JSON line on stdout: NCCL writes its version banner to fd 1, so everything but the final line goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
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

    # ---- parity at the headline size: the committed oracle result (tools/make_golden_headline.py) ----
    parity = None
    gold = golden_headline()
    if rank == 0 and gold and N == gold["config"]["N"] and D == gold["config"]["D"]:
        ref = np.array([gold["lml"]] + list(gold["grad"]))
        err = float(np.max(np.abs(res - ref) / np.abs(ref)))
        parity = {"max_rel_diff_vs_oracle": err, "tolerance": 1e-8, "ok": bool(err <= 1e-8),
                  "oracle": "tests/golden/headline_N32768.json (oracle/ref_numpy.py run once at N=32768)"}
        assert err <= 1e-8, f"headline LML/gradient differ from the committed oracle result: {res} vs {ref}"

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

    peak, probe, achieved, nb, m_t, k_t = None, None, None, None, None, None
    if rank == 0:
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

    # ---- the row-sharded matrix-free paths (all ranks: the calls are collective) ----
    extra = None
    if not args.no_extra:
        del x, y, model, xh, yh
        torch.cuda.empty_cache()
        pk = torch.tensor([peak or 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.broadcast(pk, 0)
            ops.comm_init_from_torch_distributed()
        try:
            extra = other_configs(float(pk.item()), world, rank, float(os.environ.get("GPXB_BENCH_EXTRA_SCALE", "1.0")))
        except Exception as e:  # noqa: BLE001
            extra = [{"error": f"{type(e).__name__}: {e}"[:300]}]

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    traffic = None
    tf = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    cpu = None
    if world == 1 and not args.no_cpu:
        # bounded sample of the same workload, MEASURED (no extrapolation): the oracle at CPU_SAMPLE_N rows; the
        # full-size figure is what `--impl reference` measures (and tools/make_golden_headline.py recorded)
        n0 = min(CPU_SAMPLE_N, N)
        t_s, tm, _ = cpu_measure(n0, 2, 1)
        r = N / n0
        cubic = tm.get("potrf", 0.0) + tm.get("potri", 0.0)
        t_N = cubic * r**3 + (t_s - cubic) * r**2  # LAPACK dpotrf / dpotri are O(N^3), everything else O(N^2)
        g = gold or {}
        cpu = {"value": 1.0 / t_N, "unit": "evals/s", "cores": cpu_cores(), "kind": "port",
               "sample": (f"oracle/ref_numpy.py lml_grad_dense_lowmem (NumPy/OpenBLAS restatement of GPX _lml_dense + "
                          f"analytic gradient; JAX absent) measured on the bounded sample N={n0}, D={D}: {t_s:.2f} s per "
                          "LML+grad (" + ", ".join(f"{k} {v:.2f} s" for k, v in tm.items()) + f"), scaled to N={N}: "
                          f"dpotrf+dpotri x{r:.0f}^3, the rest x{r:.0f}^2 = {t_N:.0f} s.  The same oracle MEASURED at the "
                          f"full size: bench.py --impl reference; the committed golden run took "
                          f"{g.get('wall_seconds', float('nan')):.0f} s on {g.get('cores')} cores "
                          "(tests/golden/headline_N32768.json)"),
               "sample_N": n0, "sample_seconds": t_s, "scaled_to_full_size": n0 != N,
               "full_size_measured": {"seconds": g.get("wall_seconds"), "cores": g.get("cores"),
                                      "evals_per_s": (1.0 / g["wall_seconds"]) if g.get("wall_seconds") else None,
                                      "source": "tests/golden/headline_N32768.json"}}

    step_tflops = alg_flops(N, D) / (ms * 1e-3) / 1e12
    line = {
        "metric": "LML+grad evals/s (exact GPR, Matern-5/2, N=32768, D=16, FP64)",
        "value": world * 1e3 / ms, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"exact GPR Matern-5/2 N={N} D={D} LML+grad (BASELINE configs[1])",
                   "parallelism": "replicas only" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (A is %.1f GB)" % (8.0 * N * N / 1e9),
                   "lml": float(res[0]), "grad": [float(res[1]), float(res[2])], "parity": parity},
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
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
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
                    help="skip the other_configs section (the row-sharded SGPR / RPCholesky / Lanczos / PCG paths, about "
                         "90 s on one GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
