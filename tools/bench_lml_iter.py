"""BASELINE configs[4] end to end: the iterative LML (PCG fit with the RPCholesky preconditioner + SLQ
log-determinant, 30 Rademacher probes x 50 Lanczos steps) of an SE-kernel GPR on N x 16 synthetic data,
row-sharded over the ranks it is launched with (1 GPU: run directly; N GPUs: under torchrun).  Device
time, max over ranks; one JSON line on rank 0.

    [torchrun ...] tools/bench_lml_iter.py [N] [n_pivots] [P] [m]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402
from gpx_b200.models._iterative import slq_logdet  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    N, npiv, P, m = (a + [1_000_000, 100, 30, 50][len(a):])[:4]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rank = 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank = dist.get_rank()
        ops.comm_init_from_torch_distributed()
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((N, 16), dtype=torch.float64, device="cuda", generator=g)
    y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn((N, 1), dtype=torch.float64, device="cuda", generator=g)
    ell, sigma = 4.0, 0.5
    key = np.array([0, 2023], dtype=np.uint32)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.cpu()), out

    t_pcg, (c, mu, iters) = timed(lambda: ops.gpr_fit_iter("se", x, y, ell, sigma, npiv, key))
    U = ops.rademacher_probes(key, N, P, True)
    t_slq, (al, be, flag) = timed(lambda: ops.lanczos("se", x, ell, sigma**2 + 1e-10, U, m))
    logdet = slq_logdet(al.cpu().numpy(), be.cpu().numpy(), N)
    quad = float(((y - mu) * c).sum().cpu())
    lml = (-0.5 * quad - 0.5 * logdet - 0.5 * N * np.log(2 * np.pi)) / N
    if rank == 0:
        fl_slq = float(N) ** 2 * (2 * 16 + 4 + 2 * P) * m
        fl_pcg = float(N) ** 2 * (2 * 16 + 4 + 2) * iters
        print(json.dumps(dict(what="lml_iter_end_to_end", n_gpus=world, N=N, D=16, P=P, m=m, n_pivots=npiv,
                              pcg_iters=int(iters), s_pcg=t_pcg, s_slq=t_slq, s_total=t_pcg + t_slq,
                              evals_per_s=1.0 / (t_pcg + t_slq), tflops_slq=fl_slq / t_slq / 1e12,
                              tflops_pcg=fl_pcg / t_pcg / 1e12, lml=lml, breakdown=int(flag.cpu()[0]))), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
