"""Scaling of the row-sharded matrix-free paths (BASELINE configs[3] part 2 and configs[4]).
Run directly (1 GPU) or under torchrun (one rank per GPU).  Times are device times, max over
ranks.  One JSON line per measurement on rank 0.

    [torchrun ...] tools/bench_sharded.py [N_slq] [m] [N_sgpr] [M]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from gpx_b200 import ops  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    N1, m, N2, M = (a + [2_000_000, 3, 1_000_000, 4096][len(a):])[:4]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rank = 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank = dist.get_rank()
        ops.comm_init_from_torch_distributed()

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
        return float(t.item()), out

    key = np.array([0, 2023], dtype=np.uint32)
    g = torch.Generator(device="cuda").manual_seed(4)
    P, D = 30, 16
    if N1 > 0:
        x = torch.randn((N1, D), dtype=torch.float64, device="cuda", generator=g)
        U = ops.rademacher_probes(key, N1, P, True)
        ops.lanczos("se", x[:4096].contiguous(), 4.0, 0.25, U[:4096].contiguous(), 2) if world == 1 else None
        t, (al, be, fl) = timed(lambda: ops.lanczos("se", x, 4.0, 0.25 + 1e-10, U, m))
        flops = m * float(N1) ** 2 * (2 * D + 4 + 2 * P)
        if rank == 0:
            print(json.dumps(dict(what="lanczos_block_steps", n_gpus=world, N=N1, D=D, P=P, m=m, s=t,
                                  s_per_step=t / m, tflops_total=flops / t / 1e12,
                                  tflops_per_gpu=flops / t / 1e12 / world,
                                  alpha00=float(al[0, 0].cpu()), scaling="strong")), flush=True)
        del x, U
    if N2 > 0:
        g2 = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randn((N2, 32), dtype=torch.float64, device="cuda", generator=g2)
        y = torch.sin(x.sum(1, keepdim=True))
        idx = torch.randperm(N2, device="cuda", generator=g2)[:M].contiguous()
        xl = ops.gather_rows(x, idx)
        ops.sgpr_lml("se", x[:65536].contiguous(), y[:65536].contiguous(), xl, 32**0.5, 0.1) if world == 1 else None
        t, out = timed(lambda: ops.sgpr_lml("se", x, y, xl, 32**0.5, 0.1))
        fl = 2.0 * N2 * M * 32 + 2.0 * M * M * N2 + 2.0 * M**3 / 3
        if rank == 0:
            print(json.dumps(dict(what="sgpr_lml_woodbury", n_gpus=world, N=N2, M=M, D=32, s=t,
                                  tflops_total=fl / t / 1e12, tflops_per_gpu=fl / t / 1e12 / world,
                                  lml=float(out.cpu()[0]), scaling="strong")), flush=True)
    if N2 > 0 and os.environ.get("BENCH_RPCHOL", "1") == "1":
        t, (_, piv) = timed(lambda: ops.rpcholesky("se", x, M, 32**0.5, key, True, want_factor=False))
        alg = 8.0 * N2 * M * (M - 1) / 2 + 5 * 8.0 * N2 * M
        if rank == 0:
            print(json.dumps(dict(what="rpcholesky", n_gpus=world, N=N2, M=M, D=32, s=t, gbs_total=alg / t / 1e9,
                                  hbm_frac_per_gpu=alg / t / 1e9 / world / 6554.2,
                                  pivots_head=piv[:6].cpu().tolist(), scaling="strong")), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
