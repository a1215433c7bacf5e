"""Diagnostic (torchrun, one rank per GPU): bandwidth of NCCL all-gather / send-recv through torch.distributed and of
a raw peer copy through a CUDA-IPC mapping, to tell an NCCL transport problem from a link problem."""
import ctypes as C
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 34 * 1_000_000  # doubles per rank (272 MB)
    mine = torch.randn(n, dtype=torch.float64, device="cuda")
    full = torch.empty(n * world, dtype=torch.float64, device="cuda")
    out = {"rank": rank}

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms = timed(lambda: dist.all_gather_into_tensor(full, mine))
    out["allgather_ms"] = ms
    out["allgather_recv_GBps"] = n * 8 * (world - 1) / ms / 1e6

    def sr():
        ops = [dist.P2POp(dist.isend, mine, (rank + 1) % world), dist.P2POp(dist.irecv, full[:n], (rank - 1) % world)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    ms = timed(sr)
    out["sendrecv_ms"] = ms
    out["sendrecv_GBps"] = n * 8 / ms / 1e6
    # raw IPC peer copy
    rt = C.CDLL("libcudart.so.12")
    handle = (C.c_char * 64)()
    base = C.c_void_p(full.untyped_storage().data_ptr())
    out["ipc_get"] = rt.cudaIpcGetMemHandle(handle, base)
    h = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device="cuda")
    allh = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(allh, h)
    peer = (rank + 1) % world
    ph = (C.c_char * 64).from_buffer_copy(bytes(allh[peer].cpu().tolist()))
    pp = C.c_void_p()
    out["ipc_open"] = rt.cudaIpcOpenMemHandle(C.byref(pp), ph, 1)
    if out["ipc_open"] == 0:
        def pc():
            rt.cudaMemcpyAsync(pp, C.c_void_p(mine.data_ptr()), C.c_size_t(n * 8), 3, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        ms = timed(pc)
        out["peer_copy_ms"] = ms
        out["peer_copy_GBps"] = n * 8 / ms / 1e6
    print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
