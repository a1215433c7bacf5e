"""Feasibility probe for the peer-memory exchange (one process per GPU, launched with torchrun): can a rank map
another rank's device buffer through CUDA IPC and write to it over NVLink?  Prints one JSON line per rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_probe.py
"""
import ctypes as C
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = C.CDLL("libcudart.so.12")
    out = {"rank": rank, "can_access_peer": [torch.cuda.can_device_access_peer(local, d) for d in range(world) if d != local]}
    buf = C.c_void_p()
    assert rt.cudaMalloc(C.byref(buf), 1 << 20) == 0
    rt.cudaMemset(buf, 0, 1 << 20)
    handle = (C.c_char * 64)()
    out["ipc_get"] = rt.cudaIpcGetMemHandle(handle, buf)
    h = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device="cuda")
    allh = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(allh, h)
    peer = (rank + 1) % world
    ph = (C.c_char * 64).from_buffer_copy(bytes(allh[peer].cpu().tolist()))
    pp = C.c_void_p()
    out["ipc_open"] = rt.cudaIpcOpenMemHandle(C.byref(pp), ph, 1)
    if out["ipc_open"] == 0:
        # write my rank + 1 into the peer's buffer, then read my own buffer after a barrier
        src = torch.full((1024,), float(rank + 1), dtype=torch.float64, device="cuda")
        out["memcpy_to_peer"] = rt.cudaMemcpy(pp, C.c_void_p(src.data_ptr()), 8192, 3)
        torch.cuda.synchronize()
        dist.barrier()
        mine = torch.empty(1024, dtype=torch.float64, device="cuda")
        rt.cudaMemcpy(C.c_void_p(mine.data_ptr()), buf, 8192, 3)
        torch.cuda.synchronize()
        out["received_from_prev_rank"] = float(mine[0].cpu())
        # bandwidth of a large peer copy
        big = C.c_void_p()
        n = 256 << 20
        rt.cudaMalloc(C.byref(big), n)
    print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
