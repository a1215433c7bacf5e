"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded paths: every rank derives
the same 1024-aligned row partition, the ranges tile [0, N), and the 128-byte communicator id
travels intact through the broadcast used by ops.comm_init_from_torch_distributed."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpx_b200 import ops

    out = {}
    for N in (1, 255, 256, 1000, 4096, 1_000_000, 2_000_000):
        b, e = ops.shard_range(N, world, rank)
        t = torch.tensor([b, e], dtype=torch.int64)
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, t)
        out[N] = [tuple(int(v) for v in a) for a in allr]
        # every rank computes every other rank's range identically
        for r in range(world):
            assert ops.shard_range(N, world, r) == out[N][r]
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    got = ops.broadcast_bytes(payload, 128, 0)
    assert got == bytes(range(128))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_row_partition_and_id_broadcast_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = res[0][1]
    for N, rr in ranges.items():
        assert rr[0][0] == 0 and rr[-1][1] == N
        for a, b in zip(rr[:-1], rr[1:]):
            assert a[1] == b[0]  # contiguous, no gap, no overlap
        for b, e in rr[:-1]:
            assert (b % 1024 == 0 and e % 1024 == 0) or e == N
    assert res[0][1] == res[1][1]


def test_row_partition_properties_for_every_world_size():
    """gpxb_shard_range (host-only): for 1..8 ranks the ranges tile [0, N) in order, every interior boundary is a
    multiple of 1024 (the RPCholesky block-sum granularity) and the loads differ by at most one block."""
    import numpy as np

    from gpx_b200 import ops

    rng = np.random.default_rng(0)
    sizes = [1, 2, 1023, 1024, 1025, 4096, 10_000, 1_000_000, 2_000_000] + [int(v) for v in rng.integers(1, 3_000_000, 40)]
    for N in sizes:
        for world in range(1, 9):
            rr = [ops.shard_range(N, world, r) for r in range(world)]
            assert rr[0][0] == 0 and rr[-1][1] == N
            for (b0, e0), (b1, e1) in zip(rr[:-1], rr[1:]):
                assert e0 == b1 and b0 <= e0
            for b, e in rr:
                assert b % 1024 == 0 or b == N
            blocks = [-(-(e - b) // 1024) for b, e in rr]
            assert max(blocks) - min(blocks) <= 1
