"""world_size-2 gloo test (CPU) of the host side of the multi-GPU path in bench.py: the 128-byte communicator id travels
from rank 0 to every rank, query shards are contiguous, disjoint and cover the batch, and the "every replica answers the
probe batch like rank 0" check trips when one rank differs.  (The NCCL side lives in the library, multi.cu, and is
exercised on the GPU box by tests/test_gpu_multi.py.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # communicator id: created by rank 0, identical everywhere afterwards
        uid = torch.from_numpy(np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, np.uint8))
        dist.broadcast(uid, 0)
        ok = bool(np.array_equal(uid.numpy(), np.arange(128, dtype=np.uint8)))
        # shards
        for n in (0, 1, 7, 10000, 1000001):
            lo, hi = bench.shard_bounds(n, rank, world)
            t = torch.tensor([lo, hi], dtype=torch.int64)
            g = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(g, t)
            b = [x.tolist() for x in g]
            ok &= b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            ok &= max(x[1] - x[0] for x in b) - min(x[1] - x[0] for x in b) <= 1
        # replica check: equal answers pass, one differing rank trips every rank
        for bad in (False, True):
            ids = torch.arange(50, dtype=torch.int64)
            if bad and rank == 1:
                ids[7] = 99
            ref = ids.clone()
            dist.broadcast(ref, 0)
            same = torch.tensor([int(torch.equal(ids, ref))])
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            ok &= int(same.item()) == (0 if bad else 1)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_world2_host_protocol():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
