"""world_size-2 gloo test (CPU) of the multi-GPU host logic: index replication protocol (header -> alloc ->
per-array broadcast -> commit), query sharding, answer all-gather in rank order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeIndex:
    """host-memory stand-in with the blob protocol of hnswlib-rs_b200.hnsw.Hnsw"""

    def __init__(self, build):
        self.committed = False
        self.arrays = []
        if build:
            rng = np.random.default_rng(7)
            self.n, self.d = 1000, 24
            self.arrays = [rng.random((self.n, 32), dtype=np.float32), rng.integers(0, self.n, (self.n, 16)).astype(np.uint32),
                           np.zeros(0, np.uint32), np.arange(self.n, dtype=np.uint64)]

    def blob_header(self):
        h = np.zeros(16, np.uint64)
        h[0], h[1], h[2] = 0x68623230306e7377, self.n, self.d
        for i, a in enumerate(self.arrays):
            h[4 + i] = a.nbytes
        return h

    def blob_alloc(self, h):
        assert int(h[0]) == 0x68623230306e7377
        self.n, self.d = int(h[1]), int(h[2])
        self.arrays = [np.zeros(int(h[4 + i]), np.uint8) for i in range(4)]

    def blobs(self):
        return [(a.ctypes.data, a.nbytes) for a in self.arrays]

    def blob_commit(self):
        self.committed = True


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import importlib
    rep = importlib.import_module("hnswlib-rs_b200.replicate")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = FakeIndex(build=(rank == 0))
    nbytes = rep.broadcast_index(idx, 0, "cpu")
    ref = FakeIndex(build=True)
    same = all(np.array_equal(np.frombuffer(a.tobytes(), np.uint8), np.frombuffer(b.tobytes(), np.uint8))
               for a, b in zip(idx.arrays, ref.arrays))
    # query sharding + answers gathered in rank order
    nq = 10
    lo, hi = rep.shard_bounds(2 * nq, rank, world)
    local = torch.arange(lo, hi, dtype=torch.int64).reshape(nq, 1) * 10
    allv = rep.all_gather_answers(local, world)
    q.put((rank, nbytes, same, idx.committed or rank == 0, allv.reshape(-1).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds(pkg):
    import importlib
    rep = importlib.import_module("hnswlib-rs_b200.replicate")
    for n in (0, 1, 7, 10000, 1000003):
        for w in (1, 2, 3, 8):
            b = [rep.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_replication_protocol_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nbytes, same, committed, gathered in res:
        assert same and committed
        assert nbytes == 1000 * 32 * 4 + 1000 * 16 * 4 + 1000 * 8
        assert gathered == [i * 10 for i in range(20)]
