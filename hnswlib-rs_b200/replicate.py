"""Multi-GPU plumbing: replicate a frozen index over torch.distributed and shard queries.

The reference is single-process (rayon over independent queries, /root/reference/src/hnsw.rs:1612-1635); on
GPUs the same independence means: replicate the read-only graph, give every rank a contiguous shard of the
queries, gather the answers.  No collective sits on the traversal's critical path.
torch.distributed is plumbing only (NCCL broadcast / all-gather on tensors that alias the library's device
arrays); with the gloo backend and host pointers the same code runs on CPU for the protocol tests.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def as_byte_tensor(ptr, nbytes, device):
    """uint8 tensor aliasing `nbytes` at raw pointer `ptr` (device = 'cpu' or 'cuda:N')."""
    if str(device).startswith("cuda"):
        return torch.as_tensor(_DevMem(ptr, nbytes), device=device)
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8))


def shard_bounds(n, rank, world):
    """contiguous shard [lo, hi) of n items for `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_index(h, src, device, group=None):
    """Replicate the index of rank `src` into the (empty) handles of all other ranks.

    `h` needs blob_header() -> u64[16], blob_alloc(header), blobs() -> [(ptr, nbytes)], blob_commit()
    (Hnsw in hnsw.py; the C ABI calls are hnsw_b200_blob_*).  Returns the number of bytes broadcast."""
    rank = dist.get_rank(group)
    hdr = torch.from_numpy(h.blob_header().astype(np.int64)) if rank == src else torch.zeros(16, dtype=torch.int64)
    hdr = hdr.to(device)
    dist.broadcast(hdr, src, group=group)
    if rank != src:
        h.blob_alloc(hdr.cpu().numpy().astype(np.uint64))
    total = 0
    for ptr, nb in h.blobs():
        if nb == 0:
            continue
        t = as_byte_tensor(ptr, nb, device)
        dist.broadcast(t, src, group=group)   # ncclBroadcast over NVLink/NVSwitch when device is cuda
        total += nb
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()
    if rank != src:
        h.blob_commit()
    return total


def all_gather_answers(local, world, group=None):
    """local: tensor [nq_local, ...] (same shape on every rank) -> [world, nq_local, ...] on every rank, rank order."""
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)   # ncclAllGather (concatenated along dim 0, rank order)
    return out.view((world,) + tuple(local.shape))
