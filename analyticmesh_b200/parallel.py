"""Multi-GPU plumbing: one process per GPU.  The default sharded march (include/am_b200.h: am_set_shard_p2p)
exchanges polygons and winner masks with the engine's own kernels over NVLink peer memory; torch.distributed
is only used to swap the CUDA IPC handles at set-up (make_allgather).  The round-1 scheme (am_set_shard /
am_set_shard_nccl: all-reduce of the level's polygons) is kept as a cross-check.

Per BFS level every rank clips the states it owns; the level's polygon scratch is zero in every slot
a rank does not own, so an integer all-reduce(SUM) over the int32 view of the buffer is the exact
union (each 32-bit word is non-zero on at most one rank).  Everything else (frontier, visited set,
stitching) is replicated and deterministic, hence bit-identical on all ranks.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist


class _DevMem:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n_int32):
        self.__cuda_array_interface__ = {"shape": (int(n_int32),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2}


def make_allreduce(group=None, device=None):
    """Returns allreduce(ptr, n_int32) for cuam.set_shard.  `device` None/cuda: `ptr` is device memory
    and NCCL is used; device == 'cpu': `ptr` is host memory (gloo) -- used by the CPU tests."""
    is_cpu = device is not None and torch.device(device).type == "cpu"

    def allreduce(ptr, n, stream=None):
        if is_cpu:
            buf = (ctypes.c_int32 * n).from_address(ptr)
            t = torch.from_numpy(np.frombuffer(buf, dtype=np.int32))
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return
        t = torch.as_tensor(_DevMem(ptr, n), device=device if device is not None else "cuda")
        if stream:
            # enqueue the collective in order on the engine's stream: no host synchronisation
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            torch.cuda.current_stream().synchronize()

    return allreduce


def owner_of_seeds(n_seeds_unique, world):
    """Seed state i (after de-duplication, in id order) is owned by rank i % world; children inherit the
    owner of the parent that discovered them (csrc/frontier.cuh finalize_kernel)."""
    return np.arange(n_seeds_unique) % world


def broadcast_bytes(payload, group=None, device=None):
    """rank 0's `payload` (bytes, fixed length 128) on every rank, through torch.distributed."""
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if payload is not None:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def make_allgather(group=None, device=None):
    """Returns allgather(payload: bytes) -> list[bytes] over the ranks of `group` (rank order), for
    cuam.set_shard_p2p.  device None: cuda tensors with NCCL, cpu tensors with gloo."""
    def allgather(payload):
        dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
        mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, mine, group=group)
        return [bytes(p.cpu().numpy().tobytes()) for p in parts]

    return allgather
