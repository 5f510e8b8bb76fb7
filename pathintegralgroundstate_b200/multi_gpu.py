"""Multi-GPU plumbing: one process per GPU, independent Markov chains sharded
over ranks, ONE collective per Monte-Carlo block -- a sum all-reduce of the
block accumulator vector (12 energy sums, 24 counters, g(r), S(k), n(r)
histograms; ~3 KB) over NCCL/NVLink.  Chains never exchange configurations
(SURVEY.md section 8(e)), so there is no other data-path traffic.

torch.distributed is used for the rendezvous and the collective only.
"""
from __future__ import annotations

import os

import numpy as np


def shard_chains(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of chain ids: (first, count) for this rank."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def chain_seed(seed: int, first_chain: int) -> int:
    """MT19937 replay only: rank r's local chain i is seeded with sgrnd(seed + first + i).  Prefer
    PigsCuda(..., seed=seed, chain_offset=first): the library then mixes the GLOBAL chain index into both the Philox
    stream and the MT seed, so chain c of the job draws the same numbers whatever the number of ranks."""
    return int(seed) + int(first_chain)


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8", data=(int(ptr), False), version=2)


def init_process_group(backend: str | None = None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the env)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def allreduce_block(sim, device_tensor=None):
    """Sum the chain-summed block vector of `sim` (a PigsCuda) over all ranks, in
    place in the library's device buffer, and return the unpacked global block
    result.  With one rank this is just a device->host read."""
    import torch
    import torch.distributed as dist

    ptr, n = sim.block_vector()
    sim.sync()
    t = device_tensor
    if t is None:
        t = torch.as_tensor(_CudaArray(ptr, n), device=f"cuda:{sim.p.device}")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    vec = t.cpu().numpy()
    return sim.unpack_block_vector(vec)


def allreduce_vector_host(vec: np.ndarray) -> np.ndarray:
    """The same reduction for a host vector (gloo; used by the CPU tests of the
    N>1 path and by drivers that already hold the vector on the host)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64).copy())
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
