"""Multi-GPU sharding of a batch of comparisons.

Every comparison touches only its own inputs and nonces plus the shared read-only keys
(``initiator.py:228-564`` / ``keyholder.py:181-287`` are pure per-comparison functions), so the batch
shards by comparison index: rank ``j`` of ``G`` owns the contiguous block ``shard_range(B, j, G)``, keys
are replicated, and there is NO collective on the data path.  ``torch.distributed`` (NCCL on GPUs, gloo
in the CPU tests) only carries the barrier, the max-over-ranks of the timings and -- when a caller wants
the whole result on every rank -- one final all-gather of the result limbs.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Half-open index range of ``rank``'s block: blocks are contiguous, ordered by rank, cover
    ``[0, total)`` exactly, and differ in size by at most one (the first ``total % world`` ranks get the extra)."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError("bad shard specification")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_range(total, r, world) for r in range(world))]


def shard_arrays(arrays: dict, rank: int, world: int) -> dict:
    """Slice every ``[B, ...]`` array of a workload dict down to this rank's block (views, no copies)."""
    total = next(iter(arrays.values())).shape[0]
    lo, hi = shard_range(total, rank, world)
    return {k: v[lo:hi] for k, v in arrays.items()}


def all_gather_rows(local: np.ndarray, total: int, group=None) -> np.ndarray:
    """Concatenate the per-rank result blocks (``[b_j, ...]`` uint32 / uint8 arrays) in rank order on every rank.
    Uses ``torch.distributed`` with whatever backend the process group has (gloo: CPU tensors; NCCL: staged
    through the current CUDA device).  Blocks may be ragged (``total`` not divisible by the world size)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, its shard has {sizes[rank]}")
    tail = local.shape[1:]
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    as_i = np.int32 if local.dtype == np.uint32 else local.dtype
    pad = max(sizes)
    buf = torch.zeros((pad,) + tail, dtype=torch.from_numpy(np.zeros(1, dtype=as_i)).dtype, device=device)
    if sizes[rank]:
        buf[:sizes[rank]] = torch.from_numpy(np.ascontiguousarray(local).view(as_i)).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = [o[:n].cpu().numpy().view(local.dtype) for o, n in zip(out, sizes)]
    return np.concatenate(parts, axis=0)
