"""Multi-GPU plumbing of the path: persons are independent, so a batch of crops is split into
contiguous per-rank shards (full weight replica per GPU) and the ONLY exchange is one
all-gather of the decoded (B_local, K, 7) fp32 records (replaces
mmengine's pickled ``collect_results``).  Works with any ``torch.distributed`` backend: NCCL on
the GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``n`` persons: the first ``n % world`` ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank records (B_local, K, 7) into (n_total, K, 7) on every rank, in person
    order.  Equal shards use a single ``all_gather_into_tensor``; ragged shards are padded to
    the largest shard so it is still ONE collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} records, its shard of {n_total} is {sizes[rank]}")
    cap = max(sizes)
    send = local.contiguous()
    if send.shape[0] != cap:
        send = torch.cat([send, send.new_zeros((cap - send.shape[0],) + tuple(send.shape[1:]))], 0)
    out = send.new_empty((world * cap,) + tuple(send.shape[1:]))
    dist.all_gather_into_tensor(out, send, group=group)
    if all(s == cap for s in sizes):
        return out
    return torch.cat([out[r * cap: r * cap + sizes[r]] for r in range(world)], 0)
