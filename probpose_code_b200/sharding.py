"""Multi-GPU plumbing of the path: persons are independent, so a batch of crops is split into
contiguous per-rank shards (full weight replica per GPU) and the ONLY exchange is one
all-gather of the decoded (B_local, K, 7) fp32 records (replaces
mmengine's pickled ``collect_results``).  Works with any ``torch.distributed`` backend: NCCL on
the GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import Optional, Tuple

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


def nccl_comm_ptr(device: torch.device, group=None) -> int:
    """The ncclComm_t of torch.distributed's NCCL process group for `device` (``ProcessGroupNCCL._comm_ptr()``), for
    ``pp_allgather``.  The communicator must exist: call after ``init_process_group(..., device_id=device)`` or after
    a first collective."""
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    backend = pg._get_backend(torch.device(device))
    ptr = backend._comm_ptr()
    if not ptr:
        raise RuntimeError("the NCCL communicator of this process group is not initialised yet")
    return int(ptr)


class RecordGatherer:
    """The per-step all-gather of the decoded records, off the compute stream.

    Two send buffers (the decode kernel writes its records straight into ``send_buffer(i)``) and two receive buffers,
    allocated once; ``gather(i)`` enqueues ``pp_allgather`` (one ``ncclAllGather`` behind the C ABI) on a side stream
    that waits for step i's decode, so the collective of step i runs under the kernels of step i + 1; the compute
    stream only waits for a buffer's previous gather before the decode of step i + 2 overwrites it."""

    def __init__(self, batch_local: int, world: int, device, keypoints: int = 17, floats: int = 7, group=None):
        from ._lib import lib  # noqa: F401  (fails loudly if the library is missing)
        self.device = torch.device(device)
        self.world, self.batch = world, batch_local
        self.count = batch_local * keypoints * floats
        self.send = [torch.empty((batch_local, keypoints, floats), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.recv = [torch.empty((world * batch_local, keypoints, floats), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.stream = torch.cuda.Stream(self.device)
        self.decoded = [torch.cuda.Event() for _ in range(2)]
        self.t0 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.t1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.pending = [False, False]
        self._timed = [True, True]
        self.comm = nccl_comm_ptr(self.device, group)
        self._ms, self._n = 0.0, 0

    def send_buffer(self, i: int) -> torch.Tensor:
        j = i & 1
        if self.pending[j]:  # step i - 2's gather still reads this buffer
            torch.cuda.current_stream(self.device).wait_event(self.t1[j])
        return self.send[j]

    def gather(self, i: int) -> torch.Tensor:
        """Enqueue the all-gather of ``send_buffer(i)`` (written on the current stream); returns the receive buffer,
        valid after :meth:`wait` (or after the side stream reaches this point)."""
        from ._lib import check, lib
        j = i & 1
        self._harvest(j)
        self.decoded[j].record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.decoded[j])
            self.t0[j].record(self.stream)
            check(lib().pp_allgather(self.comm, self.send[j].data_ptr(), self.recv[j].data_ptr(), self.count,
                                     self.stream.cuda_stream), "pp_allgather")
            self.t1[j].record(self.stream)
        self.pending[j] = True
        self._timed[j] = False
        return self.recv[j]

    def _harvest(self, j: int) -> None:
        """Adds slot j's last gather to the statistics if it has completed (never waits: the host usually runs several
        steps ahead of the device, so most gathers are only counted by :meth:`gather_ms`' final look)."""
        if self.pending[j] and not self._timed[j] and self.t1[j].query():
            self._ms += self.t0[j].elapsed_time(self.t1[j])
            self._n += 1
            self._timed[j] = True

    def wait(self) -> None:
        """The current stream waits for every outstanding gather."""
        cur = torch.cuda.current_stream(self.device)
        for j in range(2):
            if self.pending[j]:
                cur.wait_event(self.t1[j])

    def reset_stats(self) -> None:
        """Forget the gather timings so far (the first collective of a communicator sets up its channels)."""
        self._ms, self._n = 0.0, 0

    def gather_ms(self) -> Optional[float]:
        """Mean device time of a gather (side-stream events) over the gathers that had completed when their slot was
        next looked at, plus - after waiting for the side stream - the outstanding ones."""
        self.stream.synchronize()
        for j in range(2):
            self._harvest(j)
        return self._ms / self._n if self._n else None
