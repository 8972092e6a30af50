"""World-size-2 gloo run of the multi-GPU plumbing (shard split + the single all-gather)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_is_a_partition():
    from probpose_code_b200.sharding import shard_range

    for n in (0, 1, 7, 64, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    from probpose_code_b200.sharding import gather_records, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_total * 17 * 7, dtype=torch.float32).reshape(n_total, 17, 7)
    lo, hi = shard_range(n_total, rank, world)
    out = gather_records(full[lo:hi].clone(), n_total)
    q.put((rank, bool(torch.equal(out, full))))
    try:
        gather_records(full[:0], n_total)
    except ValueError:
        q.put((rank, True))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_gather_records_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500 + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = [q.get(timeout=5) for _ in range(4)]
    assert all(ok for _, ok in res)
