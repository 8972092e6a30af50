"""pp_allgather behind the C ABI on a real NCCL communicator (world size 1 on the single test GPU: the collective
degenerates to a copy, but the whole chain - ProcessGroupNCCL._comm_ptr() -> pp_allgather -> ncclAllGather resolved from
the loaded libnccl, side stream, double-buffered send / receive - is the one the multi-GPU bench runs)."""
import os

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nccl_world1():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29581")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        yield dev
    finally:
        dist.destroy_process_group()


def test_record_gatherer_world1(nccl_world1):
    from probpose_code_b200 import synth
    from probpose_code_b200.engine import Engine
    from probpose_code_b200.sharding import RecordGatherer
    dev = nccl_world1
    eng = Engine(precision="fp16x3", max_batch=4).load_state_dict(synth.make_state_dict(seed=0))
    g = RecordGatherer(4, 1, dev)
    crops = [synth.make_crops(4, seed=60 + i).to(dev) for i in range(5)]
    want = [eng.infer(c).clone() for c in crops]
    outs = []
    for i, c in enumerate(crops):
        eng.infer(c, out=g.send_buffer(i))  # the decode kernel writes the send buffer itself
        outs.append((i, g.gather(i)))
        if i >= 1:  # the previous step's gather has been overlapped with this step: read it back now
            g.wait()
            torch.cuda.synchronize()
    g.wait()
    torch.cuda.synchronize()
    # the two newest receive buffers hold steps 3 and 4
    assert torch.equal(outs[4][1], want[4]) and torch.equal(outs[3][1], want[3])
    assert g.gather_ms() is not None and g.gather_ms() >= 0.0


def test_pp_allgather_argument_checks(nccl_world1):
    import ctypes as C
    from probpose_code_b200 import _lib
    lib = _lib.lib()
    x = torch.zeros(8, device=nccl_world1)
    assert lib.pp_allgather(None, x.data_ptr(), x.data_ptr(), 8, None) == -1  # NULL communicator
    assert b"communicator" in lib.pp_last_error()
    from probpose_code_b200.sharding import nccl_comm_ptr
    comm = nccl_comm_ptr(nccl_world1)
    assert lib.pp_allgather(comm, x.data_ptr(), x.data_ptr(), 0, None) == 0  # empty shard: nothing to do
    assert lib.pp_allgather(comm, None, x.data_ptr(), 8, None) == -1
