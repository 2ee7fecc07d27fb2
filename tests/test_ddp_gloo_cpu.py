"""world_size-2 gloo test (CPU) of the data-parallel plumbing: flat buffers, ONE all-reduce, 1/world scaling, batch sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genrec_b200.hstu import HSTU
    from genrec_b200.optim import FlatBuffers, allreduce_gradients
    from genrec_b200.data import shard_batch
    torch.manual_seed(0)                      # identical init on every rank
    m = HSTU(30, 16, 64, 2, 1)
    keys = [k for k, _ in m.named_parameters()]
    fb = FlatBuffers(m)
    ok = all(p.data_ptr() >= fb.flat.data_ptr() and p.grad.data_ptr() >= fb.grad.data_ptr() for p in m.parameters())
    ok &= all(p.data_ptr() % 256 == fb.flat.data_ptr() % 256 for p in m.parameters())
    # rank-specific gradients, written through the per-parameter views
    for i, p in enumerate(m.parameters()):
        p.grad.fill_(float((rank + 1) * (i + 1)))
    scale = allreduce_gradients(fb)
    got = [float(p.grad.flatten()[0]) * scale for p in m.parameters()]
    want = [sum((r + 1) * (i + 1) for r in range(world)) / world for i in range(len(keys))]
    ok &= got == want and scale == 1.0 / world
    # identical parameters on both ranks after the same update
    with torch.no_grad():
        fb.flat.add_(fb.grad, alpha=-0.1 * scale)
    t = fb.flat.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok &= torch.equal(t, fb.flat)
    # batch sharding: disjoint, covering, contiguous
    ids = torch.arange(10 * 4).view(10, 4)
    mine = shard_batch({"input_ids": ids}, rank, world)["input_ids"]
    gathered = [torch.zeros(5, 4, dtype=ids.dtype) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok &= torch.equal(torch.cat(gathered), ids)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_allreduce_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
