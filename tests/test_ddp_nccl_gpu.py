"""N-rank NCCL data-parallel training == 1-rank training on the same global batch.

Every rank holds the same initial parameters and takes a contiguous shard of the global batch; after each backward the flat
gradient is all-reduced (NCCL over NVLink) and the fused Adam applies 1/world.  With full-length sequences every rank has the same
number of valid targets, so DDP's "mean of per-rank means" equals the single-rank mean over the whole batch and the two runs
must agree to fp32 reduction-order noise.  Skipped with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

V, L, D, H, NB, BG, STEPS = 300, 40, 64, 2, 2, 16, 3


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _batches():
    from genrec_b200.data import synthetic_batch
    return [synthetic_batch(BG, L, V, seed=100 + i, full_length=True) for i in range(STEPS)]


def _train(dev, rank, world, group=None):
    from genrec_b200.hstu import HSTU
    from genrec_b200.optim import FlatAdam
    torch.manual_seed(0)
    m = HSTU(V, L, D, H, NB, dropout=0.0).to(dev).train()
    opt = FlatAdam(m, lr=1e-3, betas=(0.9, 0.98), process_group=group)
    import torch.distributed as dist
    per = BG // world
    losses, grads = [], []
    for ids, ts, tg in _batches():
        sl = slice(rank * per, (rank + 1) * per)
        _, loss = m(ids[sl].to(dev), ts[sl].to(dev), tg[sl].to(dev))
        loss.backward()
        g = opt.grad.detach().clone()                      # this rank's gradient; averaged over ranks below for the comparison
        if world > 1:
            dist.all_reduce(g)
        grads.append((g / world).cpu())
        opt.step()                                         # peer-memory one-pass step, or NCCL all-reduce + Adam (GRB_DP=nccl)
        losses.append(loss.detach())
    torch.cuda.synchronize(dev)
    return opt.flat.detach().cpu().clone(), torch.stack(losses).cpu(), grads, opt.dp_mode


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    flat, losses, grads, mode = _train(dev, rank, world)
    gathered = [torch.zeros_like(losses).to(dev) for _ in range(world)]
    dist.all_gather(gathered, losses.to(dev))
    if rank == 0:
        out["flat"] = flat
        out["grads"] = grads
        out["mode"] = mode
        out["loss"] = torch.stack([g.cpu() for g in gathered]).mean(0)
    ref = flat.to(dev).clone()
    dist.broadcast(ref, 0)
    out[f"same{rank}"] = bool(torch.equal(ref.cpu(), flat))          # replicas stay bit-identical
    dist.destroy_process_group()


@pytest.mark.parametrize("dp", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_rank_nccl_step_equals_single_rank_step(world, dp, monkeypatch):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.setenv("GRB_DP", dp)          # inherited by the spawned ranks
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    flat1, loss1, grads1, _ = _train(torch.device("cuda:0"), 0, 1)
    assert all(out[f"same{r}"] for r in range(world))
    print("data-parallel step:", out["mode"])
    assert out["mode"] == "nccl-allreduce" if dp == "nccl" else out["mode"].startswith(("peer-", "nccl-"))
    torch.testing.assert_close(out["loss"], loss1, rtol=1e-4, atol=1e-5)
    n = min(out["flat"].numel(), flat1.numel())               # the flat buffers are padded to a multiple of 64 * world at the END
    g0, g1 = out["grads"][0][:n], grads1[0][:n]               # first step: identical parameters, only the reduction order differs
    torch.testing.assert_close(g0, g1, rtol=1e-3, atol=1e-4 * g1.abs().max().item())   # (bf16 noise is 4e-3 of the max)
    # Adam turns a gradient into +-lr whatever its size, so an element whose gradient is reduction-order noise may move the other
    # way; everything else must agree
    far = ((out["flat"][:n] - flat1[:n]).abs() > 1e-4).float().mean().item()
    assert far < 1e-3, far
