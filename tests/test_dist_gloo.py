"""CPU, world_size 2 over gloo: the N>1 host logic (batch sharding, the single gradient all-reduce, max-over-ranks)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from x2i_b200 import dist as xdist
    r, lr, w = xdist.init("gloo")
    assert (r, w) == (rank, world)
    # data-parallel distillation semantics: every rank's loss is sum/bsz_local, grads averaged over ranks
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    data = torch.randn(4, 6)
    lo, hi = xdist.shard_range(4, rank, world)
    loss = lin(data[lo:hi]).pow(2).sum() / (hi - lo)
    loss.backward()
    n = xdist.allreduce_mean_grads_(lin.parameters())
    t = xdist.max_over_ranks(1.0 + rank, torch.device("cpu"))
    xdist.barrier()
    q.put((rank, n, lin.weight.grad.clone(), lin.bias.grad.clone(), t, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_grad_allreduce_equals_big_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    data = torch.randn(4, 6)
    (lin(data).pow(2).sum() / 4).backward()  # single-process, whole batch
    for rank, n, gw, gb, t, span in res:
        assert n == 21
        torch.testing.assert_close(gw, lin.weight.grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(gb, lin.bias.grad, rtol=1e-5, atol=1e-6)
        assert t == 2.0
    assert [r[5] for r in res] == [(0, 2), (2, 4)]
