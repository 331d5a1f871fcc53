"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, max-over-ranks timing, gradient averaging."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from cotnet_b200 import dist as cdist
    r, lr, w = cdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = cdist.shard_range(513, r, w)
    mx = cdist.max_over_ranks(10.0 + rank, "cpu")
    sm = cdist.sum_over_ranks(hi - lo, "cpu")
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    local = lin.weight.grad.clone()
    cdist.allreduce_grads_(list(lin.parameters()), w)
    cdist.barrier()
    # flat gradient bucket: grads are views, one all-reduce(mean), parameters broadcast from rank 0
    torch.manual_seed(100 + rank)                     # different initial weights per rank on purpose
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Flatten(), torch.nn.Linear(4, 2))
    net = net.to(memory_format=torch.channels_last)
    cdist.broadcast_module_(net, 0)
    w_after_bcast = net[0].weight.detach().clone()
    fg = cdist.FlatGrads(net.parameters())
    fg.zero_()
    net(torch.full((2, 3, 3, 3), float(rank + 1))).sum().backward()
    assert fg.attached()
    flat_local = fg.flat.clone()
    fg.all_reduce_mean_()
    views_ok = all(torch.equal(p.grad.reshape(-1), p.grad.contiguous().reshape(-1)) for p in net.parameters())
    q.put((rank, lo, hi, mx, sm, local.tolist(), lin.weight.grad.tolist(), w_after_bcast.reshape(-1).tolist(),
           flat_local.tolist(), fg.flat.tolist(), net[0].weight.grad.reshape(-1).tolist(), views_ok))   # plain lists: no fd passing
    dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=60) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, lo0, hi0, mx0, sm0, l0, g0, wb0, fl0, fa0, cg0, ok0), (_, lo1, hi1, mx1, sm1, l1, g1, wb1, fl1, fa1, cg1, ok1) = res
    assert wb0 == wb1                                         # broadcast_module_: replicas start identical
    fl0, fl1, fa0, fa1 = [torch.tensor(t) for t in (fl0, fl1, fa0, fa1)]
    assert torch.allclose(fa0, fa1) and torch.allclose(fa0, (fl0 + fl1) / 2)   # FlatGrads: one all-reduce(mean)
    assert ok0 and ok1 and fl0.abs().sum() > 0 and cg0 == cg1   # p.grad views see the reduced values
    assert (lo0, hi0, lo1, hi1) == (0, 257, 257, 513)        # shards tile the range, remainder on rank 0
    assert mx0 == mx1 == 11.0                                 # slowest rank defines the time
    assert sm0 == sm1 == 513.0
    g0, g1, l0, l1 = [torch.tensor(t) for t in (g0, g1, l0, l1)]
    assert torch.allclose(g0, g1) and torch.allclose(g0, (l0 + l1) / 2)   # averaged gradients agree on all ranks


def test_shard_range_edges():
    from cotnet_b200.dist import shard_range
    assert [shard_range(5, r, 8) for r in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]
    assert shard_range(0, 0, 2) == (0, 0)
