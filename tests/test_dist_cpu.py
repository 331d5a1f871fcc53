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


# ------------------------------------------------------------------------------------------------ TrainStep host logic over gloo
def _trainstep_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import copy
    import torch.distributed as dist
    from cotnet_b200 import dist as cdist, trainer
    cdist.init_from_env(backend="gloo")

    class HostTrainStep(trainer.TrainStep):
        """The product's bucket / chunk / hook / all-reduce logic on CPU tensors: only the two kernel launches are replaced by
        torch ops (test infrastructure; the product class refuses CPU models)."""
        _host_logic_only = True

        def _launch_gather(self, tab, ent, bucket, grads):
            for (_, off, n, _), g in zip(ent, grads):
                bucket[off:off + n] = g.reshape(-1).to(bucket.dtype)

        def optimizer_step(self):
            for P, M, G, E, h in ((self.P_big, self.M_big, self.G_big, self.E_big, self.hyper),
                                  (self.P_small, self.M_small, self.G_small, self.E_small, self.hyper_small)):
                lr, mu, wd, dec, gs = [float(v) for v in h]
                g = G.float() * gs + wd * P
                M.mul_(mu).add_(g)
                P.sub_(lr * (g + mu * M))
                if E is not None:
                    E.mul_(dec).add_((1 - dec) * P)

    torch.manual_seed(7)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, bias=True), torch.nn.ReLU(), torch.nn.Conv2d(8, 8, 3), torch.nn.Flatten(),
                              torch.nn.Linear(8 * 4 * 4, 16), torch.nn.ReLU(), torch.nn.Linear(16, 5))
    ref = copy.deepcopy(net)
    lr, mu, wd = 0.1, 0.9, 1e-2
    ts = HostTrainStep(net, lr=lr, momentum=mu, weight_decay=wd, nesterov=True, ema_decay=0.9, amp_dtype=None, weights="fp32",
                       comm_chunks=2, overlap=True)
    assert len(ts.plan["chunks"]) == 2 and ts.overlap and ts.world == 2
    decay = [p for n, p in ref.named_parameters() if not (p.dim() == 1 or n.endswith(".bias"))]
    no_decay = [p for n, p in ref.named_parameters() if (p.dim() == 1 or n.endswith(".bias"))]
    opt = torch.optim.SGD([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": wd}], lr=lr, momentum=mu, nesterov=True)
    flushed_in_hooks = []
    orig_flush = ts._flush_chunk
    in_backward = {"v": False}

    def spy(c):
        flushed_in_hooks.append((c, in_backward["v"]))
        orig_flush(c)
    ts._flush_chunk = spy
    orig_fb = ts.forward_backward

    worst = 0.0
    for step in range(2):
        g = torch.Generator().manual_seed(100 * step + rank)
        x = torch.randn(4, 3, 8, 8, generator=g)
        y = torch.randint(0, 5, (4,), generator=g)
        in_backward["v"] = True
        ts.step_eager(x, y)
        in_backward["v"] = False
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        for p in ref.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        opt.step()
        ms = ts.master_state()
        for n, p in ref.named_parameters():
            worst = max(worst, (ms[n] - p.detach()).abs().max().item())
    q.put((rank, worst, [c for c, _ in flushed_in_hooks], float(ts.P_big.sum()), float(ts.hyper[4])))
    dist.destroy_process_group()


def test_trainstep_bucket_logic_two_rank_gloo():
    """TrainStep's flat buckets, chunk planning, post-accumulate hooks (last chunk flushed first, during backward),
    chunked all-reduce(mean) and the update arithmetic, 2 ranks over gloo, against optim.SGD(nesterov, add_weight_decay) on
    all-reduced gradients."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_trainstep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, worst, order, psum, gscale in res:
        assert worst <= 1e-5, worst
        assert order == [1, 0, 1, 0], order                 # backward produces the LAST chunk's gradients first, every step
        assert abs(gscale - 0.5) < 1e-12                    # gloo: SUM all-reduce, 1/world applied by the update
    assert abs(res[0][3] - res[1][3]) <= 1e-4               # replicas stay in lock-step


def test_plan_flat_partition():
    from cotnet_b200 import backbone, trainer
    m = backbone.cotnet50()
    plan = trainer.plan_flat(list(m.named_parameters()), 3)
    big, small = plan["big"], plan["small"]
    assert len(big) + len(small) == len(list(m.parameters()))
    assert all(p.dim() >= 2 for _, p, _ in big) and all(p.dim() == 1 for _, p, _ in small)      # optim_factory.add_weight_decay split
    offs = [o for _, _, o in big]
    assert offs == sorted(offs) and all(o % trainer.ALIGN == 0 for o in offs)
    assert all(o2 - o1 >= p.numel() for (_, p, o1), o2 in zip(big, offs[1:] + [plan["n_big"]]))    # slots do not overlap
    ch = plan["chunks"]
    assert len(ch) == 3 and ch[0][0] == 0 and ch[-1][1] == plan["n_big"]
    assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))                                           # contiguous cover
    assert sorted(i for c in ch for i in c[2]) == list(range(len(big)))
    sizes = [hi - lo for lo, hi, _ in ch]
    assert max(sizes) <= 1.6 * (plan["n_big"] / 3)                                                  # roughly balanced
