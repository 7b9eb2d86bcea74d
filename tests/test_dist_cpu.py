"""CPU suite: the N>1 plumbing (unit sharding, max-over-ranks timing) under gloo with world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semabs_b200 import dist as sd

    r, w = sd.rank_world()
    units = list(range(17))
    mine = sd.shard_units(units, r, w)
    step_ms = 10.0 + 5.0 * rank  # rank 1 is the slow one
    out.put((rank, mine, sd.max_over_ranks(step_ms), sd.sum_over_ranks(len(mine))))
    dist.destroy_process_group()


def test_sharding_and_timing_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    (r0, u0, t0, n0), (r1, u1, t1, n1) = res
    assert u0 + u1 == list(range(17)) and abs(len(u0) - len(u1)) <= 1  # every unit exactly once, balanced
    assert t0 == t1 == 15.0  # max over ranks
    assert n0 == n1 == 17.0


def test_shard_units_properties():
    from semabs_b200.dist import shard_units

    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            parts = [shard_units(list(range(n)), r, world) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semabs_b200 import train

    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2)),
          torch.nn.Parameter(torch.zeros(4))]
    ps[0].grad = torch.full((5, 3), float(rank + 1))
    ps[1].grad = torch.arange(7.0) * (rank + 1)
    ps[2].grad = None                                   # unused on every rank (e.g. VOOL's completion-net sampler): stays None
    ps[3].grad = torch.ones(4) if rank == 0 else None   # used on one rank only (a relation embedding): zeros from the other
    train.all_reduce_gradients(ps)
    out.put((rank, [None if p.grad is None else p.grad.clone() for p in ps]))
    dist.destroy_process_group()


def test_gradient_all_reduce_world2():
    """DDP-style gradient averaging (utils.py:255-258) incl. the unused-parameter cases of SURVEY.md §8e."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    for rank, g in res:
        assert torch.equal(g[0], torch.full((5, 3), 1.5))
        assert torch.equal(g[1], torch.arange(7.0) * 1.5)
        assert g[2] is None
        # used on one rank only: every rank gets the average, so every rank's LAMB updates it identically
        assert torch.equal(g[3], torch.full((4,), 0.5))


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semabs_b200 import train

    ps = [torch.nn.Parameter(torch.zeros(*s)) for s in ((4, 3), (6,), (2, 2, 2), (5,), (3,))]
    with train.GradientBuckets() as gb:
        assert train.GradientBuckets.active() is gb
        # what unet3d_bwd.UNetBackward.backward does: a dict that grows level by level, handed over after each level
        grads = {}
        grads[ps[0]] = torch.full((4, 3), float(rank + 1))
        grads[ps[1]] = torch.arange(6.0) * (rank + 1)
        gb.submit(grads, list(grads)[0:])
        grads[ps[2]] = torch.full((2, 2, 2), 10.0 * rank)
        gb.submit(grads, list(grads)[2:])
        gb.submit(grads, [])  # a level without parameters sends nothing
        gb.join(grads)
        for p, g in grads.items():  # autograd would do this with the node's return values
            p.grad = g
    assert train.GradientBuckets.active() is None
    ps[3].grad = torch.full((5,), float(rank))  # a parameter outside the UNet: the flat all-reduce takes it
    before = [ps[i].grad.clone() for i in range(3)]
    train.all_reduce_gradients(ps, skip=gb.reduced)
    assert all(torch.equal(a, ps[i].grad) for i, a in enumerate(before)), "bucketed gradients were reduced twice"
    out.put((rank, gb.n_buckets, len(gb.reduced), [None if p.grad is None else p.grad.clone() for p in ps]))
    dist.destroy_process_group()


def test_gradient_buckets_world2(monkeypatch):
    """Per-level buckets reduced during backward (train.GradientBuckets) + the flat all-reduce of the rest give what
    one flat all-reduce of everything gives (utils.py:255-258 semantics)."""
    from semabs_b200 import train

    with train.GradientBuckets() as off:  # no process group: a no-op context
        assert not off.enabled and train.GradientBuckets.active() is None
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    for rank, n_buckets, n_reduced, g in res:
        assert (n_buckets, n_reduced) == (2, 3)
        assert torch.equal(g[0], torch.full((4, 3), 1.5))
        assert torch.equal(g[1], torch.arange(6.0) * 1.5)
        assert torch.equal(g[2], torch.full((2, 2, 2), 5.0))
        assert torch.equal(g[3], torch.full((5,), 0.5))
        assert g[4] is None
