"""world_size-2 gloo tests of the host-side N>1 logic (no GPU): shard bounds, identical Pool permutation on
every rank, gradient all-reduce."""
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


def _worker(rank, world, port, q, done):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tgpose_b200 import parallel
    lo, hi = parallel.shard_bounds(33, rank, world)
    parallel.seed_for_forward(7)
    perm = torch.randperm(1028)[:257]
    lin = torch.nn.Linear(4, 3)
    with torch.no_grad():
        lin.weight.fill_(1.0)
        lin.bias.fill_(0.0)
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    nb = parallel.allreduce_gradients(lin.parameters(), world)
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)          # a flat gradient arena (ranger.Ranger.flat_grads)
    nflat = parallel.allreduce_flat(flat, world, bucket_bytes=16)       # 4-element slices -> 3 collectives
    # overlapped all-reduce over a flat arena: three parameters whose gradients are views of one buffer, 2 slices
    # (16 bytes = 4 elements each would give 3; use 6-element slices -> 2), one parameter WITHOUT a gradient this step
    ps = [torch.nn.Parameter(torch.full((4,), 1.0)), torch.nn.Parameter(torch.full((3,), 2.0)), torch.nn.Parameter(torch.full((5,), 3.0))]
    arena = torch.zeros(12)
    offs = [0, 4, 7]
    for p_, o in zip(ps, offs):
        p_.grad = arena[o:o + p_.numel()].view(p_.shape)
    ov = parallel.OverlappedAllReduce(arena, ps, offs, bucket_bytes=24, world=world)
    ov.begin()
    ((ps[0] * (rank + 1)).sum() + (ps[1] * 10 * (rank + 1)).sum()).backward()       # ps[2] gets no gradient
    n_slices = ov.finish()
    ov_res = (arena.tolist(), n_slices, ov.launched_early, all(p_.grad.data_ptr() == arena[o:].data_ptr() for p_, o in zip(ps, offs)))
    # plain lists, not tensors: a tensor travels as a shared-memory handle that dies with its producer
    q.put((rank, lo, hi, perm[:8].tolist(), lin.weight.grad.tolist(), lin.bias.grad.tolist(), nb, flat.tolist(), nflat, ov_res))
    dist.destroy_process_group()
    done.wait(timeout=120)          # stay alive until the parent has drained the queue


def test_two_rank_sharding_and_grad_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    done = ctx.Event()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, done)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    done.set()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [tuple(torch.tensor(v) if isinstance(v, list) and i in (4, 5, 7) else v for i, v in enumerate(t)) for t in res]
    ov0, ov1 = res[0][9], res[1][9]
    res = [t[:9] for t in res]
    # averaged over the two ranks: d/dp0 = (1 + 2) / 2, d/dp1 = 10 * 1.5, p2 untouched (0); gradients stayed arena views
    assert ov0[0] == ov1[0] == [1.5] * 4 + [15.0] * 3 + [0.0] * 5
    assert ov0[1] == 2 and ov0[3] and ov1[3]
    # slice 0 = elements 0..5 (p0, head of p1) went out from the hooks during backward; slice 1 (tail of p1, p2) waits for
    # p2, which got no gradient, and is launched by finish()
    assert ov0[2] == ov1[2] == 1
    (r0, lo0, hi0, perm0, gw0, gb0, nb0, f0, nf0), (r1, lo1, hi1, perm1, gw1, gb1, nb1, f1, nf1) = res
    assert nf0 == nf1 == 3 and torch.equal(f0, f1) and torch.allclose(f0, torch.arange(10, dtype=torch.float32) * 1.5)
    assert (lo0, hi0, lo1, hi1) == (0, 17, 17, 33)           # contiguous, disjoint, covering
    assert perm0 == perm1                                     # same Pool permutation on every rank
    # rank r: d/dW sum(lin(x)) = 2*(r+1) per entry; average over ranks = 3
    assert torch.allclose(gw0, torch.full((3, 4), 3.0)) and torch.equal(gw0, gw1)
    assert torch.allclose(gb0, torch.full((3,), 2.0)) and torch.equal(gb0, gb1)
    assert nb0 == 1


def test_shard_bounds_cover_everything():
    from tgpose_b200.parallel import shard_bounds
    for n in (1, 7, 32, 8192):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
