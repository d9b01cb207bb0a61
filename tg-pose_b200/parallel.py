"""Multi-GPU plumbing: one process per GPU, the batch of clouds sharded, no data-path collective.

Every op of the path is per-cloud (kNN, gathers, convs, the ORL mean is over the points of ONE cloud,
gcn3d.py:216; chamfer is per batch element), so inference shards the batch contiguously and needs no
collective; the reference itself is single-device (trainer/RL_TDA.py:27).  Training adds exactly one
collective, a gradient all-reduce (sum then / world) over NCCL (gloo in the CPU tests).
Pool_layer draws its permutation from the CPU generator (gcn3d.py:242): every rank must seed it identically
before each forward so that a sharded run equals the single-GPU run bit for bit.
"""
import os

import torch
import torch.distributed as dist


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_bounds(n_items, rank, world):
    """contiguous split of n_items over `world` ranks; the first n_items % world ranks get one more."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world):
    lo, hi = shard_bounds(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


def seed_for_forward(seed):
    """same CPU-RNG state on every rank before a forward (Pool_layer's randperm)."""
    torch.manual_seed(seed)


def allreduce_gradients(params, world=None, bucket_bytes=32 << 20):
    """average .grad over all ranks: flatten into ~32 MB buckets, one all_reduce each (NVLink/NVSwitch:
    size buckets for launch latency, not link count)."""
    if not dist.is_initialized():
        return 0
    world = world or dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    n_buckets, i = 0, 0
    while i < len(grads):
        bucket, size = [], 0
        while i < len(grads) and (not bucket or size + grads[i].numel() * 4 <= bucket_bytes):
            bucket.append(grads[i])
            size += grads[i].numel() * 4
            i += 1
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_buckets += 1
    return n_buckets


def allreduce_flat(flat, world=None, bucket_bytes=32 << 20):
    """average a flat gradient arena (ranger.Ranger.flat_grads) over all ranks in place: one all_reduce per ~32 MB
    slice, no flatten / unflatten copies."""
    if not dist.is_initialized():
        return 0
    world = world or dist.get_world_size()
    step = max(1, bucket_bytes // flat.element_size())
    n_buckets = 0
    for lo in range(0, flat.numel(), step):
        piece = flat[lo:lo + step]
        dist.all_reduce(piece, op=dist.ReduceOp.SUM)
        piece.div_(world)
        n_buckets += 1
    return n_buckets


@torch.no_grad()
def sharded_inference(net, points, cat_id, seed=7, gather=True):
    """run `net` on this rank's slice of (points, cat_id); optionally all_gather the small pose outputs."""
    rank, _, world = world_info()
    pts, cat = shard_batch([points, cat_id], rank, world)
    seed_for_forward(seed)
    out = net(pts.contiguous(), cat.contiguous())
    if not gather or world == 1 or not dist.is_initialized():
        return out
    res = {}
    for k, v in out.items():
        if v is None or v.dim() == 0:
            continue
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v.contiguous())     # requires equal shard sizes
        res[k] = torch.cat(parts, 0)
    return res
