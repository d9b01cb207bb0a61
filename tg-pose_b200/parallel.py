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


class OverlappedAllReduce:
    """DDP-style overlap of the gradient all-reduce with backward over a FLAT gradient arena (ranger.Ranger.flat_grads).

    The arena is cut into ~bucket_bytes slices.  Every parameter whose gradient view lies (partly) in a slice holds one
    count of it; a post-accumulate-grad hook releases the count, and the slice whose count reaches zero is all-reduced
    at once with async_op=True -- NCCL orders the collective behind the kernels already queued on the compute stream and
    runs it on its own stream while backward keeps launching.  finish() launches whatever is left (parameters that got no
    gradient this step), waits, and divides by the world size.  Slices are launched in a fixed order per rank only if the
    hooks fire in the same order on every rank, which holds for identical replicas (same autograd graph)."""

    def __init__(self, flat, params, offsets, bucket_bytes=32 << 20, world=None):
        self.flat = flat
        self.world = world or dist.get_world_size()
        step = max(1, bucket_bytes // flat.element_size())
        self.bounds = [(lo, min(lo + step, flat.numel())) for lo in range(0, flat.numel(), step)]
        self.members = [0] * len(self.bounds)
        self.param_slices = []
        for p, off in zip(params, offsets):
            first, last = off // step, (off + max(p.numel(), 1) - 1) // step
            sl = list(range(first, min(last, len(self.bounds) - 1) + 1))
            self.param_slices.append(sl)
            for i in sl:
                self.members[i] += 1
        self.handles = []
        self.left = list(self.members)
        self.launched = [False] * len(self.bounds)
        self.active = False
        for p, sl in zip(params, self.param_slices):
            p.register_post_accumulate_grad_hook(self._make_hook(sl))

    def _make_hook(self, sl):
        def hook(_p):
            if not self.active:
                return
            for i in sl:
                self.left[i] -= 1
                if self.left[i] == 0:
                    self._launch(i)
        return hook

    def _launch(self, i):
        if self.launched[i]:
            return
        lo, hi = self.bounds[i]
        self.handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        self.launched[i] = True

    def begin(self):
        self.left = list(self.members)
        self.launched = [False] * len(self.bounds)
        self.handles = []
        self.active = True

    def finish(self):
        self.active = False
        early = sum(self.launched)
        for i in range(len(self.bounds)):
            self._launch(i)
        for h in self.handles:
            h.wait()
        self.flat.div_(self.world)
        self.launched_early = early
        return len(self.bounds)


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
