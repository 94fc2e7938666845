"""Pool sharding for multi-GPU scoring.

Images are independent, so the ordered index list the reference's SubsetSequentialSampler walks
(ll4al/data/sampler.py:3-16) is dealt out round-robin: rank r scores positions r, r+world, ...  Round-robin
(rather than contiguous blocks) keeps the shards balanced when the pool is ordered by shape.  The only
collective on the path is one all-gather of [n_local, 1 + (C-1)] values at the end; ``merge_shards`` puts
the gathered rows back into loader order.
"""
import numpy as np


def shard_indices(n, rank, world):
    return np.arange(rank, n, world)


def local_count(n, rank, world):
    return len(range(rank, n, world))


def padded_count(n, world):
    return (n + world - 1) // world


def merge_shards(parts, n, world):
    """parts[r] = rows scored by rank r (in its local order) -> array in global loader order."""
    first = np.asarray(parts[0])
    out = np.zeros((n,) + first.shape[1:], dtype=first.dtype)
    for r in range(world):
        idx = shard_indices(n, r, world)
        out[idx] = np.asarray(parts[r])[:len(idx)]
    return out


def get_uncertainty_sharded(eng, pool, augs, rank, world, group=None, n=None, chunk=None):
    """Score this rank's shard, all-gather (consistency, class vector) rows, return them in loader order on every
    rank.  ``pool`` is either a sequence of u8 HWC images or a callable ``index -> image`` together with ``n`` (the
    pool size): a rank only ever materialises the images of its own shard, one scoring chunk at a time, so neither
    the pool nor the shard has to fit in host memory (cfg-5's 118k images are 378 GB).

    Python's ``random`` stream (cutout) is consumed per rank in shard order; seed it per rank if draws matter."""
    import torch
    import torch.distributed as dist
    from . import api
    fetch = pool if callable(pool) else (lambda i: pool[i])
    n = len(pool) if n is None else n
    mine = shard_indices(n, rank, world)
    step = chunk or 2 * eng.images_per_chunk(max(1, len(api._aug_kinds(augs))))
    cons, cls = [], []
    for pos in range(0, len(mine), step):
        c, v = api.score_images(eng, [api._to_u8(fetch(int(i))) for i in mine[pos:pos + step]], augs)
        cons.extend(c)
        cls.extend(v)
    c1 = eng.num_classes - 1
    rows = np.zeros((padded_count(n, world), 1 + c1), dtype=np.float64)
    if len(mine):
        rows[:len(mine), 0] = cons
        rows[:len(mine), 1:] = np.stack(cls) if len(cls) else 0
    if world == 1:
        parts = [rows]
    else:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.from_numpy(rows).to(dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        parts = [o.cpu().numpy() for o in out]
    merged = merge_shards(parts, n, world)
    return [float(v) for v in merged[:, 0]], [merged[i, 1:].copy() for i in range(n)]
