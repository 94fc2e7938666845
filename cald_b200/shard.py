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


def get_uncertainty_sharded(eng, images_u8, augs, rank, world, group=None):
    """Score this rank's shard, all-gather (consistency, class vector) rows, return them in loader order
    on every rank.  ``images_u8`` is the FULL ordered pool (each rank only touches its shard)."""
    import torch
    import torch.distributed as dist
    from . import api
    n = len(images_u8)
    mine = shard_indices(n, rank, world)
    cons, cls = api.score_images(eng, [images_u8[i] for i in mine], augs)
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
