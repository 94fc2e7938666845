"""world_size-2 gloo test of the N>1 path: shard -> (fake) score -> all-gather -> loader order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeEngine:
    """stands in for the CUDA engine: score = mean pixel, class vector = per-channel stats"""
    num_classes = 4
    fetched = 0

    def images_per_chunk(self, n_augs):
        return 1  # => 4 images per scoring call: the shard is walked in several chunks

    def last_ref_counts(self, n):
        return np.ones(n, dtype=np.int32)

    def score(self, images, kinds, bp, u, noise=None, swap_perms=None):
        cons = np.array([float(im.mean()) for im in images])
        cls = np.stack([np.array([im[..., c].max() for c in range(3)], dtype=np.float64) for im in images])
        return cons, cls, 0


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cald_b200 import shard
    rs = np.random.RandomState(0)
    pool = [rs.randint(0, 256, (4, 5, 3)).astype(np.uint8) for _ in range(n)]
    cons, cls = shard.get_uncertainty_sharded(FakeEngine(), pool, ['flip'], rank, world)
    want = [float(im.mean()) for im in pool]
    ok = np.allclose(cons, want) and all(np.array_equal(c, [im[..., k].max() for k in range(3)])
                                         for c, im in zip(cls, pool))
    # the pool as an index -> image callable: a rank must only ever fetch the images of its own shard
    fetched = []

    def fetch(i):
        fetched.append(i)
        return pool[i]
    cons2, cls2 = shard.get_uncertainty_sharded(FakeEngine(), fetch, ['flip'], rank, world, n=n)
    ok = ok and cons2 == cons and all(np.array_equal(a, b) for a, b in zip(cls, cls2))
    ok = ok and fetched == list(range(rank, n, world))
    # the host selection on the gathered rows gives every rank the same picks (cald_train.py:439-448)
    from cald_b200 import api
    labeled = [((None,), ({"labels": torch.tensor([1, 2])},)), ((None,), ({"labels": torch.tensor([3])},))]
    picks = api.select(cons, cls, list(range(100, 100 + n)), labeled, 2)
    ret["picks%d" % rank] = [int(v) for v in picks]
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_allgather_restores_loader_order():
    for n in (7, 10):
        mgr = mp.Manager()
        ret = mgr.dict()
        port = 29511 + n
        mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]
        assert ret["picks0"] == ret["picks1"] and len(ret["picks0"]) == 2
