"""On-device selection (cald_select) against the reference's own picks (fixtures written by the
unmodified cald_train.cls_kldiv / inline selection) and against the host path api.select at cfg-5 size."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Labeled:
    def __init__(self, rows):
        self.rows = rows

    def __iter__(self):
        for r in self.rows:
            yield (None,), ({"labels": torch.from_numpy(r[r >= 0])},)


@pytest.fixture(scope="module")
def eng():
    from cald_b200.engine import Engine
    return Engine(depth=50, num_classes=21, min_size=320, max_size=512, workspace_bytes=1 << 30)


def test_reference_selection_fixture(eng):
    from cald_b200 import api
    g = np.load(os.path.join(GOLD, "selection.npz"))
    subset = [int(v) for v in g["subset"]]
    new = api.select(list(g["uncertainty"]), [c for c in g["cls"]], subset, Labeled(g["labels"]), int(g["budget"]),
                     engine=eng)
    assert [int(v) for v in new] == [int(v) for v in g["new_labeled"]]      # same images in the same pick order


@pytest.mark.parametrize("tag", ["frcnn_r50", "retina_r50"])
def test_pool_fixture_selection(eng, tag):
    """the 100-image pools: the RetinaNet one has 12 images without detections (all-zero class vectors, score 0.0):
    more than the budget of 10, and cls_kldiv returns all of them (cald_train.py:246-249)"""
    from cald_b200 import api
    g = np.load(os.path.join(GOLD, "pool_%s_nc21.npz" % tag))
    new = api.select(list(g["consistency"]), [c for c in g["cls"]], [int(v) for v in g["subset"]],
                     Labeled(g["label_rows"]), int(g["budget"]), engine=eng)
    assert sorted(int(v) for v in new) == sorted(int(v) for v in g["selected"])
    host = api.select(list(g["consistency"]), [c for c in g["cls"]], [int(v) for v in g["subset"]],
                      Labeled(g["label_rows"]), int(g["budget"]))
    assert sorted(int(v) for v in new) == sorted(int(v) for v in host)


@pytest.mark.parametrize("uniform", [False, True])
def test_cfg5_size_matches_host_path(eng, uniform):
    """118k-image COCO-shaped pool, budget 1000, mr 1.2 (cald_train.py:306, 516): device picks == host picks"""
    from cald_b200 import api
    rs = np.random.RandomState(7)
    n, c1, budget = 118000, 90, 1000
    unc = rs.uniform(0, 1, n)
    cls = rs.uniform(0, 1, (n, c1)) * (rs.uniform(0, 1, (n, c1)) > 0.8)
    zero = np.argsort(unc)[[3, 500, 1100]]
    cls[zero] = 0.0
    rows = np.full((300, 8), -1, dtype=np.int64)
    for r in rows:
        k = rs.randint(1, 8)
        r[:k] = rs.randint(1, c1 + 1, k)
    subset = list(range(5000, 5000 + n))
    old = api.uniform
    api.uniform = uniform
    try:
        host = api.select(list(unc), [c for c in cls], subset, Labeled(rows), budget)
        dev = api.select(list(unc), [c for c in cls], subset, Labeled(rows), budget, engine=eng)
    finally:
        api.uniform = old
    assert [int(v) for v in dev] == [int(v) for v in host]
    assert len(dev) == budget
