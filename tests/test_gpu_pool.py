"""Pool-level parity (BASELINE.json configs[0] / north_star): the engine against what the UNMODIFIED reference returned
for a 100-image VOC2007-shaped pool (tests/golden/make_golden_pool.py): per-image consistency scores, class vectors, the
images selected at budget 10, and the RNG stream position -- Faster R-CNN R50-FPN and RetinaNet R50-FPN.

Hard bounds (no "most images" averages): the fraction of images within 1e-3, the worst deviation, identical selection.
Every image outside 1e-3 is explained stage by stage in profiles/r02_parity.md (tools/pool_diagnose.py)."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']


def _pool(kind):
    from cald_b200 import synth
    from cald_b200.engine import Engine, ARCH_FRCNN, ARCH_RETINANET
    tag = "frcnn_r50" if kind == "frcnn" else "retina_r50"
    path = os.path.join(GOLD, "pool_%s_nc21.npz" % tag)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % path)
    g = np.load(path)
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    w = synth.planted_frcnn_weights(50, 21, 0) if kind == "frcnn" else synth.planted_retinanet_weights(21, 0)
    eng = Engine(depth=50, num_classes=21, min_size=int(g["min_size"]), max_size=int(g["max_size"]),
                 arch_id=ARCH_FRCNN if kind == "frcnn" else ARCH_RETINANET)
    eng.load_state_dict(w)
    return g, imgs, eng


class _Labeled:
    def __init__(self, rows):
        self.rows = rows

    def __iter__(self):
        import torch
        for r in self.rows:
            yield (None,), ({"labels": torch.from_numpy(r[r >= 0])},)


@pytest.mark.parametrize("kind", ["frcnn", "retina"])
def test_pool_scores_selection_and_class_vectors(kind):
    from cald_b200 import api
    g, imgs, eng = _pool(kind)
    cons, cls = [], []
    for k, im in enumerate(imgs):
        random.seed(int(g["seeds"][k]))   # the fixture reseeds per image: one image's cutout draws do not move the next
        c, v = api.score_images(eng, [im], AUGS)
        cons.append(c[0])
        cls.append(v[0])
    cons, cls = np.array(cons), np.array(cls)
    err = np.abs(cons - g["consistency"])
    cerr = np.abs(cls - g["cls"]).max(axis=1)
    print("%s pool: |score - reference| median %.2e p90 %.2e max %.2e; > 1e-3: %s" % (
        kind, np.median(err), np.percentile(err, 90), err.max(), np.where(err > 1e-3)[0].tolist()))
    # the reference's own two runs (8 vs 3 intra-op threads) agree exactly on this pool, so every deviation is ours
    assert np.abs(g["noise_consistency"] - g["consistency"]).max() == 0.0
    assert np.median(err) <= 5e-6
    if kind == "frcnn":
        # measured (profiles/r02_parity.md): max 6.3e-6 over the 100 images
        assert err.max() <= 1e-4, (err.max(), np.where(err > 1e-4)[0])
        # image 37: two detections of the reference view whose scores differ by 8.5e-7 swap ranks 86 / 87, which moves
        # the 50-point linspace sub-sample (cald_train.py:110-113) to the other one; its score is unaffected
        assert (cerr > 1e-3).sum() <= 1, np.where(cerr > 1e-3)[0]
    else:
        # image 24 (1.27e-3): one class-0 box of the reference view sits on the NMS IoU threshold 0.5 to within 1e-6
        assert (err > 1e-3).sum() <= 1 and err.max() <= 2e-3, (err.max(), np.where(err > 1e-3)[0])
        assert (cerr > 1e-3).sum() <= 2, np.where(cerr > 1e-3)[0]
    # selection with the reference's inline code path on the engine's scores: the identical index set (north_star)
    sel = api.select(list(cons), [c for c in cls], list(g["subset"]), _Labeled(g["label_rows"]), int(g["budget"]))
    assert sorted(int(v) for v in sel) == sorted(int(v) for v in g["selected"])
    k = int(g["budget"])
    # (stable sort: the RetinaNet pool has 12 images that tie at exactly 0.0 -- no reference detections)
    assert set(np.argsort(cons, kind="stable")[:k]) == set(np.argsort(g["consistency"], kind="stable")[:k])
    eng.close()


@pytest.mark.parametrize("kind", ["frcnn", "retina"])
def test_pool_one_seed_batched_run_keeps_the_rng_stream(kind):
    """One seed for the whole pool, all images in one call (batched passes): every cutout accept / reject decision of
    100 images has to be the reference's for python's generator to end where the reference left it."""
    from cald_b200 import api
    g, imgs, eng = _pool(kind)
    random.seed(int(g["stream_seed"]))
    cons, cls = api.score_images(eng, imgs, AUGS)
    tail = random.random()
    err = np.abs(np.array(cons) - g["stream_consistency"])
    print("%s one-seed run: > 1e-3: %s max %.2e" % (kind, np.where(err > 1e-3)[0].tolist(), err.max()))
    assert tail == float(g["stream_rng_tail"])
    assert (err <= 1e-3).mean() >= 0.99
    eng.close()
