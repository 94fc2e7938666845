"""Parity at the shapes of BASELINE.json configs[0] (VOC2007 shape: 375x500 -> 600x800, nc = 21, R50, F,C,D,R) and
configs[3] (R101, VOC2012 shape, three augmentations), engine vs the CPU oracle restatement of the reference, plus
the selection step on the engine's scores (identical index set, north_star)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_scores(fwd, imgs, augs, nc, seeds):
    from oracle import cald_oracle as co
    cons, cls = [], []
    for im, s in zip(imgs, seeds):
        random.seed(s)
        c, v = co.score_image(fwd, im, augs, nc, 1.3)
        cons.append(float(c))
        cls.append(v)
    return cons, cls


def _engine_scores(eng, imgs, augs, seeds):
    from cald_b200 import api
    cons, cls = [], []
    for im, s in zip(imgs, seeds):
        random.seed(s)
        c, v = api.score_images(eng, [im], augs)
        cons.append(c[0])
        cls.append(v[0])
    return cons, cls


def test_cfg1_voc2007_shape_scores_and_selection():
    from cald_b200 import api, synth
    from cald_b200.engine import Engine
    from oracle import frcnn_oracle as fo
    augs = ['flip', 'cut_out', 'smaller_resize', 'rotation']
    w = synth.planted_frcnn_weights(50, 21, 0)
    wt = {k: torch.from_numpy(v) for k, v in w.items()}
    cfg = fo.Cfg(50, 21, 600, 1000)                                   # cald_train.py:340
    eng = Engine(depth=50, num_classes=21, min_size=600, max_size=1000, max_views_per_pass=8)
    eng.load_state_dict(w)
    imgs = [synth.synth_image(200 + i, 375, 500) for i in range(7)] + [synth.synth_image(300, 500, 375)]
    seeds = [5000 + i for i in range(len(imgs))]
    want, want_cls = _oracle_scores(lambda x: fo.forward(x, wt, cfg), imgs, augs, 21, seeds)
    got, got_cls = _engine_scores(eng, imgs, augs, seeds)
    err = np.abs(np.array(got) - np.array(want))
    print("cfg-1 engine", np.round(got, 5), "oracle", np.round(want, 5), "err", err)
    assert err.max() <= 1e-3 and np.median(err) <= 2e-5, err
    # selection: the k most inconsistent images (np.argsort ascending, cald_train.py:439-441) ...
    k = 3
    assert set(np.argsort(got)[:k]) == set(np.argsort(want)[:k])
    # ... and the full mutual-information step on both score sets picks the same images (cald_train.py:439-448)
    rs = np.random.RandomState(1)
    labeled = [[{"labels": torch.from_numpy(rs.randint(1, 21, rs.randint(1, 5)))}] for _ in range(10)]

    class LL:
        def __iter__(self):
            for t in labeled:
                yield (None,), tuple(t)
    subset = list(range(100, 100 + len(imgs)))
    a = api.select(got, got_cls, subset, LL(), 3)
    b = api.select(want, [np.asarray(v) for v in want_cls], subset, LL(), 3)
    assert sorted(int(v) for v in a) == sorted(int(v) for v in b)
    eng.close()


def test_cfg4_r101_three_augmentations():
    from cald_b200 import synth
    from cald_b200.engine import Engine
    from oracle import frcnn_oracle as fo
    augs = ['flip', 'cut_out', 'smaller_resize']                      # A = 3
    w = synth.planted_frcnn_weights(101, 21, 0)
    wt = {k: torch.from_numpy(v) for k, v in w.items()}
    cfg = fo.Cfg(101, 21, 600, 1000)
    eng = Engine(depth=101, num_classes=21, min_size=600, max_size=1000, max_views_per_pass=8)
    eng.load_state_dict(w)
    imgs = [synth.synth_image(400 + i, 375, 500) for i in range(3)]
    seeds = [6000 + i for i in range(len(imgs))]
    want, want_cls = _oracle_scores(lambda x: fo.forward(x, wt, cfg), imgs, augs, 21, seeds)
    got, got_cls = _engine_scores(eng, imgs, augs, seeds)
    err = np.abs(np.array(got) - np.array(want))
    print("cfg-4 engine", np.round(got, 5), "oracle", np.round(want, 5), "err", err)
    assert err.max() <= 1e-3, err
    eng.close()
